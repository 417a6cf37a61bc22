"""Index sharding across the GPUs of one box: node-range shards (or one island per GPU, the
reference's `graphs: HashMap<String, StoredIndex>` loop, src/indexer/service.rs:777-797), every
rank searches its shard with the same queries, ONE all-gather of the per-shard (dist, id) lists
over NCCL, then a per-query merge by (dist, id) (src/core/search.rs:211-237).

torch.distributed is plumbing here (process group, all-gather); search and merge are the CUDA
kernels behind the C ABI.  The collective also runs on gloo/CPU tensors, which is how the host
logic is tested without GPUs.
"""
import torch
import torch.distributed as dist

from .core import merge_topk_dev


def shard_range(n, rank, world):
    """Contiguous node range [lo, hi) owned by `rank`; ranges tile [0, n) exactly."""
    return (rank * n) // world, ((rank + 1) * n) // world


def local_to_global(ids, base):
    """Shard-local ids -> global ids; the all-ones padding (ISL_INVALID_ID == -1 as int64) stays."""
    return torch.where(ids >= 0, ids + base, ids)


def gather_topk(ids, dist_, group=None):
    """All-gather of per-shard lists: ids [nq,k] int64, dist [nq,k] f32 -> ([W,nq,k], [W,nq,k])."""
    world = dist.get_world_size(group)
    nq = ids.shape[0]
    # the output is the concatenation along dim 0 (the form gloo and NCCL both accept)
    g_ids = torch.empty((world * nq,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    g_dst = torch.empty((world * nq,) + tuple(dist_.shape[1:]), dtype=dist_.dtype, device=dist_.device)
    dist.all_gather_into_tensor(g_ids, ids.contiguous(), group=group)
    dist.all_gather_into_tensor(g_dst, dist_.contiguous(), group=group)
    return g_ids.view((world,) + tuple(ids.shape)), g_dst.view((world,) + tuple(dist_.shape))


class ShardedLeannIndex:
    """One shard of a node-range-sharded LeannIndex living on this rank's GPU."""

    def __init__(self, index, base, n_total, group=None):
        self.index = index
        self.base = int(base)
        self.n_total = int(n_total)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def search_batch_dev(self, q, k, ef, ids, dst, cnt, out_ids, out_dst, stats=None):
        """q [nq,d] f32 on this GPU; ids/dst/cnt: local scratch; out_ids/out_dst [nq,k]: merged result
        (identical on every rank).  Returns the tensor holding the final ids."""
        nq, d = q.shape
        self.index.search_batch_dev(q.data_ptr(), nq, d, k, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(),
                                    stats.data_ptr() if stats is not None else None)
        if self.world == 1:
            return ids, dst
        g_ids, g_dst = gather_topk(local_to_global(ids, self.base), dst, self.group)
        torch.cuda.current_stream().synchronize()  # the merge runs on the library's stream
        merge_topk_dev(g_ids.data_ptr(), g_dst.data_ptr(), self.world, nq, k, out_ids.data_ptr(), out_dst.data_ptr())
        return out_ids, out_dst
