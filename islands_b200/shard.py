"""Index sharding across the GPUs of one box: node-range shards (or one island per GPU, the
reference's `graphs: HashMap<String, StoredIndex>` loop, src/indexer/service.rs:777-797).  Every
rank searches its shard with the same queries, the per-shard top-k lists travel as packed 16-byte
(dist, global id) records in ONE exchange, then a per-query merge by (dist, id)
(src/core/search.rs:211-237).

The product path is entirely behind the C ABI (`isl_index_search_sharded*`, csrc/api_shard.cu):
search kernel, NCCL all-gather (or peer stores over NVLink) and merge kernel on one stream inside
the library.  torch.distributed is only the host channel that hands NCCL's unique id to the ranks.
The helpers below (ranges, id mapping, record packing, the one-all-gather exchange) are the host
logic; they also run on gloo/CPU tensors, which is how they are tested without GPUs.
"""
import numpy as np
import torch
import torch.distributed as dist

from .core import ShardComm

# isl_shard_record (include/islands_b200.h): {f32 dist, u32 reserved, u64 id}
RECORD_DTYPE = np.dtype([("dist", "<f4"), ("reserved", "<u4"), ("id", "<u8")])
INVALID_ID = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_range(n, rank, world):
    """Contiguous node range [lo, hi) owned by `rank`; ranges tile [0, n) exactly."""
    return (rank * n) // world, ((rank + 1) * n) // world


def local_to_global(ids, base):
    """Shard-local ids -> global ids; the all-ones padding (ISL_INVALID_ID == -1 as int64) stays."""
    return torch.where(ids >= 0, ids + base, ids)


def pack_records(ids, dist_, base=0):
    """ids [nq,k] u64 (ISL_INVALID_ID padded), dist [nq,k] f32 -> isl_shard_record array [nq,k] with global ids."""
    ids = np.asarray(ids, np.uint64)
    rec = np.zeros(ids.shape, RECORD_DTYPE)
    rec["dist"] = np.asarray(dist_, np.float32)
    rec["id"] = np.where(ids == INVALID_ID, INVALID_ID, ids + np.uint64(base))
    return rec


def gather_records(rec, group=None):
    """ONE all-gather of the packed records: [nq,k] -> [world,nq,k] (the library does the same with
    ncclAllGather on bytes; here it runs over whatever backend the group has, gloo included)."""
    world = dist.get_world_size(group)
    mine = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
    out = torch.empty((world * mine.numel(),), dtype=torch.uint8)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.numpy().view(RECORD_DTYPE).reshape((world,) + rec.shape)


def make_shard_comm(group=None, device=None):
    """Creates the library's communicator for the ranks of `group`: rank 0 draws NCCL's unique id, the host
    channel (torch.distributed, any backend) broadcasts its 128 bytes, every rank joins."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ShardComm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group, device=device)
    return ShardComm(rank, world, box[0])


class ShardedLeannIndex:
    """One shard of a node-range- or island-sharded LeannIndex living on this rank's GPU."""

    def __init__(self, index, base, n_total, comm):
        self.index = index
        self.base = int(base)
        self.n_total = int(n_total)
        self.comm = comm

    def search_batch_dev(self, q, k, ef, out_ids, out_dst, out_cnt=None):
        """q [nq,d] f32 on this GPU (the same on every rank); out_ids [nq,k] int64 / out_dst [nq,k] f32 receive
        the merged result, identical on every rank.  One library call: search -> exchange -> merge."""
        nq, d = q.shape
        self.index.search_sharded_dev(self.comm, self.base, q.data_ptr(), nq, d, k, ef, out_ids.data_ptr(), out_dst.data_ptr(),
                                      out_cnt.data_ptr() if out_cnt is not None else None)
        return out_ids, out_dst

    def search_batch(self, queries, k, ef):
        """Host buffers in, host results out (numpy)."""
        return self.index.search_sharded(self.comm, self.base, queries, k, ef)
