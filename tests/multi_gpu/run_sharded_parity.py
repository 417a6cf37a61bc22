"""Sharded search on G real GPUs vs the oracle (launched with torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
      tests/multi_gpu/run_sharded_parity.py [--out FILE]

Every rank builds ALL G oracle sub-graphs (small, deterministic), adopts its own with from_csr, and runs
isl_index_search_sharded (host buffers) and _dev, first with ncclAllGather, then with the peer-store exchange.
The merged ids / distances / counts on every rank must equal the oracle's per-shard searches merged by
orc.merge_topk over global ids — bit for bit.  Rank 0 prints and (with --out) writes one result line."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--nodes", type=int, default=4000)
    ap.add_argument("--dim", type=int, default=64)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")  # host channel only: the data path's NCCL communicator lives in the library

    from islands_b200 import LeannConfig, LeannIndex
    from islands_b200.shard import ShardedLeannIndex, make_shard_comm, shard_range
    from oracle import pyoracle as orc

    n, d, nq, k, ef = a.nodes, a.dim, 256, 10, 48
    rng = np.random.RandomState(17)
    x = (rng.rand(n, d).astype(np.float32) * 2 - 1)
    x[n - 300:] = x[:300]  # duplicates across shards: exact distance ties in the merge
    q = np.concatenate([x[:64], (np.random.RandomState(18).rand(nq - 64, d).astype(np.float32) * 2 - 1)])
    cfg = LeannConfig()
    INVALID = np.uint64(0xFFFFFFFFFFFFFFFF)
    ids_l, dst_l, mine = [], [], None
    threads = max(1, (os.cpu_count() or 1) // world)
    for g in range(world):
        lo, hi = shard_range(n, g, world)
        levels = orc.draw_levels(100 + g, hi - lo, cfg.ml, cfg.max_layers)
        off, nbrs, entry, _ = orc.leann_build(cfg._s, x[lo:hi], levels, batch=16, threads=threads)
        ids, dst, _ = orc.leann_search(cfg._s, x[lo:hi], off, nbrs, entry, q, k, ef, threads=threads)
        ids_l.append(np.where(ids == INVALID, INVALID, ids + np.uint64(lo)))
        dst_l.append(dst)
        if g == rank:
            mine = (lo, hi, levels, off, nbrs, entry)
    o_ids, o_dst, o_cnt = orc.merge_topk(np.stack(ids_l), np.stack(dst_l), k)
    # the same for "PQ ADC traversal + exact rerank": one quantizer for the whole index, per-shard codes
    from islands_b200 import PQConfig, ProductQuantizer

    pq = ProductQuantizer(d, PQConfig(16, 32, 6, 5))
    pq.train(x[:2000])
    books = pq.codebooks()
    codes = pq.encode(x)
    a_ids, a_dst = [], []
    for g in range(world):
        lo, hi = shard_range(n, g, world)
        levels = orc.draw_levels(100 + g, hi - lo, cfg.ml, cfg.max_layers)
        off, nbrs, entry, _ = orc.leann_build(cfg._s, x[lo:hi], levels, batch=16, threads=threads)
        ids, dst, _ = orc.leann_search_adc_rerank(cfg._s, x[lo:hi], off, nbrs, entry, books, codes[lo:hi], q, k, ef, threads=threads)
        a_ids.append(np.where(ids == INVALID, INVALID, ids + np.uint64(lo)))
        a_dst.append(dst)
    oa_ids, oa_dst, oa_cnt = orc.merge_topk(np.stack(a_ids), np.stack(a_dst), k)

    lo, hi, levels, off, nbrs, entry = mine
    idx = LeannIndex.from_csr(cfg, x[lo:hi], off, nbrs, levels, entry)
    idx.attach_pq(pq, codes[lo:hi])
    comm = make_shard_comm()
    sharded = ShardedLeannIndex(idx, lo, n, comm)
    tq = torch.from_numpy(q).to(dev)
    t_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    t_dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
    t_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    ok, timing = True, {}
    for engine in ("nccl", "peer"):
        if engine == "peer":
            comm.enable_peer_exchange(nq * k)
        for step in range(4):
            ids, dst, cnt = sharded.search_batch(q, k, ef)
            ok &= bool(np.array_equal(ids, o_ids) and np.array_equal(dst.view(np.uint32), o_dst.view(np.uint32)) and np.array_equal(cnt, o_cnt))
            sharded.search_batch_dev(tq, k, ef, t_ids, t_dst, t_cnt)
            ok &= bool(np.array_equal(t_ids.cpu().numpy().view(np.uint64), o_ids)
                       and np.array_equal(t_dst.cpu().numpy().view(np.uint32), o_dst.view(np.uint32))
                       and np.array_equal(t_cnt.cpu().numpy().astype(np.uint32), o_cnt))
            ids, dst, cnt = idx.search_sharded_adc(comm, lo, q, k, ef)  # sharded ADC traversal + exact rerank
            ok &= bool(np.array_equal(ids, oa_ids) and np.array_equal(dst.view(np.uint32), oa_dst.view(np.uint32)) and np.array_equal(cnt, oa_cnt))
        sharded.search_batch_dev(tq, k, ef, t_ids, t_dst, t_cnt)
        timing[engine] = [round(v, 4) for v in comm.last_timing()]
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        line = {"test": "sharded search (exact and ADC traversal + rerank) vs oracle sub-graph searches + orc_merge_topk", "world": world, "n": n, "d": d, "nq": nq, "k": k, "ef": ef,
                "bit_exact_on_every_rank": all(flags), "per_rank": flags,
                "last_call_ms_rank0 (search, exchange, merge)": timing}
        print(json.dumps(line))
        if a.out:
            with open(a.out, "w") as f:
                f.write(json.dumps(line) + "\n")
    comm.free()
    dist.barrier()
    dist.destroy_process_group()
    if not all(flags):
        sys.exit(1)


if __name__ == "__main__":
    main()
