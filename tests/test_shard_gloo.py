"""N>1 host logic on CPU (gloo, world_size 2): node-range sharding, local->global ids, the single
all-gather of packed 16-byte isl_shard_record entries (the format the library exchanges with NCCL),
and the merge semantics (checked with the oracle's merge; the product's merge is a CUDA kernel and
is covered by the GPU tests: tests/test_sharded_search.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, nq, k, ef, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from islands_b200 import LeannConfig
    from islands_b200.shard import gather_records, pack_records, shard_range
    from oracle import pyoracle as orc

    rng = np.random.RandomState(0)
    x = (rng.rand(n, d).astype(np.float32) * 2 - 1)
    q = (rng.rand(nq, d).astype(np.float32) * 2 - 1)
    lo, hi = shard_range(n, rank, world)
    cfg = LeannConfig()
    levels = orc.draw_levels(3, hi - lo, cfg.ml, cfg.max_layers)
    off, nbrs, entry, _ = orc.leann_build(cfg._s, x[lo:hi], levels, batch=8)
    # per-shard search (the oracle stands in for the GPU kernel on this CPU-only box)
    ids, dst, _ = orc.leann_search(cfg._s, x[lo:hi], off, nbrs, entry, q, k, ef)
    rec = pack_records(ids, dst, base=lo)
    g = gather_records(rec)  # ONE all-gather of the packed records
    assert g.shape == (world, nq, k) and g.dtype.itemsize == 16
    g_ids, g_dst = np.ascontiguousarray(g["id"]), np.ascontiguousarray(g["dist"])
    # every rank holds the same gathered lists, laid out [parts][nq][k]
    chk = torch.from_numpy(g_ids.astype(np.int64))
    ref = chk.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(chk, ref)
    assert np.array_equal(g[rank], rec)
    assert np.array_equal(g_ids[rank][ids != np.uint64(0xFFFFFFFFFFFFFFFF)], (ids + np.uint64(lo))[ids != np.uint64(0xFFFFFFFFFFFFFFFF)])
    m_ids, m_dst, m_cnt = orc.merge_topk(g_ids, g_dst, k)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), m_ids)
    if rank == 0:
        np.save(os.path.join(out_dir, "dist.npy"), m_dst)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_gloo_world2(tmp_path):
    n, d, nq, k, ef, world = 600, 16, 40, 5, 32, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, d, nq, k, ef, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "merged_0.npy")
    b = np.load(tmp_path / "merged_1.npy")
    assert np.array_equal(a, b)  # identical merged result on every rank
    # merged ids are global, in range, sorted by distance, and recall vs brute force is sane
    rng = np.random.RandomState(0)
    x = (rng.rand(n, d).astype(np.float32) * 2 - 1)
    q = (rng.rand(nq, d).astype(np.float32) * 2 - 1)
    assert a.max() < n
    md = np.load(tmp_path / "dist.npy")
    assert (np.diff(md, axis=1) >= 0).all()
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    gt = np.argsort(-(qn @ xn.T), axis=1)[:, :k]
    recall = np.mean([len(set(a[i]) & set(gt[i])) / k for i in range(nq)])
    assert recall > 0.8, recall


def test_shard_ranges_tile_the_index():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from islands_b200.shard import local_to_global, shard_range

    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, i, world) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    ids = torch.tensor([[0, 5, -1]], dtype=torch.int64)
    assert local_to_global(ids, 100).tolist() == [[100, 105, -1]]
