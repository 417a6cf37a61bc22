"""Sharded search behind the C ABI (row e; service.rs:777-801, search.rs:211-237): G node-range shards, each with
its own sub-graph, every query on every shard, per-shard top-k as packed 16-byte records, merge by (dist, id).

Oracle side: orc.leann_search on each sub-graph + orc.merge_topk over global ids.  GPU side, on ONE device:
  * the G shards' halves (isl_index_search_packed_dev) written into slot g of a [G][nq][k] record buffer and merged
    by isl_merge_packed_dev — the exact kernels a G-rank run uses either side of the exchange;
  * a real communicator of world size 1 (NCCL inside the library): isl_index_search_sharded / _dev end to end,
    with ncclAllGather and with the peer-store exchange.
The G > 1 exchange itself needs G GPUs: tests/multi_gpu/run_sharded_parity.py (torchrun) checks it against the
same oracle; its committed result is profiles/r02_sharded_parity_*.txt."""
import os

import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu

INVALID = np.uint64(0xFFFFFFFFFFFFFFFF)


def _shards(orc, n, d, G, seed, dup=0):
    from islands_b200 import LeannConfig
    from islands_b200.shard import shard_range

    rng = np.random.RandomState(seed)
    x = uniform(rng, n, d)
    if dup:
        x[n - dup:] = x[:dup]  # exact ties across shards: the (dist, id) rule must decide identically
    cfg = LeannConfig()
    out = []
    for g in range(G):
        lo, hi = shard_range(n, g, G)
        levels = orc.draw_levels(seed + 10 + g, hi - lo, cfg.ml, cfg.max_layers)
        off, nbrs, entry, _ = orc.leann_build(cfg._s, x[lo:hi], levels, batch=16, threads=os.cpu_count() or 1)
        out.append((lo, hi, levels, off, nbrs, entry))
    return cfg, x, out


def _oracle_merged(orc, cfg, x, shards, q, k, ef):
    ids_l, dst_l = [], []
    for lo, hi, levels, off, nbrs, entry in shards:
        ids, dst, _ = orc.leann_search(cfg._s, x[lo:hi], off, nbrs, entry, q, k, ef, threads=os.cpu_count() or 1)
        ids_l.append(np.where(ids == INVALID, INVALID, ids + np.uint64(lo)))
        dst_l.append(dst)
    return orc.merge_topk(np.stack(ids_l), np.stack(dst_l), k)


@pytest.mark.parametrize("G", [2, 4, 8])
def test_emulated_shards_equal_oracle_merge(gpu_lib, orc, G):
    import torch

    from islands_b200 import LeannIndex
    from islands_b200.core import merge_packed_dev

    n, d, nq, k, ef = 2400, 48, 96, 10, 40
    cfg, x, shards = _shards(orc, n, d, G, seed=21 + G, dup=200)
    q = np.concatenate([x[:32], uniform(np.random.RandomState(3), nq - 32, d)])  # queries equal to stored (duplicated) rows
    o_ids, o_dst, o_cnt = _oracle_merged(orc, cfg, x, shards, q, k, ef)

    dev = torch.device("cuda", 0)
    tq = torch.from_numpy(q).to(dev)
    rec = torch.empty((G, nq, k, 16), dtype=torch.uint8, device=dev)
    handles = []
    for g, (lo, hi, levels, off, nbrs, entry) in enumerate(shards):
        idx = LeannIndex.from_csr(cfg, x[lo:hi], off, nbrs, levels, entry)
        idx.search_packed_dev(lo, tq.data_ptr(), nq, d, k, ef, rec[g].data_ptr())
        handles.append(idx)
    m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    m_dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
    m_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    merge_packed_dev(rec.data_ptr(), G, nq, k, m_ids.data_ptr(), m_dst.data_ptr(), m_cnt.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(m_cnt.cpu().numpy().astype(np.uint32), o_cnt)
    assert np.array_equal(m_ids.cpu().numpy().view(np.uint64), o_ids)
    assert np.array_equal(m_dst.cpu().numpy().view(np.uint32), o_dst.view(np.uint32))
    # the records of every shard are that shard's own search result with global ids
    r = rec.cpu().numpy().view(np.dtype([("dist", "<f4"), ("reserved", "<u4"), ("id", "<u8")])).reshape(G, nq, k)
    for g, (lo, hi, levels, off, nbrs, entry) in enumerate(shards):
        ids, dst, _ = orc.leann_search(cfg._s, x[lo:hi], off, nbrs, entry, q, k, ef, threads=os.cpu_count() or 1)
        assert np.array_equal(r[g]["id"], np.where(ids == INVALID, INVALID, ids + np.uint64(lo)))
        assert np.array_equal(r[g]["dist"].view(np.uint32), dst.view(np.uint32))
        assert not r[g]["reserved"].any()
    for h in handles:
        h.free()


@pytest.mark.parametrize("peer", [False, True])
def test_world1_communicator_end_to_end(gpu_lib, orc, peer):
    """NCCL inside the library on one rank: search -> exchange -> merge on one stream equals the plain search with
    global ids; host-buffer and device-buffer entry points; ncclAllGather and the peer-store exchange."""
    import torch

    from islands_b200 import LeannIndex
    from islands_b200.core import ShardComm
    from islands_b200.shard import ShardedLeannIndex

    n, d, nq, k, ef, base = 1500, 32, 64, 10, 48, 1_000_000_000_000  # ids beyond 2^32: the u64 id path
    cfg, x, shards = _shards(orc, n, d, 1, seed=5)
    lo, hi, levels, off, nbrs, entry = shards[0]
    q = uniform(np.random.RandomState(4), nq, d)
    o_ids, o_dst, o_cnt = orc.leann_search(cfg._s, x, off, nbrs, entry, q, k, ef)
    idx = LeannIndex.from_csr(cfg, x, off, nbrs, levels, entry)
    comm = ShardComm(0, 1, ShardComm.unique_id())
    if peer:
        comm.enable_peer_exchange(nq * k)
    sharded = ShardedLeannIndex(idx, base, n, comm)
    for _ in range(3):  # several steps: the peer exchange alternates its two gather buffers
        ids, dst, cnt = sharded.search_batch(q, k, ef)
        assert np.array_equal(cnt, o_cnt)
        assert np.array_equal(ids, o_ids + np.uint64(base))
        assert np.array_equal(dst.view(np.uint32), o_dst.view(np.uint32))
    dev = torch.device("cuda", 0)
    tq = torch.from_numpy(q).to(dev)
    t_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    t_dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
    t_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    sharded.search_batch_dev(tq, k, ef, t_ids, t_dst, t_cnt)
    assert np.array_equal(t_ids.cpu().numpy().view(np.uint64), o_ids + np.uint64(base))
    assert np.array_equal(t_dst.cpu().numpy().view(np.uint32), o_dst.view(np.uint32))
    s_ms, x_ms, m_ms = comm.last_timing()
    assert s_ms > 0 and x_ms >= 0 and m_ms > 0
    # an empty shard still takes part in the exchange and contributes nothing
    empty = LeannIndex.from_csr(cfg, np.zeros((0, d), np.float32), [0], [], None, None)
    ids, dst, cnt = empty.search_sharded(comm, 0, q, k, ef)
    assert (ids == INVALID).all() and np.isinf(dst).all() and not cnt.any()
    comm.free()
    idx.free()


def test_sharded_argument_errors(gpu_lib, orc):
    from islands_b200 import DimensionMismatch, InvalidArgument, LeannIndex
    from islands_b200.core import ShardComm

    cfg, x, shards = _shards(orc, 300, 16, 1, seed=6)
    lo, hi, levels, off, nbrs, entry = shards[0]
    idx = LeannIndex.from_csr(cfg, x, off, nbrs, levels, entry)
    comm = ShardComm(0, 1, ShardComm.unique_id())
    with pytest.raises(DimensionMismatch):
        idx.search_sharded(comm, 0, np.zeros((2, 8), np.float32), 5, 16)
    with pytest.raises(InvalidArgument):
        ShardComm(3, 2, ShardComm.unique_id())
    comm.free()
    idx.free()


def test_real_multi_gpu_exchange_when_available(gpu_lib):
    """With >= 2 GPUs on the box: the torchrun parity program (NCCL and peer-store exchange) must exit 0."""
    import socket
    import subprocess
    import sys

    g = gpu_lib.isl_device_count()
    if g < 2:
        pytest.skip("one GPU: the cross-GPU exchange is covered by tests/multi_gpu/run_sharded_parity.py under gpurun --gpus N")
    world = min(g, 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tests", "multi_gpu", "run_sharded_parity.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"bit_exact_on_every_rank": true' in r.stdout
