#!/usr/bin/env python
"""bench.py — LEANN search hot path on B200: QPS at recall@10 >= 0.95 (1M x 768, M=30, top-10).

One step = one batch of `nq` queries through the batched best-first search of the resident
index (islands_b200 C ABI).  See DESIGN.md "measurement" for every definition used here.

  python bench.py --gpus 1 --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference ...                     # the reference's CPU algorithm (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...        # index sharded by node range + NCCL merge
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_TOP = 10
RECALL_TARGET = 0.95
EF_LADDER = [16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="islands_b200", choices=["islands_b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--dataset", type=str, default="latent32", choices=["latent32", "uniform"])
    ap.add_argument("--ef", type=int, default=0, help="0 = smallest ef of the ladder with recall@10 >= 0.95")
    ap.add_argument("--build-batch", type=int, default=4096)
    ap.add_argument("--no-uniform", action="store_true", help="skip the secondary uniform-data measurement")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline sample")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic data (generated on the GPU; seeds fixed)
# ---------------------------------------------------------------------------------------------
def make_data(torch, dataset, n, nq, d, dev):
    """latent32: x = z A + 0.05 eps, z ~ N(0, I_32), A ~ N(0, 1/32)^{32 x d}: d-dimensional vectors
    with the low intrinsic dimension typical of learned embeddings (graph ANN reaches 0.95 recall).
    uniform: i.i.d. U[-1,1), the distribution of the reference's criterion benches
    (benches/hnsw_benchmarks.rs:9-14) — a worst case for ANY graph index at d=768."""
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    if dataset == "latent32":
        a = torch.randn((32, d), generator=g, device=dev) / 32 ** 0.5
        x = torch.randn((n, 32), generator=g, device=dev) @ a + 0.05 * torch.randn((n, d), generator=g, device=dev)
        g.manual_seed(43)
        q = torch.randn((nq, 32), generator=g, device=dev) @ a + 0.05 * torch.randn((nq, d), generator=g, device=dev)
    else:
        x = torch.rand((n, d), generator=g, device=dev) * 2 - 1
        g.manual_seed(43)
        q = torch.rand((nq, d), generator=g, device=dev) * 2 - 1
    return x.contiguous(), q.contiguous()


def ground_truth(torch, x, q, k, chunk=1024):
    """Exact cosine top-k by brute force (measurement infrastructure, not the hot path)."""
    xn = torch.nn.functional.normalize(x, dim=1)
    qn = torch.nn.functional.normalize(q, dim=1)
    out = [(qn[s:s + chunk] @ xn.T).topk(k, dim=1).indices for s in range(0, q.shape[0], chunk)]
    return torch.cat(out)


def recall_at_k(torch, ids, gt):
    """ids, gt: [m, k] int64 on the same device."""
    hit = (ids.unsqueeze(2) == gt.unsqueeze(1)).any(dim=2).float().sum(dim=1)
    return float((hit / gt.shape[1]).mean().item())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nme, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def algorithmic_bytes(stats, d, nq, k):
    """DESIGN.md / SURVEY §8(d): B = n_dist*4d + n_edge*4 + n_hop*16 + nq*(4d + 12k)."""
    nh, ne, nd = int(stats[:, 0].sum()), int(stats[:, 1].sum()), int(stats[:, 2].sum())
    return nd * 4 * d + ne * 4 + nh * 16 + nq * (4 * d + 12 * k), dict(n_hop=nh / nq, n_edge=ne / nq, n_dist=nd / nq)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def cpu_port_qps(orc, cfg, xh, off, nbrs, entry, qh, ef, threads, seconds):
    """Oracle (CPU port of leann.rs:868-988) on a bounded sample, all host threads."""
    probe = min(qh.shape[0], 4 * threads)
    t0 = time.perf_counter()
    orc.leann_search(cfg._s, xh, off, nbrs, entry, qh[:probe], K_TOP, ef, threads=threads)
    dt = time.perf_counter() - t0
    m = int(max(probe, min(qh.shape[0], probe * seconds / max(dt, 1e-6))))
    t0 = time.perf_counter()
    ids, _, _ = orc.leann_search(cfg._s, xh, off, nbrs, entry, qh[:m], K_TOP, ef, threads=threads)
    dt = time.perf_counter() - t0
    return m / dt, m, ids


def main():
    a = parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and rank != 0:
        return  # the CPU arm runs on rank 0 alone
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: islands_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1 and a.impl != "reference"
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from islands_b200 import LeannConfig, LeannIndex, _ffi
    from islands_b200.shard import ShardedLeannIndex, shard_range

    lib = _ffi.load()
    cfg = LeannConfig()  # paper_default: m=30, m0=60, efC=128, cosine, hub-preserving pruning 2%
    n, d, nq = a.n, a.d, a.nq
    parts = world if use_dist else 1
    lo, hi = shard_range(n, rank if use_dist else 0, parts)

    x, q = make_data(torch, a.dataset, n, nq, d, dev)
    n_gt = min(nq, 1000)
    gt = ground_truth(torch, x, q[:n_gt], K_TOP)
    shard = x[lo:hi].contiguous()
    del x
    torch.cuda.empty_cache()

    t0 = time.perf_counter()
    index = LeannIndex(cfg)
    index.build_dev(shard.data_ptr(), hi - lo, d, seed=7, batch=a.build_batch)
    build_s = time.perf_counter() - t0

    ids = torch.empty((nq, K_TOP), dtype=torch.int64, device=dev)
    dst = torch.empty((nq, K_TOP), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    stats = torch.zeros((nq, 5), dtype=torch.int64, device=dev)
    m_ids = torch.empty((nq, K_TOP), dtype=torch.int64, device=dev)
    m_dst = torch.empty((nq, K_TOP), dtype=torch.float32, device=dev)
    sharded = ShardedLeannIndex(index, lo, n)

    def step_device(nqq, ef, with_stats=False):
        """One pass of the hot path with inputs resident in HBM: per-shard search, then (N > 1) one
        all-gather of the (dist, id) lists and the per-query merge.  Returns the final ids tensor."""
        out_ids, _ = sharded.search_batch_dev(q[:nqq], K_TOP, ef, ids, dst, cnt, m_ids, m_dst, stats if with_stats else None)
        return out_ids

    # ---- ef: smallest rung with recall@10 >= 0.95 (setup, untimed) ---------------------------------
    def recall_for(ef):
        out = step_device(nq, ef)  # full batch keeps the gather shapes fixed
        return recall_at_k(torch, out[:n_gt], gt)

    curve = {}
    if a.ef > 0:
        ef = a.ef
        curve[ef] = recall_for(ef)
    else:
        ef = EF_LADDER[-1]
        for e in EF_LADDER:
            curve[e] = recall_for(e)
            if curve[e] >= RECALL_TARGET:
                ef = e
                break
    recall = curve[ef]

    # =========================================================================================
    if a.impl == "reference":
        from oracle import pyoracle as orc

        threads = os.cpu_count() or 1
        g = index.graph
        xh = shard.cpu().numpy()
        qh = q.cpu().numpy()
        off, nbrs, entry = g.node_offsets, g.neighbors, g.entry_point
        # size one step: ~ (cpu budget / steps) seconds of all-core work
        qps0, m0, _ = cpu_port_qps(orc, cfg, xh, off, nbrs, entry, qh, ef, threads, 2.0)
        per_step = int(max(threads, min(nq, qps0 * max(1.0, 60.0 / max(1, a.steps + a.warmup)))))
        for _ in range(a.warmup):
            orc.leann_search(cfg._s, xh, off, nbrs, entry, qh[:per_step], K_TOP, ef, threads=threads)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            r_ids, _, _ = orc.leann_search(cfg._s, xh, off, nbrs, entry, qh[:per_step], K_TOP, ef, threads=threads)
        dt = time.perf_counter() - t0
        qps = per_step * a.steps / dt
        m = min(per_step, n_gt)
        rec = recall_at_k(torch, torch.from_numpy(r_ids[:m].astype(np.int64)).to(dev), gt[:m])
        sample = f"{per_step} of {nq} queries per step, ef={ef}, graph built by the GPU library in setup (untimed)"
        print(json.dumps({
            "impl": "reference", "metric": "QPS at recall@10>=0.95 (1M x 768, M=30, top-10)", "value": qps, "unit": "queries/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n} x {d} f32 {a.dataset}, LEANN graph m=30 m0=60 efC=128, nq={nq} (sampled {per_step}), top-10, exact traversal, cosine",
                       "ef": ef, "recall_at_10": rec, "kind": "oracle port of src/core/leann.rs:868-988 (Rust reference cannot be built here)"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return

    # ---- counters for the roofline (one untimed pass with per-query stats) ---------------------------
    step_device(nq, ef, with_stats=True)
    torch.cuda.synchronize()
    alg_bytes, per_query = algorithmic_bytes(stats.cpu().numpy(), d, nq, K_TOP)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: K timed steps, inputs resident in HBM ------------------------------------------------
    for _ in range(a.warmup):
        step_device(nq, ef)
    kernel_ms = []
    barrier()
    lib.isl_kernel_launch_count_reset()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_device(nq, ef)
            kernel_ms.append(index.last_search_timing()[0])
        barrier()
        dt = time.perf_counter() - t0
    launches = int(lib.isl_kernel_launch_count())
    if use_dist:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    qps = nq * a.steps / dt

    # ---- e2e: host buffers in, host results out, copies inside the timed region ------------------------
    qh = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
    qh.copy_(q)
    out_ids_h = torch.empty((nq, K_TOP), dtype=torch.int64, pin_memory=True)
    out_dst_h = torch.empty((nq, K_TOP), dtype=torch.float32, pin_memory=True)
    qn = qh.numpy()

    def step_e2e():
        if not use_dist:
            return index.search_batch(qn, K_TOP, ef)  # isl_index_search: H2D queries, kernel, D2H results
        q.copy_(qh, non_blocking=True)
        out = step_device(nq, ef)
        out_ids_h.copy_(out, non_blocking=True)
        out_dst_h.copy_(m_dst, non_blocking=True)
        torch.cuda.synchronize()
        return None

    for _ in range(max(1, a.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    dt_e2e = time.perf_counter() - t0
    if use_dist:
        t = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_e2e = float(t.item())
    e2e_qps = nq * a.steps / dt_e2e

    peak, peak_src = measured_peak()
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    line = {
        "metric": "QPS at recall@10>=0.95 (1M x 768, M=30, top-10)", "value": qps, "unit": "queries/s", "n_gpus": world if use_dist else 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{n} x {d} f32 {a.dataset}, LEANN graph m=30 m0=60 efC=128 hub 2% (built on GPU in setup, {build_s:.1f}s), "
                        f"batched {nq} queries, top-10, exact traversal, cosine",
            "ef": ef, "recall_at_10": recall, "recall_curve": {str(k): round(v, 4) for k, v in curve.items()},
            "shards": parts, "shard_nodes": hi - lo, "merge": "NCCL all-gather + per-query (dist,id) merge" if use_dist else "none",
            "l2": "inputs larger than L2 (vector table %.2f GB vs 126 MB)" % ((hi - lo) * d * 4 / 1e9),
            "per_query": per_query,
        },
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * K_TOP * 12 + (0 if use_dist else nq * 4)},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "leann_search_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "peak_source": peak_src},
    }

    # ---- cpu_baseline (rank 0, single GPU run only) -------------------------------------------------
    if not use_dist:
        from oracle import pyoracle as orc

        threads = os.cpu_count() or 1
        g = index.graph
        xh = shard.cpu().numpy()
        cq, m, o_ids = cpu_port_qps(orc, cfg, xh, g.node_offsets, g.neighbors, g.entry_point, qn, ef, threads, a.cpu_seconds)
        same = bool(np.array_equal(o_ids.astype(np.int64), index.search_batch(qn[:m], K_TOP, ef)[0].astype(np.int64)))
        line["cpu_baseline"] = {"value": cq, "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": f"first {m} of {nq} queries, same graph / ef, all host threads; ids equal to GPU: {same}"}
        del xh

        # secondary: the reference benches' own distribution (uniform), reported beside the headline
        if a.dataset != "uniform" and not a.no_uniform:
            del index, shard
            torch.cuda.empty_cache()
            xu, qu = make_data(torch, "uniform", n, nq, d, dev)
            gtu = ground_truth(torch, xu, qu[:n_gt], K_TOP)
            iu = LeannIndex(cfg)
            iu.build_dev(xu.data_ptr(), n, d, seed=7, batch=a.build_batch)
            uni = {}
            for e in (64, 1024):
                iu.search_batch_dev(qu.data_ptr(), nq, d, K_TOP, e, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), stats.data_ptr())
                iu.search_batch_dev(qu.data_ptr(), nq, d, K_TOP, e, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), stats.data_ptr())
                ms = iu.last_search_timing()[0]
                b, pq_ = algorithmic_bytes(stats.cpu().numpy(), d, nq, K_TOP)
                uni[str(e)] = {"qps": nq / ms * 1e3, "recall_at_10": recall_at_k(torch, ids[:n_gt], gtu), "n_dist": pq_["n_dist"],
                               "gbps": b / ms / 1e6, "frac": b / ms / 1e6 / peak}
            line["uniform_reference_distribution"] = {
                "note": "U[-1,1)^768 (benches/hnsw_benchmarks.rs:9-14): distances concentrate, recall>=0.95 needs a near-exhaustive traversal for any graph index",
                "by_ef": uni}
    if rank == 0:
        print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
