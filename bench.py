#!/usr/bin/env python
"""bench.py — LEANN search hot path on B200: QPS at recall@10 >= 0.95 (1M x 768, M=30, top-10).

One step = one batch of `nq` queries through the batched best-first search of the resident
index (islands_b200 C ABI).  See DESIGN.md "measurement" for every definition used here.

  python bench.py --gpus 1 --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference ...                     # the reference's CPU algorithm (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...        # one island (n x d shard) per GPU, weak scaling

N > 1 (DESIGN.md "multi-GPU"): every rank owns one island / node-range shard of `n` vectors (its own
seed) — the total index is N*n vectors — and `value` counts queries routed to their island (the
reference's `index_names` filter, src/indexer/service.rs:768-771): independent units, no data-path
collective, weak scaling.  The same run also measures the sharded search (every query on every shard,
ONE exchange of packed (dist,id) records + per-query merge, service.rs:777-801) through the library's
own entry point isl_index_search_sharded_dev — with ncclAllGather and with the peer-store exchange —
and reports it under "sharded" with a search / exchange / merge breakdown.

The run exits non-zero when a parity check made along the way fails (GPU ids vs the CPU port).
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
EF_REFINE_STEP = 8  # after the coarse ladder, the last interval is refined in steps of 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="islands_b200", choices=["islands_b200", "reference"])
    # (the long spellings exist because torchrun's own parser claims "--n" / "--d" prefixes when it launches this script)
    ap.add_argument("--n", "--island-nodes", dest="n", type=int, default=1_000_000)
    ap.add_argument("--d", "--dim", dest="d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--dataset", type=str, default="latent32", choices=["latent32", "uniform"])
    ap.add_argument("--ef", type=int, default=0, help="0 = smallest ef of the ladder with recall@10 >= 0.95")
    ap.add_argument("--build-batch", type=int, default=4096)
    ap.add_argument("--no-uniform", action="store_true", help="skip the secondary uniform-data measurement")
    ap.add_argument("--no-hnsw", action="store_true", help="skip the secondary HnswGraph (configs[0]) measurement")
    ap.add_argument("--no-encoder", action="store_true", help="skip the secondary recompute-encoder measurement")
    ap.add_argument("--no-adc", action="store_true", help="skip the secondary PQ ADC traversal + exact rerank measurement")
    ap.add_argument("--pq-m", type=int, default=32, help="subquantizers of the ADC secondary")
    ap.add_argument("--pq-ksub", type=int, default=128, help="centroids per subquantizer of the ADC secondary "
                    "(128: an 8 KB bfloat16 table per query in shared memory keeps twice the warps resident of 256)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--build-sample", type=int, default=10000, help="nodes of the construction cpu_baseline sample (oracle build + GPU build of the same prefix, graphs compared)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)  # the timing rules ask for at least three untimed steps; the JSON line reports what was done
    a.steps = max(a.steps, 1)
    return a


# ---------------------------------------------------------------------------------------------
# synthetic data (generated on the GPU; seeds fixed)
# ---------------------------------------------------------------------------------------------
def make_data(torch, dataset, n, nq, d, dev, seed=42, qseed=43):
    """latent32: x = z A + 0.05 eps, z ~ N(0, I_32), A ~ N(0, 1/32)^{32 x d}: d-dimensional vectors
    with the low intrinsic dimension typical of learned embeddings (graph ANN reaches 0.95 recall).
    uniform: i.i.d. U[-1,1), the distribution of the reference's criterion benches
    (benches/hnsw_benchmarks.rs:9-14) — a worst case for ANY graph index at d=768."""
    g = torch.Generator(device=dev)
    g.manual_seed(42)  # the latent basis is shared by every island and every query
    if dataset == "latent32":
        a = torch.randn((32, d), generator=g, device=dev) / 32 ** 0.5
        g.manual_seed(seed + 7)
        x = torch.randn((n, 32), generator=g, device=dev) @ a + 0.05 * torch.randn((n, d), generator=g, device=dev)
        g.manual_seed(qseed)
        q = torch.randn((nq, 32), generator=g, device=dev) @ a + 0.05 * torch.randn((nq, d), generator=g, device=dev)
    else:
        g.manual_seed(seed)
        x = torch.rand((n, d), generator=g, device=dev) * 2 - 1
        g.manual_seed(qseed)
        q = torch.rand((nq, d), generator=g, device=dev) * 2 - 1
    return x.contiguous(), q.contiguous()


def ground_truth_scores(torch, x, q, k, rows=1 << 20):
    """Exact cosine top-k by brute force (measurement infrastructure, not the hot path): similarities and ids.
    The base vectors are streamed in blocks of `rows`, so a 12.5M-node shard needs no second copy of itself."""
    qn = torch.nn.functional.normalize(q, dim=1)
    best_v = torch.full((q.shape[0], k), -float("inf"), device=q.device)
    best_i = torch.zeros((q.shape[0], k), dtype=torch.int64, device=q.device)
    for s in range(0, x.shape[0], rows):
        xb = torch.nn.functional.normalize(x[s:s + rows], dim=1)
        t = (qn @ xb.T).topk(min(k, xb.shape[0]), dim=1)
        v = torch.cat([best_v, t.values], dim=1)
        i = torch.cat([best_i, t.indices + s], dim=1)
        sel = v.topk(k, dim=1)
        best_v, best_i = sel.values, i.gather(1, sel.indices)
        del xb, t
    return best_v, best_i


def ground_truth(torch, x, q, k):
    return ground_truth_scores(torch, x, q, k)[1]


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
        self.source = "nvidia-smi"
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process (a sample costs well under a millisecond, so even a sub-second timed region gets dozens);
        # falls back to the nvidia-smi command line of the profiling recipe when the binding is missing.
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.source = "nvml (the counters nvidia-smi prints)"
            while not self._stop.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for nme, b in bits.items():
                    if r & b:
                        self.reasons.add(nme)
                self._stop.wait(0.01)
            return
        except Exception:
            pass
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
            self._stop.wait(0.05)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


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


def workload_config(a, n, d, nq, parts, ef):
    """`config` of the JSON line — the same dict in both arms (ours and --impl reference)."""
    return {
        "workload": f"{n} x {d} f32 {a.dataset} per GPU, LEANN graph m=30 m0=60 efC=128 hub 2% (built on the GPU in setup), "
                    f"batched {nq} queries per GPU per step, top-10, exact traversal (leann.rs:868-988), cosine",
        "ef": ef, "islands": parts, "island_nodes": n, "total_nodes": parts * n,
        "l2": "inputs larger than L2 (vector table %.2f GB vs 126 MB)" % (n * d * 4 / 1e9),
    }


def calibrate_ef(recall_for, target, fixed=0):
    """Smallest ef with recall >= target: coarse ladder, then the last interval in steps of 8."""
    curve = {}
    if fixed > 0:
        curve[fixed] = recall_for(fixed)
        return fixed, curve
    ef, prev = EF_LADDER[-1], 0
    for e in EF_LADDER:
        curve[e] = recall_for(e)
        if curve[e] >= target:
            ef = e
            break
        prev = e
    if curve[ef] >= target and ef - prev > EF_REFINE_STEP:
        for e in range(prev + EF_REFINE_STEP, ef, EF_REFINE_STEP):
            curve[e] = recall_for(e)
            if curve[e] >= target:
                ef = e
                break
    return ef, dict(sorted(curve.items()))


def main():
    a = parse_args()
    # stdout carries exactly one line, the JSON result: everything else a library may print there (NCCL announces its
    # version on stdout when the first communicator is created) goes to stderr
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
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

    from islands_b200 import LeannConfig, LeannIndex, PQConfig, ProductQuantizer, _ffi
    from islands_b200.shard import ShardedLeannIndex

    lib = _ffi.load()
    failures = []  # parity checks made along the way; any entry makes the run exit non-zero
    cfg = LeannConfig()  # paper_default: m=30, m0=60, efC=128, cosine, hub-preserving pruning 2%
    n, d, nq = a.n, a.d, a.nq
    island = rank if use_dist else 0  # one island of n vectors per GPU (weak scaling)

    x, q = make_data(torch, a.dataset, n, nq, d, dev, seed=42 + 1000 * island, qseed=43 + 1000 * island)
    n_gt = min(nq, 1000)
    gt = ground_truth(torch, x, q[:n_gt], K_TOP)

    t0 = time.perf_counter()
    index = LeannIndex(cfg)
    index.build_dev(x.data_ptr(), n, d, seed=7, batch=a.build_batch)
    build_s = time.perf_counter() - t0

    ids = torch.empty((nq, K_TOP), dtype=torch.int64, device=dev)
    dst = torch.empty((nq, K_TOP), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    stats = torch.zeros((nq, 5), dtype=torch.int64, device=dev)

    def step_device(ef, with_stats=False, queries=None):
        """One pass of the hot path with inputs resident in HBM: this island's batch through the
        batched best-first search (isl_index_search_dev).  Returns the ids tensor."""
        qq = q if queries is None else queries
        index.search_batch_dev(qq.data_ptr(), nq, d, K_TOP, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(),
                               stats.data_ptr() if with_stats else None)
        return ids

    def all_max(v):
        if not use_dist:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- ef: smallest value with recall@10 >= 0.95 on every island (setup, untimed) -----------------
    ef, curve = calibrate_ef(lambda e: recall_at_k(torch, step_device(e)[:n_gt], gt), RECALL_TARGET, a.ef)
    ef = int(all_max(ef))
    recall = curve[ef] if ef in curve else recall_at_k(torch, step_device(ef)[:n_gt], gt)

    metric_name = "QPS at recall@10>=0.95 (1M x 768, M=30, top-10)"
    # =========================================================================================
    if a.impl == "reference":
        from oracle import pyoracle as orc

        threads = os.cpu_count() or 1
        g = index.graph
        xh = x.cpu().numpy()
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
        print(file=result_out, flush=True, *[json.dumps({
            "impl": "reference", "metric": metric_name, "value": qps, "unit": "queries/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, n, d, nq, a.gpus, ef),
            "details": {"kind": "oracle port of src/core/leann.rs:868-988 (the Rust reference cannot be built here: no cargo / rustc)",
                        "queries_per_step": per_step, "recall_at_10": rec,
                        "n_gpus_note": "queries are routed to their island, so the CPU's cost per query does not depend on the number of islands: "
                                       "rank 0 searches its own island's graph with all host threads"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        })])
        return

    # ---- counters for the roofline (one untimed pass with per-query stats) ---------------------------
    step_device(ef, with_stats=True)
    torch.cuda.synchronize()
    alg_bytes, per_query = algorithmic_bytes(stats.cpu().numpy(), d, nq, K_TOP)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        return all_max(time.perf_counter() - t0)

    # ---- value: K timed steps, inputs resident in HBM ------------------------------------------------
    for _ in range(a.warmup):
        step_device(ef)
    kernel_ms = []
    barrier()
    lib.isl_kernel_launch_count_reset()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_device(ef)
            kernel_ms.append(index.last_search_timing()[0])
        barrier()
        dt = time.perf_counter() - t0
    launches = int(lib.isl_kernel_launch_count())
    dt = all_max(dt)
    qps = world * nq * a.steps / dt if use_dist else nq * a.steps / dt

    # ---- e2e: host buffers in, host results out, copies inside the timed region ------------------------
    qh = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
    qh.copy_(q)
    qn = qh.numpy()
    dt_e2e = timed(lambda: index.search_batch(qn, K_TOP, ef), a.steps, max(1, a.warmup))  # isl_index_search
    e2e_qps = (world if use_dist else 1) * nq * a.steps / dt_e2e
    lat = []  # batch latency of the synchronous host-buffer call (p50 / p99), outside the timed regions
    for _ in range(max(a.steps, 10)):
        t0 = time.perf_counter()
        index.search_batch(qn, K_TOP, ef)
        lat.append((time.perf_counter() - t0) * 1e3)

    peak, peak_src = measured_peak()
    traffic = None  # DRAM bytes of this launch from the committed ncu --set full capture, when it is the same workload
    try:
        with open(os.path.join(ROOT, "profiles", "r02_search_traffic.json")) as f:
            tj = json.load(f)
        if tj["workload"] == {"n": n, "d": d, "nq": nq, "dataset": a.dataset, "ef": ef, "k": K_TOP}:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    parts = world if use_dist else 1
    line = {
        "metric": metric_name, "value": qps, "unit": "queries/s", "n_gpus": parts,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, n, d, nq, parts, ef),
        "details": {"recall_at_10": recall, "recall_curve": {str(k): round(v, 4) for k, v in curve.items()}, "build_s": build_s,
                    "routing": "queries routed to their island (index_names filter, service.rs:768-771); no data-path collective" if use_dist else "single island",
                    "per_query": per_query},
        "batch_latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "calls": len(lat),
                             "what": f"isl_index_search, host buffers, {nq} queries per call (this rank)"},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": parts * nq * d * 4,
                "d2h_bytes_per_step": parts * (nq * K_TOP * 12 + nq * 4)},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "leann_search_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
    }

    # ---- N > 1: the sharded search (every query on every shard + ONE exchange + merge), library entry point ---------
    if use_dist:
        from islands_b200.shard import make_shard_comm

        _, q_all = make_data(torch, a.dataset, 1, nq, d, dev, seed=1, qseed=43)  # the same batch on every rank
        sc, li = ground_truth_scores(torch, x, q_all[:n_gt], K_TOP)
        g_sc = torch.empty((world * n_gt, K_TOP), dtype=sc.dtype, device=dev)
        g_id = torch.empty((world * n_gt, K_TOP), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(g_sc, sc.contiguous())
        dist.all_gather_into_tensor(g_id, (li + rank * n).contiguous())
        g_sc = g_sc.view(world, n_gt, K_TOP).permute(1, 0, 2).reshape(n_gt, -1)
        g_id = g_id.view(world, n_gt, K_TOP).permute(1, 0, 2).reshape(n_gt, -1)
        gt_all = g_id.gather(1, g_sc.topk(K_TOP, dim=1).indices)
        comm = make_shard_comm(device=dev)  # NCCL communicator inside the library; torch only carries the unique id
        sharded = ShardedLeannIndex(index, rank * n, world * n, comm)
        m_ids = torch.empty((nq, K_TOP), dtype=torch.int64, device=dev)
        m_dst = torch.empty((nq, K_TOP), dtype=torch.float32, device=dev)

        def step_all(e):
            return sharded.search_batch_dev(q_all, K_TOP, e, m_ids, m_dst, cnt)[0]

        ef_all, curve_all = calibrate_ef(lambda e: recall_at_k(torch, step_all(e)[:n_gt], gt_all), RECALL_TARGET)
        ef_all = int(all_max(ef_all))

        def measure_engine():
            for _ in range(a.warmup):
                step_all(ef_all)
            parts_ms = []
            barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                step_all(ef_all)
                parts_ms.append(comm.last_timing())
            barrier()
            dt_ = all_max(time.perf_counter() - t0)
            pm = np.asarray(parts_ms, np.float64).mean(axis=0)
            # max over ranks of every stage (the slowest rank sets the step; a fast rank's "exchange" is mostly waiting)
            return dt_, [all_max(pm[0]), all_max(pm[1]), all_max(pm[2])], [float(v) for v in pm]

        dt_nccl, st_nccl, mine_nccl = measure_engine()
        ids_nccl = m_ids.clone()
        comm.enable_peer_exchange(nq * K_TOP)
        dt_peer, st_peer, mine_peer = measure_engine()
        same_engines = bool(torch.equal(ids_nccl, m_ids))
        rec_all = recall_at_k(torch, m_ids[:n_gt], gt_all)
        best = min(dt_nccl, dt_peer)
        line["sharded"] = {
            "what": "every query searches all shards (service.rs:777-801): isl_index_search_sharded_dev = search kernel -> ONE exchange of packed "
                    "16-byte (dist, global id) records -> merge kernel, one stream, no host synchronisation in between",
            "total_nodes": world * n, "shards": world, "ef": ef_all, "recall_at_10": rec_all,
            "recall_curve": {str(k): round(v, 4) for k, v in curve_all.items()},
            "qps": nq * a.steps / best, "ms_per_step": best / a.steps * 1e3,
            "nccl_allgather": {"qps": nq * a.steps / dt_nccl, "ms_per_step": dt_nccl / a.steps * 1e3,
                               "stage_ms_max_over_ranks": {"search": st_nccl[0], "exchange": st_nccl[1], "merge": st_nccl[2]},
                               "stage_ms_rank0": {"search": mine_nccl[0], "exchange": mine_nccl[1], "merge": mine_nccl[2]}},
            "peer_stores": {"qps": nq * a.steps / dt_peer, "ms_per_step": dt_peer / a.steps * 1e3,
                            "stage_ms_max_over_ranks": {"search": st_peer[0], "exchange": st_peer[1], "merge": st_peer[2]},
                            "stage_ms_rank0": {"search": mine_peer[0], "exchange": mine_peer[1], "merge": mine_peer[2]},
                            "how": "the search kernel stores each finished query's records into every rank's gather buffer over NVLink (CUDA IPC "
                                   "mappings); a flag handshake replaces the collective"},
            "engines_agree": same_engines,
            "exchange_bytes_per_rank_per_step": nq * K_TOP * 16,
            "limiter": "the search kernel" if st_nccl[0] > 5 * (st_nccl[1] + st_nccl[2]) else "see stage_ms",
            "note": "queries are replicated, so QPS does not grow with N; the index does (capacity scaling); exchange_ms on a rank includes waiting for the slowest rank"}
        if not same_engines:
            failures.append("sharded search: NCCL and peer-store exchange returned different ids")
        comm.free()

    # ---- cpu_baseline + secondary workloads (single GPU run only) ------------------------------------
    if not use_dist:
        from oracle import pyoracle as orc

        threads = os.cpu_count() or 1
        g = index.graph
        xh = x.cpu().numpy()
        cq, m, o_ids = cpu_port_qps(orc, cfg, xh, g.node_offsets, g.neighbors, g.entry_point, qn, ef, threads, a.cpu_seconds)
        same = bool(np.array_equal(o_ids.astype(np.int64), index.search_batch(qn[:m], K_TOP, ef)[0].astype(np.int64)))
        # the reference's own execution model is one thread (search_batch is a sequential map, search.rs:179-181)
        cq1, m1, _ = cpu_port_qps(orc, cfg, xh, g.node_offsets, g.neighbors, g.entry_point, qn, ef, 1, min(a.cpu_seconds, 5.0))
        if not same:
            failures.append(f"headline workload: GPU ids differ from the CPU port on the first {m} queries")
        line["cpu_baseline"] = {"value": cq, "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": f"first {m} of {nq} queries, same graph / ef, all host threads; ids equal to GPU: {same}",
                                "single_thread": {"value": cq1, "unit": "queries/s", "sample": f"first {m1} queries, one thread "
                                                  "(the reference's execution model, search.rs:179-181)"}}

        # BASELINE configs[3]: graph construction 1M x 768, efConstruction=128, hub-preserving pruning, on the GPU
        bs = index.last_build_stats()
        b_build = bs["n_dist"] * 4 * d + bs["n_edge"] * 4 + bs["n_hop"] * 16 + n * 4 * d + 2 * bs["edges"] * 4
        ns = min(n, a.build_sample)
        xs = xh[:ns]
        lv = orc.draw_levels(7, ns, cfg.ml, cfg.max_layers)
        t0 = time.perf_counter()
        o_off, o_nb, o_entry, _ = orc.leann_build(cfg._s, xs, lv, batch=a.build_batch, threads=threads)
        cpu_build_s = time.perf_counter() - t0
        gs = LeannIndex(cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gs.build(xs, ns, levels=lv, batch=a.build_batch)
        gpu_sample_s = time.perf_counter() - t0
        gg = gs.graph
        same_graph = bool(np.array_equal(gg.node_offsets, o_off) and np.array_equal(gg.neighbors, o_nb) and gg.entry_point == o_entry)
        if not same_graph:
            failures.append(f"construction: the GPU graph of the first {ns} nodes differs from the CPU port's")
        gs.free()
        line["build"] = {
            "what": f"BASELINE configs[3]: LeannIndex::build {n} x {d}, efConstruction=128, hub-preserving pruning 2%, rounds of {a.build_batch} inserts (leann.rs:560-833)",
            "seconds": build_s, "inserts_per_s": n / build_s, "gpu_rounds_ms": bs["rounds_ms"], "search_kernel_ms": bs["search_ms"],
            "search_share_of_rounds": bs["search_ms"] / max(bs["rounds_ms"], 1e-9), "rounds": bs["rounds"], "edges": bs["edges"],
            "per_insert": {"n_dist": bs["n_dist"] / n, "n_edge": bs["n_edge"] / n, "n_hop": bs["n_hop"] / n},
            "roofline": {"bound": "hbm", "kernel": "leann_search_kernel (efConstruction searches of every round)", "algorithmic_bytes": b_build,
                         "achieved": (b_build - 2 * bs["edges"] * 4) / max(bs["search_ms"], 1e-9) / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": (b_build - 2 * bs["edges"] * 4) / max(bs["search_ms"], 1e-9) / 1e6 / peak,
                         "whole_build_gbps": b_build / max(bs["rounds_ms"], 1e-9) / 1e6, "peak_source": peak_src},
            "cpu_baseline": {"value": ns / cpu_build_s, "unit": "inserts/s", "cores": threads, "kind": "port",
                             "sample": f"oracle port of leann.rs:560-833 (same round model) on the first {ns} vectors, all host threads; "
                                       f"GPU on the same sample: {ns / gpu_sample_s:.0f} inserts/s (host buffers), graphs bit-identical: {same_graph}"}}

        # secondary: BASELINE configs[1] names "PQ ADC traversal + exact rerank" — a mode the reference
        # specifies (docs/leann-specification.md:223-269) but does not implement; measured beside the headline.
        if not a.no_adc:
            pq_m = a.pq_m
            pq_ksub = a.pq_ksub
            pq = ProductQuantizer(d, PQConfig(pq_m, pq_ksub, 8, 1))
            pq.train(xh[:20000])
            pq_codes = pq.encode(xh)
            index.attach_pq(pq, pq_codes)

            def adc_recall(e):
                r = index.search_adc_rerank_batch(qn, K_TOP, e)[0]
                return recall_at_k(torch, torch.from_numpy(r[:n_gt].astype(np.int64)).to(dev), gt)

            ef_adc, curve_adc = calibrate_ef(adc_recall, RECALL_TARGET)
            ids_bitset, dist_bitset, _, st = index.search_adc_rerank_batch(qn, K_TOP, ef_adc, stats=True)
            for _ in range(max(3, a.warmup)):  # timed without statistics: the traversal then runs without the visited bitset
                ids_free, dist_free, _ = index.search_adc_rerank_batch(qn, K_TOP, ef_adc)
            ms_l = []
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                index.search_adc_rerank_batch(qn, K_TOP, ef_adc)  # isl_index_search_adc_rerank: host buffers in, host results out
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / a.steps
            for _ in range(3):  # kernel time (CUDA events inside the library), read outside the wall-clock loop
                index.search_adc_rerank_batch(qn, K_TOP, ef_adc)
                ms_l.append(index.last_search_timing()[0])
            same_adc = bool(np.array_equal(ids_bitset, ids_free) and np.array_equal(dist_bitset.view(np.uint32), dist_free.view(np.uint32)))
            if not same_adc:
                failures.append("ADC traversal: bitset-free results differ from the visited-bitset results")
            # the oracle's twin of this mode (orc_leann_search_adc_rerank: same bfloat16 table rule, same loop) on a bounded
            # prefix of the same queries, same graph, codebooks, codes and ef: ids and distance bits must be equal
            m_adc = min(nq, 2000)
            t0 = time.perf_counter()
            o_ids_adc, o_dist_adc, _ = orc.leann_search_adc_rerank(cfg._s, xh, g.node_offsets, g.neighbors, g.entry_point, pq.codebooks(),
                                                                   pq_codes, qn[:m_adc], K_TOP, ef_adc, threads=threads)
            cpu_adc_qps = m_adc / (time.perf_counter() - t0)
            same_adc_cpu = bool(np.array_equal(o_ids_adc.astype(np.int64), ids_free[:m_adc].astype(np.int64))
                                and np.array_equal(o_dist_adc.view(np.uint32), dist_free[:m_adc].view(np.uint32)))
            if not same_adc_cpu:
                failures.append(f"ADC traversal + rerank: GPU results differ from the CPU port on the first {m_adc} queries")
            ms = float(np.mean(ms_l))
            b = int(st.n_adc.sum()) * pq_m + int(st.n_edge.sum()) * 4 + int(st.n_hop.sum()) * 16 + int(st.n_rerank.sum()) * 4 * d \
                + nq * (4 * d + 12 * K_TOP + pq_m * pq_ksub * 4)
            line["adc_rerank"] = {"what": "BASELINE configs[1]: 1M x 768 CSR graph, batched 10k queries, PQ ADC traversal + exact rerank, 1 B200",
                                  "pq_m": pq_m, "pq_ksub": pq_ksub, "ef": ef_adc, "recall_at_10": curve_adc[ef_adc], "kernel_qps": nq / ms * 1e3,
                                  "kernel_ms": ms, "e2e_qps_host_buffers": nq / wall, "e2e_steps": a.steps,
                                  "algorithmic_gbps": b / ms / 1e6, "frac": b / ms / 1e6 / peak,
                                  "n_adc": float(st.n_adc.mean()), "n_rerank": float(st.n_rerank.mean()),
                                  # every traversal access is one 32-byte sector (a code row, an 8-sector list per hop): the rate against the
                                  # measured ceiling of independent random sectors out of L2 (profiles/r01_sector_ceiling.txt: 210-220 G/s)
                                  "traversal_sectors_per_s": (int(st.n_adc.sum()) * (pq_m // 32 if pq_m >= 32 else 1) + int(st.n_hop.sum()) * 8) / (ms * 1e-3),
                                  "frac_of_l2_sector_ceiling": (int(st.n_adc.sum()) * (pq_m // 32 if pq_m >= 32 else 1) + int(st.n_hop.sum()) * 8) / (ms * 1e-3) / 215e9,
                                  "bitset_free_results_equal_bitset_results": same_adc,
                                  "cpu_baseline": {"value": cpu_adc_qps, "unit": "queries/s", "cores": threads, "kind": "port",
                                                   "sample": f"orc_leann_search_adc_rerank on the first {m_adc} queries, same graph / codebooks / codes / ef, "
                                                             f"all host threads; ids and distance bits equal to GPU: {same_adc_cpu}"},
                                  "bound": "issue slots of the per-hop chain (profiles/r02_adc_bag_ncu.txt: 60 % busy at 21 resident queries per SM); the byte "
                                           "roofline is not the limiter: every traversal access is one 32-byte sector out of L2, only the exact rerank streams from HBM",
                                  "note": "parity unpinned: no reference implementation of this mode exists (leann.rs:54-56); oracle <-> GPU bit-exact"}
            line["roofline"]["also"] = {"adc_rerank_kernel_qps": nq / ms * 1e3, "adc_rerank_recall_at_10": curve_adc[ef_adc], "adc_rerank_frac_of_hbm_bytes": b / ms / 1e6 / peak}
        del xh

        # secondary: the recompute encoder (BASELINE configs[4]: random-init 110M bf16 encoder, the only
        # tensor-core row) — one batch of frontier nodes through the tcgen05 encoder.
        if not a.no_encoder:
            from islands_b200 import Encoder, EncoderConfig

            enc = Encoder(EncoderConfig()).init_random(seed=46, stddev=0.02)
            eb, es = 2048, 64
            gtok = torch.Generator(device=dev)
            gtok.manual_seed(45)
            tok = torch.randint(1, 30000, (eb, es), generator=gtok, device=dev, dtype=torch.int32)
            eln = torch.full((eb,), es, device=dev, dtype=torch.int32)
            eout = torch.empty((eb, 768), device=dev)
            torch.cuda.synchronize()
            ems = []
            for it in range(5):
                enc.embed_dev(tok.data_ptr(), eln.data_ptr(), eb, es, eout.data_ptr())
                if it >= 2:
                    ems.append(enc.last_timing())
            ms_e = float(np.mean([m for m, _ in ems]))
            fl = ems[0][1]
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    tpeak = float(json.load(f)["bf16_tflops_sustained"])
            except Exception:
                tpeak = 1400.0
            line["recompute_encoder"] = {"shape": "BERT-base 110M random init, bf16 tcgen05 GEMMs, f32 accumulate", "batch": eb, "seq_len": es,
                                         "ms": ms_e, "sequences_per_s": eb / ms_e * 1e3, "tflops": fl / ms_e / 1e9,
                                         "roofline": {"bound": "tensor", "achieved": fl / ms_e / 1e9, "peak": tpeak, "unit": "TFLOP/s",
                                                      "frac": fl / ms_e / 1e9 / tpeak, "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"}}
            del enc, tok, eout
            torch.cuda.empty_cache()

        # secondary: BASELINE configs[0] — HnswGraph (hnsw.rs) on 100k random 768-d vectors, M=30, efSearch=64, top-10
        if not a.no_hnsw:
            from islands_b200 import HnswConfig, HnswGraph

            hn = 100_000
            xh_, qh_ = make_data(torch, "uniform", hn, nq, d, dev)
            gth = ground_truth(torch, xh_, qh_[:n_gt], K_TOP)
            hg = HnswGraph(HnswConfig(m=30, m0=60, ef_construction=128, ml=1.0 / np.log(30.0)))
            t0 = time.perf_counter()
            hg.insert_batch_dev(xh_.data_ptr(), hn, d, seed=7, batch=1024)
            hbuild = time.perf_counter() - t0
            hms = []
            for it in range(4):
                hg.search_batch_dev(qh_.data_ptr(), nq, d, K_TOP, 64, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr())
                if it:
                    hms.append(hg.last_search_timing())
            line["hnsw_config0"] = {"workload": "HnswGraph 100k x 768 uniform, m=30 m0=60 efC=128, 10000 queries, ef=64, top-10 (greedy descent + layer-0 search)",
                                    "insert_s": hbuild, "inserts_per_s": hn / hbuild, "max_level": hg.max_level,
                                    "search_ms": float(np.mean(hms)), "qps": nq / float(np.mean(hms)) * 1e3,
                                    "recall_at_10": recall_at_k(torch, ids[:n_gt], gth),
                                    "note": "bit-exact with the reference's HnswGraph, including its prune_connections quirk (hnsw.rs:419-420 with :327: "
                                            "the id being inserted is filtered out of a full neighbour list), so only the first ~m0 nodes ever "
                                            "receive incoming edges and the traversal stays among them: recall and speed are the reference's"}
            # BASELINE configs[0] for the LEANN index itself: 100k x 768, ef = 64, top-10, GPU and the CPU port on the
            # same graph and queries (1000 of them on the CPU), ids compared
            lg = LeannIndex(cfg)
            t0 = time.perf_counter()
            lg.build_dev(xh_.data_ptr(), hn, d, seed=7, batch=1024)
            torch.cuda.synchronize()
            lbuild = time.perf_counter() - t0
            lms = []
            for it in range(4):
                lg.search_batch_dev(qh_.data_ptr(), nq, d, K_TOP, 64, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), 0)
                if it:
                    lms.append(lg.last_search_timing()[0])
            lrec = recall_at_k(torch, ids[:n_gt], gth)
            gids = ids[:1000].cpu().numpy().astype(np.int64)
            gg = lg.graph
            xc, qc = xh_.cpu().numpy(), qh_[:1000].cpu().numpy()
            t0 = time.perf_counter()
            oids, _, _ = orc.leann_search(cfg._s, xc, gg.node_offsets, gg.neighbors, gg.entry_point, qc, K_TOP, 64, threads=threads)
            ldt = time.perf_counter() - t0
            if not bool(np.array_equal(oids.astype(np.int64), gids)):
                failures.append("leann_config0 (100k x 768 uniform): GPU ids differ from the CPU port")
            line["leann_config0"] = {"workload": "LeannIndex 100k x 768 uniform, m=30 m0=60 efC=128, ef=64, top-10 (BASELINE configs[0])",
                                     "build_s": lbuild, "search_ms_per_10k": float(np.mean(lms)), "qps": nq / float(np.mean(lms)) * 1e3,
                                     "recall_at_10": lrec, "cpu_port_qps": 1000 / ldt, "cpu_threads": threads,
                                     "ids_equal_cpu_port": bool(np.array_equal(oids.astype(np.int64), gids)),
                                     "note": "uniform 768-d data: recall at ef = 64 is what any graph index gives there (SURVEY F10)"}
            del hg, lg, xh_, qh_, xc, qc
            torch.cuda.empty_cache()

        # secondary: the reference benches' own distribution (uniform), reported beside the headline
        if a.dataset != "uniform" and not a.no_uniform:
            del index, x
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
    if failures:
        line["parity_failures"] = failures
    if rank == 0:
        print(json.dumps(line), file=result_out, flush=True)
    if use_dist:
        dist.destroy_process_group()
    if failures:
        print("PARITY FAILURE: " + "; ".join(failures), file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
