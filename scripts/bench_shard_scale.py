"""One GPU's share of a 100M-vector index: builds and searches a 12.5M x 768 island (= 100M / 8, north_star's
sharded target) on ONE B200 and prints a JSON line with QPS at recall@10 >= 0.95, the kernel's algorithmic
bandwidth and the build time, then (ADC=1, default) the PQ ADC traversal + exact rerank mode on the same shard.
Islands are independent (island-routed queries), so the 8-GPU
number is this per-GPU figure times the scaling measured by `bench.py --gpus 8` on 1M islands.
Run under gpurun:  N=12500000 python scripts/bench_shard_scale.py
Memory: vectors 38.4 GB in the index + the same again while the synthetic set is alive (ground truth is
computed chunk by chunk before the build, the set is freed after it)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from islands_b200 import LeannConfig, LeannIndex, PQConfig, ProductQuantizer

n = int(os.environ.get("N", 12_500_000))
d, nq, k, n_gt = 768, 10_000, 10, 1000
chunk = 500_000
dev = torch.device("cuda:0")

g = torch.Generator(device=dev)
g.manual_seed(42)
basis = torch.randn((32, d), generator=g, device=dev) / 32 ** 0.5  # the latent basis of bench.make_data("latent32")
x = torch.empty((n, d), device=dev)
for s in range(0, n, chunk):
    e = min(n, s + chunk)
    g.manual_seed(1000 + s // chunk)
    x[s:e] = torch.randn((e - s, 32), generator=g, device=dev) @ basis
    x[s:e] += 0.05 * torch.randn((e - s, d), generator=g, device=dev)
g.manual_seed(43)
q = torch.randn((nq, 32), generator=g, device=dev) @ basis + 0.05 * torch.randn((nq, d), generator=g, device=dev)

# exact cosine top-k of the first n_gt queries, chunk by chunk (measurement infrastructure)
qn = torch.nn.functional.normalize(q[:n_gt], dim=1)
best_v = torch.full((n_gt, k), -2.0, device=dev)
best_i = torch.zeros((n_gt, k), dtype=torch.int64, device=dev)
for s in range(0, n, chunk):
    e = min(n, s + chunk)
    t = (qn @ torch.nn.functional.normalize(x[s:e], dim=1).T).topk(k, dim=1)
    v = torch.cat([best_v, t.values], dim=1)
    i = torch.cat([best_i, t.indices + s], dim=1)
    o = v.topk(k, dim=1)
    best_v, best_i = o.values, torch.gather(i, 1, o.indices)
gt = best_i
del qn, best_v, t, v, i, o

torch.cuda.synchronize()
t0 = time.perf_counter()
idx = LeannIndex(LeannConfig())
idx.build_dev(x.data_ptr(), n, d, seed=7, batch=int(os.environ.get("BATCH", 4096)))
torch.cuda.synchronize()
build_s = time.perf_counter() - t0
adc = os.environ.get("ADC", "1") != "0"
if adc:  # PQ codes for the ADC traversal + exact rerank mode, encoded chunk by chunk through the host API
    pq_m, pq_ksub = 32, 128
    pq = ProductQuantizer(d, PQConfig(pq_m, pq_ksub, 8, 1))
    t0 = time.perf_counter()
    pq.train(x[:20000].cpu().numpy())
    codes = np.concatenate([pq.encode(x[s:min(n, s + chunk)].cpu().numpy()) for s in range(0, n, chunk)])
    idx.attach_pq(pq, codes)
    pq_s = time.perf_counter() - t0
    del codes
del x
torch.cuda.empty_cache()

ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
dst = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
stats = torch.zeros((nq, 5), dtype=torch.int64, device=dev)


def recall_for(ef):
    idx.search_batch_dev(q.data_ptr(), nq, d, k, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), 0)
    return bench.recall_at_k(torch, ids[:n_gt], gt)


ef, curve = bench.calibrate_ef(recall_for, 0.95, int(os.environ.get("EF", 0)))
idx.search_batch_dev(q.data_ptr(), nq, d, k, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), stats.data_ptr())
b, per_query = bench.algorithmic_bytes(stats.cpu().numpy(), d, nq, k)
for _ in range(3):
    idx.search_batch_dev(q.data_ptr(), nq, d, k, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), 0)
torch.cuda.synchronize()
steps, ms = 10, []
t0 = time.perf_counter()
for _ in range(steps):
    idx.search_batch_dev(q.data_ptr(), nq, d, k, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), 0)
    ms.append(idx.last_search_timing()[0])
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f)["hbm_gbs"])
except Exception:
    peak = 6543.1
km = float(np.mean(ms))
adc_out = None
if adc:
    qh = q.cpu().numpy()

    def adc_recall(ef_):
        r = idx.search_adc_rerank_batch(qh, k, ef_)[0]
        return bench.recall_at_k(torch, torch.from_numpy(r[:n_gt].astype(np.int64)).to(dev), gt)

    ef_a, curve_a = bench.calibrate_ef(adc_recall, 0.95, int(os.environ.get("EF_ADC", 0)))
    ms_a, t0 = [], time.perf_counter()
    for _ in range(3):
        idx.search_adc_rerank_batch(qh, k, ef_a)
        ms_a.append(idx.last_search_timing()[0])
    wall_a = (time.perf_counter() - t0) / 3
    adc_out = {"pq_m": pq_m, "pq_ksub": pq_ksub, "train_encode_s": pq_s, "ef": ef_a, "recall_at_10": curve_a[ef_a], "recall_curve": curve_a,
               "kernel_ms": float(np.mean(ms_a)), "kernel_qps": nq / float(np.mean(ms_a)) * 1e3, "e2e_qps_host_buffers": nq / wall_a}
print(json.dumps({"adc_rerank": adc_out, "workload": f"{n} x {d} f32 latent32 island on one B200 (1/8 of a 100M-vector index), m=30 m0=60 efC=128 hub 2%, "
                              f"{nq} queries per step, top-{k}, exact traversal, cosine",
                  "build_s": build_s, "ef": ef, "recall_at_10": curve[ef], "recall_curve": curve, "qps": nq / wall, "ms_per_step": wall * 1e3,
                  "kernel_ms": km, "algorithmic_gbps": b / km / 1e6, "frac_of_measured_hbm_peak": b / km / 1e6 / peak,
                  "per_query": per_query, "hbm_gb_resident": torch.cuda.memory_allocated() / 1e9}), flush=True)
