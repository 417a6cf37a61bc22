"""Measurement scaffold (not product): exact kNN graph via torch on the GPU -> LeannIndex.from_csr
-> batched search; prints kernel time, counters, algorithmic bytes, GB/s and recall@10."""
import argparse
import json
import sys
import time
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islands_b200 import LeannConfig, LeannIndex  # noqa: E402


def knn_graph(x, deg, chunk=4096):
    xn = torch.nn.functional.normalize(x, dim=1)
    n = x.shape[0]
    out = torch.empty((n, deg), dtype=torch.int64, device=x.device)
    for s in range(0, n, chunk):
        sim = xn[s:s + chunk] @ xn.T
        idx = torch.arange(s, min(n, s + chunk), device=x.device)
        sim[torch.arange(idx.numel(), device=x.device), idx] = -2.0
        out[s:s + chunk] = sim.topk(deg, dim=1).indices
    return out


def ground_truth(x, q, k, chunk=1024):
    xn = torch.nn.functional.normalize(x, dim=1)
    qn = torch.nn.functional.normalize(q, dim=1)
    out = []
    for s in range(0, q.shape[0], chunk):
        out.append((qn[s:s + chunk] @ xn.T).topk(k, dim=1).indices)
    return torch.cat(out).cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--deg", type=int, default=60)
    ap.add_argument("--efs", type=str, default="64,256")
    ap.add_argument("--clusters", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    if a.clusters:
        centers = torch.rand((a.clusters, a.d), generator=g, device=dev) * 2 - 1
        assign = torch.randint(0, a.clusters, (a.n,), generator=g, device=dev)
        x = centers[assign] + 0.15 * torch.randn((a.n, a.d), generator=g, device=dev)
        qa = torch.randint(0, a.clusters, (a.nq,), generator=g, device=dev)
        q = centers[qa] + 0.15 * torch.randn((a.nq, a.d), generator=g, device=dev)
    else:
        x = torch.rand((a.n, a.d), generator=g, device=dev) * 2 - 1
        q = torch.rand((a.nq, a.d), generator=g, device=dev) * 2 - 1
    t0 = time.time()
    nb = knn_graph(x, a.deg)
    torch.cuda.synchronize()
    print(f"knn graph {time.time() - t0:.1f}s", flush=True)
    gt = ground_truth(x, q, 10)
    xh = x.cpu().numpy()
    qh = q.cpu().numpy()
    offsets = np.arange(0, (a.n + 1) * a.deg, a.deg, dtype=np.uint64)
    nbrs = nb.cpu().numpy().astype(np.uint64).reshape(-1)
    del nb, x, q
    torch.cuda.empty_cache()
    cfg = LeannConfig()
    t0 = time.time()
    idx = LeannIndex.from_csr(cfg, xh, offsets, nbrs, None, 0)
    print(f"from_csr {time.time() - t0:.1f}s", flush=True)
    for ef in [int(e) for e in a.efs.split(",")]:
        for rep in range(a.reps):
            t0 = time.time()
            ids, dist, cnt, st = idx.search_batch(qh, 10, ef, stats=True)
            wall = time.time() - t0
            ms, _ = idx.last_search_timing()
        recall = np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(a.nq)])
        nd, ne, nh = st.n_dist.sum(), st.n_edge.sum(), st.n_hop.sum()
        bytes_ = nd * 4 * a.d + ne * 4 + nh * 16 + a.nq * (4 * a.d + 12 * 10)
        print(json.dumps(dict(n=a.n, d=a.d, nq=a.nq, ef=ef, kernel_ms=round(ms, 3), wall_ms=round(wall * 1e3, 1),
                              qps_kernel=round(a.nq / ms * 1e3, 1), recall10=round(float(recall), 4),
                              n_dist=float(nd) / a.nq, n_hop=float(nh) / a.nq, n_edge=float(ne) / a.nq,
                              gbytes=round(bytes_ / 1e9, 3), gbps=round(bytes_ / ms / 1e6, 1),
                              frac_of_6543=round(bytes_ / ms / 1e6 / 6543.1, 4))), flush=True)


if __name__ == "__main__":
    main()
