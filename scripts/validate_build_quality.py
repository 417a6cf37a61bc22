"""Is the batched (round-model) construction as good a graph as the reference's sequential build?

`batch = 1` on the GPU IS the reference loop (leann.rs:578-615: one insert at a time; CSR bit-exact with the
oracle's line-by-line build, tests/test_build_parity.py).  The bench's graphs come from rounds of 4096 inserts
(DESIGN.md 3.3).  This script builds both over the same vectors and levels and compares what SURVEY 7 asks for:
degree histogram, hub retention, recall@10 at equal ef.  Output: one JSON document (profiles/r02_build_quality.json).

  python scripts/validate_build_quality.py [--n 100000] [--d 768] [--dataset latent32|uniform] [--out FILE]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def graph_stats(g, n, hub_frac=0.02):
    deg = np.diff(g.node_offsets.astype(np.int64))
    indeg = np.bincount(g.neighbors.astype(np.int64), minlength=n)
    nh = max(1, int(np.ceil(n * hub_frac)))
    hubs = np.argsort(-indeg, kind="stable")[:nh]
    return {
        "edges": int(g.neighbors.size), "out_degree_mean": float(deg.mean()), "out_degree_hist": np.bincount(deg, minlength=61).tolist(),
        "in_degree_mean": float(indeg.mean()), "in_degree_p50_p90_p99_max": [float(np.percentile(indeg, p)) for p in (50, 90, 99)] + [int(indeg.max())],
        "in_degree_zero_nodes": int((indeg == 0).sum()),
        "hub_in_degree_mean_top2pct": float(indeg[hubs].mean()), "hub_share_of_edges_top2pct": float(indeg[hubs].sum() / max(1, indeg.sum())),
    }, set(hubs.tolist()), indeg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--nodes", dest="n", type=int, default=100_000)
    ap.add_argument("--d", "--dim", dest="d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=2000)
    ap.add_argument("--dataset", default="latent32")
    ap.add_argument("--batches", default="1,1024,4096")
    ap.add_argument("--efs", default="32,64,104,128,256")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch

    import bench
    from islands_b200 import LeannConfig, LeannIndex

    dev = torch.device("cuda", 0)
    x, q = bench.make_data(torch, a.dataset, a.n, a.nq, a.d, dev)
    gt = bench.ground_truth(torch, x, q, 10)
    cfg = LeannConfig()
    ids = torch.empty((a.nq, 10), dtype=torch.int64, device=dev)
    dst = torch.empty((a.nq, 10), dtype=torch.float32, device=dev)
    cnt = torch.empty((a.nq,), dtype=torch.int32, device=dev)
    efs = [int(v) for v in a.efs.split(",")]
    out = {"what": "LeannIndex::build, sequential (batch=1 = the reference loop, leann.rs:578-615) vs round model", "n": a.n, "d": a.d,
           "dataset": a.dataset, "queries": a.nq, "config": "m=30 m0=60 efC=128 hub 2% cosine", "graphs": {}}
    ref_hubs = None
    for b in [int(v) for v in a.batches.split(",")]:
        idx = LeannIndex(cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx.build_dev(x.data_ptr(), a.n, a.d, seed=7, batch=b)  # the same level stream for every batch size
        dt = time.perf_counter() - t0
        g = idx.graph
        st, hubs, indeg = graph_stats(g, a.n)
        rec = {}
        for ef in efs:
            idx.search_batch_dev(q.data_ptr(), a.nq, a.d, 10, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr())
            rec[str(ef)] = bench.recall_at_k(torch, ids, gt)
        if ref_hubs is None:
            ref_hubs = hubs
        st.update({"build_s": dt, "recall_at_10_by_ef": rec, "hub_overlap_with_sequential": len(hubs & ref_hubs) / len(ref_hubs),
                   "build_stats": idx.last_build_stats()})
        out["graphs"][f"batch_{b}"] = st
        print(f"batch {b}: {dt:.1f}s, edges {st['edges']}, recall {rec}", file=sys.stderr, flush=True)
        idx.free()
    keys = list(out["graphs"])
    base = out["graphs"][keys[0]]["recall_at_10_by_ef"]
    out["recall_delta_vs_sequential"] = {k: {ef: out["graphs"][k]["recall_at_10_by_ef"][ef] - base[ef] for ef in base} for k in keys[1:]}
    out["max_abs_recall_delta"] = max([abs(v) for k in out["recall_delta_vs_sequential"].values() for v in k.values()] or [0.0])
    text = json.dumps(out, indent=1)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
