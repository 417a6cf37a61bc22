"""Dev helper: build once, time the search kernel at a few ef values (kernel ms via handle events)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from islands_b200 import LeannConfig, LeannIndex
n, d, nq = int(os.environ.get("N", 1000000)), 768, 10000
dev = torch.device("cuda:0")
x, q = bench.make_data(torch, os.environ.get("DATASET", "latent32"), n, nq, d, dev)
idx = LeannIndex(LeannConfig()); idx.build_dev(x.data_ptr(), n, d, seed=7, batch=4096)
ids = torch.empty((nq, 10), dtype=torch.int64, device=dev); dst = torch.empty((nq, 10), dtype=torch.float32, device=dev)
cnt = torch.empty((nq,), dtype=torch.int32, device=dev); stats = torch.zeros((nq, 5), dtype=torch.int64, device=dev)
for ef in [int(e) for e in os.environ.get("EFS", "128").split(",")]:
    idx.search_batch_dev(q.data_ptr(), nq, d, 10, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), stats.data_ptr())
    b, pq = bench.algorithmic_bytes(stats.cpu().numpy(), d, nq, 10)
    ms = []
    for _ in range(5):
        idx.search_batch_dev(q.data_ptr(), nq, d, 10, ef, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), 0)
        ms.append(idx.last_search_timing()[0])
    m = float(np.median(ms))
    print(json.dumps(dict(lib=os.environ.get("ISL_DEV_LIB_PATH", "default"), ef=ef, kernel_ms=round(m, 2), gbps=round(b / m / 1e6, 1), frac=round(b / m / 1e6 / 6543.1, 4), n_dist=pq["n_dist"])), flush=True)
