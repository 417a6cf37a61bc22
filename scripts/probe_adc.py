"""Dev probe: PQ ADC traversal + exact rerank at 1M x 768."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from islands_b200 import LeannConfig, LeannIndex, ProductQuantizer, PQConfig
n, d, nq = int(os.environ.get("N", 1000000)), 768, 10000
dev = torch.device("cuda:0")
x, q = bench.make_data(torch, os.environ.get("DATASET", "latent32"), n, nq, d, dev)
gt = bench.ground_truth(torch, x, q[:1000], 10)
idx = LeannIndex(LeannConfig()); idx.build_dev(x.data_ptr(), n, d, seed=7, batch=4096)
xh = x.cpu().numpy(); qh = q.cpu().numpy()
for m in [int(v) for v in os.environ.get("MS", "8,16,32").split(",")]:
    pq = ProductQuantizer(d, PQConfig(m, int(os.environ.get("KSUB", 256)), 8, 1))
    t0 = time.time(); pq.train(xh[:20000]); t1 = time.time()
    codes = pq.encode(xh); t2 = time.time()
    idx.attach_pq(pq, codes)
    print(json.dumps(dict(m=m, train_s=round(t1 - t0, 2), encode_s=round(t2 - t1, 2))), flush=True)
    for ef, lim in [(int(v), int(l)) for v in os.environ.get("EFS", "64,128,256,512").split(",") for l in os.environ.get("LIMITS", "0").split(",")]:
        idx.set_rerank_limit(lim)
        ids, dist, cnt, st = idx.search_adc_rerank_batch(qh, 10, ef, stats=True)
        ids, dist, cnt = idx.search_adc_rerank_batch(qh, 10, ef)
        ms, _ = idx.last_search_timing()
        rec = bench.recall_at_k(torch, torch.from_numpy(ids[:1000].astype(np.int64)).to(dev), gt)
        b = st.n_adc.sum() * m + st.n_edge.sum() * 4 + st.n_hop.sum() * 16 + st.n_rerank.sum() * 4 * d + nq * (4 * d + 120 + m * int(os.environ.get("KSUB", 256)) * 4)
        print(json.dumps(dict(m=m, ef=ef, limit=lim, kernel_ms=round(ms, 2), qps=round(nq / ms * 1e3), recall=round(rec, 4), n_adc=float(st.n_adc.mean()),
                              n_rerank=float(st.n_rerank.mean()), n_hop=float(st.n_hop.mean()), gbps=round(b / ms / 1e6, 1))), flush=True)
