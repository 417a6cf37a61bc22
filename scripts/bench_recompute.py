"""One-off measurement of BASELINE configs[4]: search with on-demand recompute at scale.
An index of N nodes keeps graph + PQ codes + one S-token row per node; the batch's distinct ADC
survivors go through the random-init BERT-base bf16 encoder; exact rerank on the recomputed rows.
The graph is built over the encoder's own outputs (so the recompute results are the stored-vector
results bit for bit — checked here on a sample), tokens come from a clustered table (near-duplicate
chunks), which is what gives the neighbourhoods a graph index needs."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islands_b200 import Encoder, EncoderConfig, LeannConfig, LeannIndex, PQConfig, ProductQuantizer

n, nq, S, k = int(os.environ.get("N", 200000)), int(os.environ.get("NQ", 256)), 64, 10
ef = int(os.environ.get("EF", 128))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(45)
clusters = max(64, n // 50)
base = torch.randint(1, 30000, (clusters, S), generator=g, device=dev, dtype=torch.int32)
def draw(count):
    t = base[torch.randint(0, clusters, (count,), generator=g, device=dev)].clone()
    flip = torch.rand((count, S), generator=g, device=dev) < 0.15
    t[flip] = torch.randint(1, 30000, (int(flip.sum()),), generator=g, device=dev, dtype=torch.int32)
    return t.contiguous(), torch.full((count,), S, device=dev, dtype=torch.int32)
tok, ln = draw(n); qtok, qln = draw(nq)
enc = Encoder(EncoderConfig()).init_random(seed=46, stddev=0.02)
emb = torch.empty((n, 768), device=dev); qemb = torch.empty((nq, 768), device=dev)
t0 = time.perf_counter(); enc.embed_dev(tok.data_ptr(), ln.data_ptr(), n, S, emb.data_ptr()); t_enc_all = time.perf_counter() - t0
enc.embed_dev(qtok.data_ptr(), qln.data_ptr(), nq, S, qemb.data_ptr())
index = LeannIndex(LeannConfig()); index.build_dev(emb.data_ptr(), n, 768, seed=7, batch=4096)
xh = emb.cpu().numpy(); qh = qemb.cpu().numpy()
pq = ProductQuantizer(768, PQConfig(32, 256, 8, 1)); pq.train(xh[:20000]); index.attach_pq(pq, pq.encode(xh))
ids_a, dist_a, _ = index.search_adc_rerank_batch(qh, k, ef)                       # stored vectors
index.set_recompute(enc, tok.cpu().numpy(), ln.cpu().numpy())
ids_b, dist_b, _ = index.search_adc_recompute_batch(qh, k, ef)                    # warm-up + check
t0 = time.perf_counter(); ids_b, dist_b, _ = index.search_adc_recompute_batch(qh, k, ef); dt = time.perf_counter() - t0
info = index.last_recompute()
xn = torch.nn.functional.normalize(emb, dim=1); qn = torch.nn.functional.normalize(qemb, dim=1)
gt = (qn @ xn.T).topk(k, dim=1).indices.cpu().numpy()
rec = float(np.mean([len(set(ids_b[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
ms_enc, fl = enc.last_timing()
# hub-embedding cache (docs/leann-specification.md:661-690): the top HUB_PCT % in-degree rows stay resident
hub = {}
for pct in [float(v) for v in os.environ.get("HUB_PCTS", "2,10").split(",") if v]:
    t0 = time.perf_counter(); index.set_hub_cache(int(n * pct / 100)); t_set = time.perf_counter() - t0
    index.search_adc_recompute_batch(qh, k, ef)
    t0 = time.perf_counter(); ids_c, dist_c, _ = index.search_adc_recompute_batch(qh, k, ef); dtc = time.perf_counter() - t0
    ic = index.last_recompute()
    hub[str(pct)] = dict(cached_nodes=ic["hub_cache_nodes"], build_s=round(t_set, 2), hits=ic["hub_cache_hits"], recomputed=ic["unique_nodes"],
                         batch_ms=round(dtc * 1e3, 1), encoder_ms=round(ic["encoder_ms"], 2),
                         identical=bool(np.array_equal(ids_b, ids_c) and np.array_equal(dist_b.view(np.uint32), dist_c.view(np.uint32))))
index.set_hub_cache(0)
# rerank limit: the traversal keeps ef survivors, only the best LIMIT per query are recomputed
lim = {}
gtn = gt
for limit in [int(v) for v in os.environ.get("LIMITS", "64,32,16").split(",") if v]:
    index.set_rerank_limit(limit)
    index.search_adc_recompute_batch(qh, k, ef)
    t0 = time.perf_counter(); ids_l, _, _ = index.search_adc_recompute_batch(qh, k, ef); dtl = time.perf_counter() - t0
    il = index.last_recompute()
    lim[str(limit)] = dict(recomputed=il["unique_nodes"], batch_ms=round(dtl * 1e3, 1), qps=round(nq / dtl, 1), encoder_ms=round(il["encoder_ms"], 2),
                           recall_at_10=float(np.mean([len(set(ids_l[i].tolist()) & set(gtn[i].tolist())) / k for i in range(nq)])))
index.set_rerank_limit(0)
# ef sweep of the ADC variant: recall is what the table distances let through to the encoder
sweep = {}
for e in [int(v) for v in os.environ.get("EFS", "256,512").split(",") if v]:
    index.search_adc_recompute_batch(qh, k, e)
    t0 = time.perf_counter(); ids_e, _, _ = index.search_adc_recompute_batch(qh, k, e); dte = time.perf_counter() - t0
    ie = index.last_recompute()
    sweep[str(e)] = dict(batch_ms=round(dte * 1e3, 1), qps=round(nq / dte, 1), recomputed=ie["unique_nodes"], encoder_ms=round(ie["encoder_ms"], 2),
                         recall_at_10=float(np.mean([len(set(ids_e[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)])))
# the reference's own recompute loop, hop by hop (isl_index_search_recompute): exact traversal, provider = encoder
hop = {}
nqx = int(os.environ.get("NQ_HOP", 64))
for e in [int(v) for v in os.environ.get("EFS_HOP", "32,64").split(",") if v]:
    t0 = time.perf_counter(); ids_h, dist_h, _, st_h = index.search_recompute_batch(qh[:nqx], k, e, stats=True); dth = time.perf_counter() - t0
    ih = index.last_recompute()
    ids_s, dist_s, _ = index.search_batch(qh[:nqx], k, e)  # stored vectors, plain exact kernel
    hop[str(e)] = dict(queries=nqx, batch_s=round(dth, 2), qps=round(nqx / dth, 2), sequences_encoded=ih["unique_nodes"], encoder_ms=round(ih["encoder_ms"], 1),
                       n_dist_per_query=float(st_h.n_dist.mean()), n_hop_per_query=float(st_h.n_hop.mean()),
                       recall_at_10=float(np.mean([len(set(ids_h[i].tolist()) & set(gt[i].tolist())) / k for i in range(nqx)])),
                       identical_to_stored_exact_search=bool(np.array_equal(ids_h, ids_s) and np.array_equal(dist_h.view(np.uint32), dist_s.view(np.uint32))))
# the encoder in split precision (ISL_ENCODER_BF16X3): what rank-level agreement with f32 embeddings costs
enc3 = Encoder(EncoderConfig(precision=1)).init_random(seed=46, stddev=0.02)
nb = 2048
out3 = torch.empty((nb, 768), device=dev)
for _ in range(2):
    enc3.embed_dev(tok.data_ptr(), ln.data_ptr(), nb, S, out3.data_ptr())
ms3, _ = enc3.last_timing()
enc.embed_dev(tok.data_ptr(), ln.data_ptr(), nb, S, emb.data_ptr()); ms1, _ = enc.last_timing()
split = dict(batch=nb, bf16_ms=round(ms1, 2), bf16x3_ms=round(ms3, 2), max_abs_diff=float((out3 - emb[:nb]).abs().max().item()))
print(json.dumps(dict(ef_sweep=sweep, per_hop_recompute=hop, split_precision=split, rerank_limit=lim, hub_cache=hub, n=n, nq=nq, S=S, ef=ef, index_encode_s=round(t_enc_all, 2), index_encode_seq_per_s=round(n / t_enc_all),
                      recompute_batch_ms=round(dt * 1e3, 1), qps=round(nq / dt, 1), unique_nodes=info["unique_nodes"],
                      traverse_ms=round(info["traverse_ms"], 2), encoder_ms=round(info["encoder_ms"], 2), rerank_ms=round(info["rerank_ms"], 2),
                      encoder_tflops=round(fl / ms_enc / 1e9, 1), recall_at_10=rec,
                      identical_to_stored=bool(np.array_equal(ids_a, ids_b) and np.array_equal(dist_a.view(np.uint32), dist_b.view(np.uint32))))))
