"""Dev probe: recall@10 of the bf16 recompute search vs the fp32 stored-vector search."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from islands_b200 import Encoder, EncoderConfig, LeannConfig, LeannIndex, PQConfig, ProductQuantizer
from oracle.encoder_oracle import bert_embed
from test_recompute_search import _token_table
import test_recompute_search as T

def run(flip, n=3000, nq=2000, S=16, k=10, ef=64, seed=11):
    rng = np.random.RandomState(seed)
    cfg_e = EncoderConfig(vocab_size=2000, hidden_size=128, num_layers=2, num_heads=2, intermediate_size=512, max_position=32)
    enc = Encoder(cfg_e).init_random(seed=3, stddev=0.08)
    base = rng.randint(1, 2000, size=(150, S))
    def draw(count):
        t = base[rng.randint(0, 150, size=count)].copy()
        f = rng.rand(count, S) < flip
        t[f] = rng.randint(1, 2000, size=int(f.sum()))
        ln = rng.randint(S // 2, S + 1, size=count)
        for i in range(count): t[i, ln[i]:] = 0
        return t.astype(np.int32), ln.astype(np.int32)
    (tok, ln), (qtok, qln) = draw(n), draw(nq)
    sd = enc.state_dict()
    vectors = bert_embed(sd, cfg_e, tok, ln); queries = bert_embed(sd, cfg_e, qtok, qln)
    index = LeannIndex(LeannConfig(m=12, m0=24, ef_construction=64)); index.build(vectors, n, seed=5, batch=64)
    pq = ProductQuantizer(128, PQConfig(16, 64, 10, 1)); pq.train(vectors); index.attach_pq(pq, pq.encode(vectors))
    ids_a, da, _ = index.search_adc_rerank_batch(queries, k, ef)
    index.set_recompute(enc, tok, ln)
    ids_b, db, _ = index.search_adc_recompute_batch(queries, k, ef)
    vn = vectors / np.linalg.norm(vectors, axis=1, keepdims=True); qn = queries / np.linalg.norm(queries, axis=1, keepdims=True)
    gt = np.argsort(-(qn @ vn.T), axis=1, kind="stable")[:, :k]
    rec = lambda ids: float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
    gap = np.median(da[:, 9] - da[:, 8])
    print(json.dumps(dict(flip=flip, ra=rec(ids_a), rb=rec(ids_b), delta=rec(ids_a) - rec(ids_b), same=float((ids_a == ids_b).mean()),
                          med_gap_9_10=float(gap), med_d10=float(np.median(da[:, 9])), max_abs_dd=float(np.abs(da - db)[ids_a == ids_b].max()))), flush=True)

for flip in (0.15, 0.3, 0.5):
    run(flip)
