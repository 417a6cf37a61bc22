"""Frozen outputs of the oracle on small seeded inputs: graphs (sequential and round-model construction), search results
for the four metrics with and without frontier pruning, PQ codebooks / codes / tables, HNSW layer lists and search
results, two-level and ADC + rerank results.  The reference holds no golden results of this kind and cannot run
here, so these values come from the oracle at the commit that wrote them — at which point every one of them had also
been reproduced by the independent second reading (tests/test_oracle_second_reading.py) and, for the algorithms the GPU
implements, bit for bit by the CUDA kernels.  Their job is to hold that state still: a later change that moves the
oracle (and with it, silently, the parity target of the GPU tests) fails tests/test_oracle_golden.py.
Usage: python tests/golden/make_search_golden.py   (writes tests/golden/search_golden.npz)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def uniform(rng, n, d):
    return (rng.rand(n, d).astype(np.float32) * 2 - 1).astype(np.float32)


def compute():
    from islands_b200 import HnswConfig, LeannConfig
    from oracle import pyoracle as orc

    out = {}
    n, d, nq = 500, 32, 16
    rng = np.random.RandomState(2024)
    v = uniform(rng, n, d)
    v[n - 50:] = v[:50]  # exact copies: distance ties
    q = uniform(rng, nq, d)
    for metric in range(4):
        cfg = LeannConfig(metric=metric)  # paper defaults: m = 30, m0 = 60, efC = 128 (leann.rs:386-403)
        levels = orc.draw_levels(5, n, cfg.ml, cfg.max_layers)
        off, nbrs, entry, max_level = orc.leann_build(cfg._s, v, levels)
        out[f"m{metric}_off"], out[f"m{metric}_nbrs"], out[f"m{metric}_levels"] = off, nbrs, levels
        out[f"m{metric}_entry"] = np.array([entry, max_level], np.int64)
        ids, dist, cnt, st = orc.leann_search(cfg._s, v, off, nbrs, entry, q, 10, 48, stats=True)
        out[f"m{metric}_ids"], out[f"m{metric}_dist"], out[f"m{metric}_cnt"] = ids, dist.view(np.uint32), cnt
        out[f"m{metric}_stats"] = np.stack([st["n_hop"], st["n_edge"], st["n_dist"]])
        if metric == 0:
            for strategy, ratio in ((0, 0.5), (1, 0.5), (2, 0.5)):
                pc = LeannConfig(metric=0, pruning_strategy=strategy, prune_ratio=ratio, prune_seed=77)
                pi, pd, _ = orc.leann_search(pc._s, v, off, nbrs, entry, q, 10, 48)
                out[f"prune{strategy}_ids"], out[f"prune{strategy}_dist"] = pi, pd.view(np.uint32)
            r_off, r_nbrs, r_entry, _ = orc.leann_build(cfg._s, v, levels, batch=32)
            out["round32_off"], out["round32_nbrs"], out["round32_entry"] = r_off, r_nbrs, np.array([r_entry], np.int64)
            cb = orc.pq_train(1, v, 4, 16, 6, 42)  # dsub = 8
            codes = orc.pq_encode(1, cb, v)
            out["pq_codebooks"], out["pq_codes"] = cb.view(np.uint32), codes
            out["pq_tables"] = orc.pq_build_tables(cb, q[0]).view(np.uint32)
            ti, td, _ = orc.leann_search_two_level(cfg._s, v, off, nbrs, entry, cb, codes, q, 10, 48, 0.25)
            out["two_level_ids"], out["two_level_dist"] = ti, td.view(np.uint32)
            ai, ad, _ = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, 10, 48)
            out["adc_ids"], out["adc_dist"] = ai, ad.view(np.uint32)
    hcfg = HnswConfig(m=6, m0=12, ef_construction=30, ml=0.7)
    hl = orc.draw_levels(8, n, hcfg.ml, hcfg.max_layers)
    g = orc.Hnsw(hcfg._s, d)
    g.insert_batch(v, hl)
    lists = []
    for i in range(n):
        for layer in range(int(hl[i]) + 1):
            lists.append(np.concatenate([[i, layer], g.neighbors(i, layer)]).astype(np.int64))
    out["hnsw_lists"] = np.concatenate([np.concatenate([[len(x)], x]) for x in lists]).astype(np.int64)
    out["hnsw_entry"] = np.array([g.entry_point(), g.max_level()], np.int64)
    hi, hd, hc = g.search(q, 10, 40)
    out["hnsw_ids"], out["hnsw_dist"], out["hnsw_cnt"] = hi, hd.view(np.uint32), hc
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "search_golden.npz")
    np.savez_compressed(path, **compute())
    print(path, os.path.getsize(path), "bytes")
