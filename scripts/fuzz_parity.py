"""Randomised parity sweep (run under gpurun): GPU search variants against the oracle on small random graphs with
tie-heavy data (copied vectors, tiny codebooks), duplicate list entries and random k / ef.  Prints one line per
mismatch and a summary; exit code 1 on any mismatch.  SEED / ROUNDS / BUDGET_S from the environment."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from islands_b200 import HnswConfig, HnswGraph, LeannConfig, LeannIndex, PQConfig, ProductQuantizer  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

rng = np.random.RandomState(int(os.environ.get("SEED", 1)))
rounds, budget = int(os.environ.get("ROUNDS", 40)), float(os.environ.get("BUDGET_S", 150))
t_start, bad, checks = time.time(), 0, 0


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2], b[2])


for r in range(rounds):
    if time.time() - t_start > budget:
        break
    n = int(rng.choice([60, 300, 1000, 2500]))
    d = int(rng.choice([32, 64, 96]))
    metric = int(rng.choice([0, 0, 1, 2]))
    m0 = int(rng.choice([8, 24, 60, 100]))
    v = (rng.rand(n, d).astype(np.float32) * 2 - 1)
    ncopy = int(rng.choice([0, n // 10, n // 3]))
    if ncopy:
        v[n - ncopy:] = v[:ncopy]                      # exact distance ties
    if rng.rand() < 0.3:
        v = np.round(v * 4) / 4                        # coarse grid: ties everywhere
    cfg = LeannConfig(metric=metric, m=max(2, m0 // 2), m0=m0, ef_construction=max(m0 + 8, 32))
    levels = orc.draw_levels(r + 5, n, cfg.ml, cfg.max_layers)
    bb = int(rng.choice([1, 16]))
    off, nbrs, entry, _ = orc.leann_build(cfg._s, v, levels, batch=bb, threads=os.cpu_count() or 1)
    nbrs = nbrs.copy()
    # graph construction (round model): GPU CSR against the oracle's, same levels and batch
    try:
        gb = LeannIndex(cfg)
        gb.build(v, n, levels=levels, batch=bb)
        gg = gb.graph
        checks += 1
        if not (np.array_equal(gg.node_offsets, off) and np.array_equal(gg.neighbors[:int(off[-1])], nbrs[:int(off[-1])])
                and int(gg.entry_point) == int(entry)):
            bad += 1
            print(f"MISMATCH build: round {r} n={n} d={d} metric={metric} m0={m0} copies={ncopy} batch={bb}", flush=True)
        del gb
    except Exception as ex:
        print(f"raised (build): round {r} n={n} d={d} m0={m0} copies={ncopy} batch={bb}", type(ex).__name__, str(ex)[:80], flush=True)
    if rng.rand() < 0.4:                               # duplicate list entries
        for node in rng.choice(n, max(1, n // 5), replace=False):
            s, e = int(off[node]), int(off[node + 1])
            if e - s > 3:
                nbrs[s + int(rng.randint(1, e - s))] = nbrs[s + int(rng.randint(0, e - s))]
    # HNSW (hnsw.rs): multi-layer insert (round model) and search, every layer's lists compared
    if n <= 1000:
        try:
            hm = int(rng.choice([4, 8, 16]))
            hcfg = HnswConfig(m=hm, m0=2 * hm, ef_construction=int(rng.choice([16, 40])), metric=metric, ml=float(rng.choice([0.5, 0.9])))
            hl = orc.draw_levels(r + 9, n, hcfg.ml, hcfg.max_layers)
            hb = int(rng.choice([1, 32]))
            og = orc.Hnsw(hcfg._s, d)
            og.insert_batch(v, hl, batch=hb, threads=8)
            hg = HnswGraph(hcfg)
            hg.insert_batch(v, hl, batch=hb)
            ok = len(hg) == len(og) and hg.entry_point == og.entry_point() and hg.max_level == og.max_level()
            for layer in range(int(hl.max()) + 1):
                deg, nb = hg.export_layer(layer)
                for i in range(n):
                    ref = og.neighbors(i, layer)
                    if ref is None:
                        ok = ok and deg[i] == -1
                    else:
                        ok = ok and deg[i] == len(ref) and np.array_equal(nb[i, :deg[i]], ref)
            hq = (rng.rand(32, d).astype(np.float32) * 2 - 1)
            for hk, hef in [(1, 1), (10, 50), (20, 20)]:
                a_ = hg.search_batch(hq, hk, hef)
                b_ = og.search(hq, hk, hef, threads=8)
                ok = ok and same(a_, b_)
            checks += 1
            if not ok:
                bad += 1
                print(f"MISMATCH hnsw: round {r} n={n} d={d} metric={metric} m={hm} batch={hb} copies={ncopy}", flush=True)
        except Exception as ex:
            print(f"raised (hnsw): round {r} n={n} d={d} metric={metric}", type(ex).__name__, str(ex)[:80], flush=True)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    nq = 64
    q = np.concatenate([(rng.rand(nq - 8, d).astype(np.float32) * 2 - 1), v[:8]])
    m = int(rng.choice([m_ for m_ in (2, 4, 8, 16, 32) if d % m_ == 0]))
    ksub = int(rng.choice([2, 4, 16, 64, 256]))
    cb = orc.pq_train(1, v[:min(n, 1000)], m, ksub, 2, r)  # Euclidean quantizer, as PQConfig defaults (pq.rs:37-45)
    codes = orc.pq_encode(1, cb, v)
    pq = ProductQuantizer(d, PQConfig(m, ksub, 2, r))
    pq.set_codebooks(cb)
    idx.attach_pq(pq, codes)
    for _ in range(4):
        k = int(rng.choice([1, 5, 10, 40]))
        ef = int(rng.choice([1, 7, 33, 64, 65, 130, 200, 257, 400, 513, 700]))
        limit = int(rng.choice([0, 0, 3, 20]))
        tag = f"round {r} n={n} d={d} metric={metric} m0={m0} copies={ncopy} pq=({m},{ksub}) k={k} ef={ef} limit={limit}"
        # exact traversal
        try:
            g = idx.search_batch(q, k, ef, stats=True)
            o = orc.leann_search(cfg._s, v, off, nbrs, entry, q, k, ef, threads=8, stats=True)
            checks += 1
            if not same(g, o) or any(not np.array_equal(getattr(g[3], f), o[3][f]) for f in ("n_hop", "n_edge", "n_dist")):
                bad += 1
                print("MISMATCH exact:", tag, flush=True)
        except Exception as ex:  # a loud failure (the tie list limit) is not a parity error; report it
            print("raised (exact):", tag, type(ex).__name__, str(ex)[:80], flush=True)
        # ADC traversal + rerank, with statistics (bitset) and without (bitset-free)
        try:
            idx.set_rerank_limit(limit)
            o = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef, threads=8, stats=True, rerank_limit=limit)
            g = idx.search_adc_rerank_batch(q, k, ef, stats=True)
            checks += 1
            if not same(g, o) or any(not np.array_equal(getattr(g[3], f), o[3][f]) for f in ("n_hop", "n_edge", "n_adc", "n_rerank")):
                bad += 1
                print("MISMATCH adc (statistics):", tag, flush=True)
            g2 = idx.search_adc_rerank_batch(q, k, ef)
            checks += 1
            if not same(g2, o):
                bad += 1
                print("MISMATCH adc (no statistics):", tag, flush=True)
        except Exception as ex:
            print("raised (adc):", tag, type(ex).__name__, str(ex)[:80], flush=True)
        idx.set_rerank_limit(0)
        # two-level search
        ratio = float(rng.choice([0.1, 0.5, 1.0]))
        try:
            o = orc.leann_search_two_level(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef, ratio, threads=8, stats=True)
            g = idx.search_two_level_batch(q, k, ef, ratio, stats=True)
            checks += 1
            if not same(g, o):
                bad += 1
                print("MISMATCH two-level:", tag, "ratio", ratio, flush=True)
        except Exception as ex:  # a loud failure (e.g. the tie list limit) is not a parity error; report it
            print("raised:", tag, type(ex).__name__, str(ex)[:100], flush=True)
print(f"fuzz: {checks} checks, {bad} mismatches, {time.time() - t_start:.0f}s", flush=True)
sys.exit(1 if bad else 0)
