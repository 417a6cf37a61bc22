"""Searcher / MultiIndexSearcher / SearchConfig / SearchResult (src/core/search.rs): the reference's
own tests restated (search.rs:251-420: config presets, builder, to_similarity, min_similarity,
include_vectors, batch, multi-index merge), with the oracle HNSW as the checker."""
import numpy as np
import pytest

from conftest import uniform


def test_search_config_and_result_host_logic():
    from islands_b200 import SearchConfig, SearchResult

    c = SearchConfig()
    assert (c.top_k, c.ef, c.include_vectors, c.include_metadata, c.min_similarity) == (10, 100, False, True, None)
    assert (SearchConfig.fast(5).top_k, SearchConfig.fast(5).ef) == (5, 10)
    assert (SearchConfig.accurate(5).top_k, SearchConfig.accurate(5).ef) == (5, 50)
    r = SearchResult(3, 1.0).with_text("x").with_metadata({"a": 1})
    assert r.id == 3 and r.text == "x" and r.metadata == {"a": 1} and r.vector is None
    assert r.to_similarity() == np.float32(0.5)             # search.rs:99-102
    assert SearchResult(0, 0.0).to_similarity() == np.float32(1.0)


@pytest.mark.gpu
def test_searcher_matches_oracle_and_filters(gpu_lib, orc):
    from islands_b200 import HnswConfig, HnswGraph, SearchConfig, Searcher

    cfg = HnswConfig(m=8, m0=16, ef_construction=40, ml=0.9)
    n, d = 500, 24
    v = uniform(np.random.RandomState(2), n, d)
    lv = orc.draw_levels(4, n, cfg.ml, cfg.max_layers)
    og = orc.Hnsw(cfg._s, d)
    og.insert_batch(v, lv, batch=16, threads=8)
    g = HnswGraph(cfg)
    g.insert_batch(v, lv, batch=16)
    q = uniform(np.random.RandomState(3), 30, d)
    s = Searcher(g).top_k(7).ef(40)
    res = s.search_batch(q)
    o_ids, o_dist, o_cnt = og.search(q, 7, 40, threads=8)
    for i in range(30):
        assert [r.id for r in res[i]] == o_ids[i, :o_cnt[i]].tolist()
        assert [r.score for r in res[i]] == o_dist[i, :o_cnt[i]].tolist()
    one = s.search(q[0])
    assert [r.id for r in one] == [r.id for r in res[0]]
    # include_vectors returns the stored vector of each hit (search.rs:160-164)
    wv = Searcher(g, SearchConfig(top_k=3, ef=20, include_vectors=True)).search(v[11])
    assert all(np.array_equal(r.vector, v[r.id]) for r in wv)
    # min_similarity keeps results with 1/(1+d) >= threshold (search.rs:171-173)
    thr = float(np.median([r.to_similarity() for r in res[0]]))
    kept = Searcher(g).top_k(7).ef(40).min_similarity(thr).search(q[0])
    assert [r.id for r in kept] == [r.id for r in res[0] if r.to_similarity() >= np.float32(thr)]
    assert Searcher(g).search_batch(np.zeros((0, d), np.float32)) == []


@pytest.mark.gpu
def test_multi_index_searcher_merges_like_the_reference(gpu_lib, orc):
    from islands_b200 import HnswConfig, HnswGraph, MultiIndexSearcher, SearchConfig

    cfg = HnswConfig(m=8, m0=16, ef_construction=40, ml=0.9)
    d = 16
    rng = np.random.RandomState(7)
    ms = MultiIndexSearcher().with_config(SearchConfig(top_k=6, ef=30))
    assert ms.search(np.zeros(d, np.float32)) == [] and ms.num_indexes() == 0
    oracles = []
    shared = uniform(rng, 20, d)  # the same vectors in two islands: exact score ties across islands
    for name, n in (("alpha", 150), ("beta", 90), ("gamma", 120)):
        v = uniform(rng, n, d)
        if name != "gamma":
            v[:20] = shared
        lv = orc.draw_levels(len(name), n, cfg.ml, cfg.max_layers)
        og = orc.Hnsw(cfg._s, d)
        og.insert_batch(v, lv, batch=8, threads=8)
        g = HnswGraph(cfg)
        g.insert_batch(v, lv, batch=8)
        ms.add_index(name, g)
        oracles.append((name, og))
    assert ms.num_indexes() == 3 and ms.total_vectors() == 360
    q = np.concatenate([shared[:5], uniform(rng, 10, d)])
    got = ms.search_batch(q)
    for i in range(q.shape[0]):
        ref = []
        for name, og in oracles:  # search.rs:214-228, then stable sort by score and truncate (:231-234)
            ids, dist, cnt = og.search(q[i:i + 1], 6, 30)
            ref += [(name, int(ids[0, j]), dist[0, j]) for j in range(cnt[0])]
        ref.sort(key=lambda t: t[2])  # Python's sort is stable, like slice::sort_by
        ref = ref[:6]
        assert [(nm, r.id, r.score) for nm, r in got[i]] == ref
