"""HNSW (hnsw.rs): GPU insert/search vs the oracle on identical vectors and levels — every layer's
lists, entry point, max_level, ids and distances bit-exact — plus the reference's own property
tests (hnsw.rs:533-700: insert ids, dimension check, search sortedness / len, empty graph)."""
import numpy as np
import pytest

from conftest import uniform


def _levels(orc, cfg, n, seed):
    return orc.draw_levels(seed, n, cfg.ml, cfg.max_layers)


def _oracle_hnsw(orc, cfg, v, levels, batch):
    g = orc.Hnsw(cfg._s, v.shape[1])
    g.insert_batch(v, levels, batch=batch, threads=8)
    return g


def test_oracle_round_model_batch1_is_sequential_insert(orc):
    """CPU: orc_hnsw_insert_batch(batch=1) == the line-by-line restatement orc_hnsw_insert."""
    from islands_b200 import HnswConfig

    cfg = HnswConfig(m=6, m0=12, ef_construction=24, ml=0.9)
    v = uniform(np.random.RandomState(1), 400, 16)
    lv = _levels(orc, cfg, 400, 2)
    a = orc.Hnsw(cfg._s, 16)
    for i in range(400):
        assert a.insert(v[i], int(lv[i])) == i
    b = _oracle_hnsw(orc, cfg, v, lv, 1)
    assert len(a) == len(b) == 400
    assert a.entry_point() == b.entry_point() and a.max_level() == b.max_level()
    for i in range(400):
        for layer in range(int(lv[i]) + 2):
            x, y = a.neighbors(i, layer), b.neighbors(i, layer)
            assert (x is None) == (y is None)
            if x is not None:
                assert np.array_equal(x, y), (i, layer)


def _compare_graph(orc_g, g, levels):
    assert len(g) == len(orc_g)
    assert g.entry_point == orc_g.entry_point()
    assert g.max_level == orc_g.max_level()
    n = len(g)
    for layer in range(int(levels.max()) + 1):
        deg, nb = g.export_layer(layer)
        for i in range(n):
            ref = orc_g.neighbors(i, layer)
            if ref is None:
                assert deg[i] == -1, (i, layer)
                continue
            assert deg[i] == len(ref), (i, layer, deg[i], len(ref))
            assert np.array_equal(nb[i, :deg[i]], ref), (i, layer)


@pytest.mark.gpu
@pytest.mark.parametrize("metric", [0, 1, 2, 3])
@pytest.mark.parametrize("batch", [1, 32])
def test_hnsw_insert_and_search_match_oracle(gpu_lib, orc, metric, batch):
    from islands_b200 import HnswConfig, HnswGraph

    n, d = (300, 24) if batch == 1 else (1500, 24)
    cfg = HnswConfig(m=8, m0=16, ef_construction=40, metric=metric, ml=0.9)
    v = uniform(np.random.RandomState(21 + metric), n, d)
    lv = _levels(orc, cfg, n, 5)
    assert lv.max() >= 2  # the upper layers are exercised
    og = _oracle_hnsw(orc, cfg, v, lv, batch)
    g = HnswGraph(cfg)
    assert g.insert_batch(v, lv, batch=batch) == 0
    _compare_graph(og, g, lv)
    q = uniform(np.random.RandomState(8), 150, d)
    for k, ef in [(1, 1), (10, 50), (20, 20)]:
        ids, dist, cnt = g.search_batch(q, k, ef)
        o_ids, o_dist, o_cnt = og.search(q, k, ef, threads=8)
        assert np.array_equal(cnt, o_cnt)
        assert np.array_equal(ids, o_ids)
        assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))


@pytest.mark.gpu
def test_hnsw_incremental_inserts_grow_the_graph(gpu_lib, orc):
    """Several insert calls (device arrays grow, a new node can become the entry point)."""
    from islands_b200 import HnswConfig, HnswGraph

    cfg = HnswConfig(m=8, m0=16, ef_construction=32, ml=0.8)
    n, d = 900, 40
    v = uniform(np.random.RandomState(3), n, d)
    lv = _levels(orc, cfg, n, 9)
    lv[700] = lv.max() + 1  # a late node above the current top layer (hnsw.rs:322-325)
    og = orc.Hnsw(cfg._s, d)
    g = HnswGraph(cfg)
    for s, e, batch in [(0, 1, 1), (1, 50, 1), (50, 400, 16), (400, 900, 64)]:
        og.insert_batch(v[s:e], lv[s:e], batch=batch, threads=8)
        assert g.insert_batch(v[s:e], lv[s:e], batch=batch) == s
    _compare_graph(og, g, lv)
    assert g.entry_point == 700
    node = g.get_node(700)
    assert node.level == lv[700] and len(node.connections) == lv[700] + 1
    assert np.array_equal(node.neighbors_at(0), og.neighbors(700, 0))
    q = uniform(np.random.RandomState(4), 64, d)
    ids, dist, cnt = g.search_batch(q, 10, 64)
    o_ids, o_dist, o_cnt = og.search(q, 10, 64, threads=8)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))


@pytest.mark.gpu
def test_hnsw_768d_default_config(gpu_lib, orc):
    from islands_b200 import HnswConfig, HnswGraph

    cfg = HnswConfig()  # m=16, m0=32, efC=200 (hnsw.rs:37-48)
    n, d = 1200, 768
    v = uniform(np.random.RandomState(13), n, d)
    lv = _levels(orc, cfg, n, 17)
    og = _oracle_hnsw(orc, cfg, v, lv, 64)
    g = HnswGraph(cfg)
    g.insert_batch(v, lv, batch=64)
    _compare_graph(og, g, lv)
    q = uniform(np.random.RandomState(14), 64, d)
    ids, dist, cnt = g.search_batch(q, 10, 100)
    o_ids, o_dist, _ = og.search(q, 10, 100, threads=8)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))


@pytest.mark.gpu
def test_hnsw_reference_properties(gpu_lib):
    """hnsw.rs:571-700: ids count up from 0, wrong dimension is rejected, results are sorted and
    have k entries, empty graph returns nothing, the stored vector is its own nearest neighbour."""
    from islands_b200 import DimensionMismatch, HnswConfig, HnswGraph

    g = HnswGraph(HnswConfig())
    assert g.is_empty() and g.dimension() is None and g.entry_point is None
    assert g.search(np.zeros(8, np.float32), 5, 10) == []
    v = uniform(np.random.RandomState(0), 200, 32)
    for i in range(10):
        assert g.insert(v[i], seed=3) == i
    g.insert_batch(v[10:], seed=3, batch=16)
    assert len(g) == 200 and g.dimension() == 32
    with pytest.raises(DimensionMismatch):
        g.insert(np.zeros(16, np.float32))
    with pytest.raises(DimensionMismatch):
        g.search(np.zeros(16, np.float32), 5, 10)
    res = g.search(v[17], 10, 50)
    assert len(res) == 10
    assert all(res[i][1] <= res[i + 1][1] for i in range(9))
    assert res[0][0] == 17 and abs(res[0][1]) < 1e-6
    assert g.get_node(10_000) is None
