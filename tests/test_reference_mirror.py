"""The reference's own unit tests that no other file of this suite re-expresses yet, one test here per test there
(file:line of the original in every docstring).  Two kinds, both CPU:

* plain-data / host-side behaviour (defaults, serde forms, provider, node, config presets) — through the Python mirror
  of the reference interface, no device involved;
* search / build / quantizer *properties* (sizes, valid ids, result counts under every pruning ratio and beam width) —
  on the oracle, which is what pins the oracle to the reference's own expectations.  The GPU side of the same properties
  follows from the bit-exact GPU-vs-oracle parity tests (test_search_parity / test_build_parity / test_hnsw_parity /
  test_pq_parity), which run on the same kinds of inputs.
"""
import json

import numpy as np
import pytest

from conftest import uniform
from islands_b200 import (CsrGraph, DimensionMismatch, DistanceMetric, HnswConfig, HnswNode, InMemoryEmbeddingProvider,
                          LeannConfig, MultiIndexSearcher, NodeNotFound, PQConfig, PruningStrategy, SearchConfig, SearchResult)
from islands_b200 import serde_json as sj
from islands_b200.storage import DeserializationError


# ---- distance.rs ------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_all_metrics_handle_equal_vectors(orc, metric):
    """distance.rs:330-340 (`test_all_metrics_handle_empty_equal_vectors`): every metric accepts a vector against itself."""
    a = np.array([1.0, 2.0, 3.0, 4.0], np.float32)
    d = orc.distance(metric, a, a)
    assert np.isfinite(d)
    if metric in (0, 1, 3):
        assert abs(d) < 1e-6  # identical vectors: cosine / euclidean / manhattan distance 0
    else:
        assert d == -30.0  # dot-product distance is the negated dot product


def test_calculate_squared_dimension_mismatch():
    """distance.rs:374-380: the length check comes before any arithmetic (and before any device work)."""
    with pytest.raises(DimensionMismatch):
        DistanceMetric(DistanceMetric.Euclidean).calculate_squared([1.0, 2.0], [1.0, 2.0, 3.0])


def test_normalized_function(orc):
    """distance.rs:382-389."""
    v = orc.normalize(np.array([3.0, 4.0], np.float32))
    assert v.shape == (2,) and abs(float(np.sqrt((v * v).sum())) - 1.0) < 1e-6
    assert np.array_equal(v, np.array([0.6, 0.8], np.float32))


def test_batch_calculate_empty():
    """distance.rs:391-399: no vectors, no distances."""
    out = DistanceMetric(DistanceMetric.Cosine).batch_calculate([1.0, 0.0], np.zeros((0, 2), np.float32))
    assert out.shape == (0,)


def test_distance_metric_default_serde_copy_debug():
    """distance.rs:401-438: default Cosine; serde `rename_all = "lowercase"`; Copy / Clone / Debug."""
    assert DistanceMetric() == DistanceMetric.Cosine
    assert sj.metric_to_json(DistanceMetric.Euclidean) == '"euclidean"'
    assert sj.metric_from_json('"manhattan"') == DistanceMetric.Manhattan
    assert sj.metric_to_json(DistanceMetric.DotProduct) == '"dotproduct"' and sj.metric_from_json('"cosine"') == DistanceMetric.Cosine
    m = DistanceMetric(DistanceMetric.DotProduct)
    assert DistanceMetric(m.value) == m and hash(DistanceMetric(m.value)) == hash(m)
    assert DistanceMetric(DistanceMetric.Euclidean).debug_name() == "Euclidean"
    for bad in ('"Euclidean"', '"l2"', "1", "{"):
        with pytest.raises(DeserializationError):
            sj.metric_from_json(bad)


# ---- hnsw.rs ----------------------------------------------------------------------------------------------------------

def test_hnsw_config_presets():
    """hnsw.rs:532-552 (`test_config_default / _fast / _accurate`) and the constructor that ignores its argument (:30-35)."""
    d = HnswConfig()
    assert (d.m, d.m0, d.ef_construction, d.max_layers, d.metric) == (16, 32, 200, 16, DistanceMetric.Cosine)
    assert abs(d.ml - 1.0 / np.log(16.0)) < 1e-15
    f, a = HnswConfig.fast(), HnswConfig.accurate()
    assert (f.m, f.m0, f.ef_construction) == (12, 24, 100) and (a.m, a.m0, a.ef_construction) == (32, 64, 400)
    assert f.max_layers == a.max_layers == 16 and f.ml == d.ml
    n = HnswConfig.new(HnswConfig.accurate())
    assert (n.m, n.m0, n.ef_construction) == (16, 32, 200)
    for c in (d, f, a):
        c.validate()


def test_node_creation_and_neighbors():
    """hnsw.rs:709-734: a node of level L owns L + 1 connection lists; an absent layer is None."""
    node = HnswNode(42, 2, [[], [], []])
    assert node.id == 42 and node.level == 2 and len(node.connections) == 3
    node = HnswNode(0, 1, [[], []])
    assert list(node.neighbors_at(0)) == []
    node.connections[0] += [1, 2]
    assert list(node.neighbors_at(0)) == [1, 2]
    assert node.neighbors_at(5) is None


def _oracle_hnsw(orc, vectors, seed):
    cfg = HnswConfig()
    g = orc.Hnsw(cfg._s, vectors.shape[1])
    levels = orc.draw_levels(seed, vectors.shape[0], cfg.ml, cfg.max_layers)
    return g, levels


@pytest.mark.parametrize("n,dim", [(10, 4), (50, 16), (100, 32)])
def test_graph_sizes(orc, n, dim):
    """hnsw.rs:736-749 and `prop_insert_increases_size` (:752-765): every insert adds exactly one node."""
    v = uniform(np.random.RandomState(999), n, dim)
    g, levels = _oracle_hnsw(orc, v, 999)
    for i in range(n):
        assert g.insert(v[i], int(levels[i])) == i
        assert len(g) == i + 1
    assert g.d == dim and g.entry_point() is not None and g.max_level() == int(levels.max())


@pytest.mark.parametrize("n,k", [(5, 1), (5, 9), (17, 4), (49, 9), (30, 1)])
def test_hnsw_search_returns_valid_ids(orc, n, k):
    """hnsw.rs:767-786: every returned id names an existing node."""
    v = uniform(np.random.RandomState(42), n, 8)
    g, levels = _oracle_hnsw(orc, v, 42)
    g.insert_batch(v, levels)
    ids, dist, cnt = g.search(np.full(8, 0.5, np.float32), min(k, n), 50)
    assert 0 < cnt[0] <= min(k, n)
    for i in ids[0, :cnt[0]]:
        assert i < n and g.node_level(int(i)) is not None
    assert (np.diff(dist[0, :cnt[0]]) >= 0).all()


def test_hnsw_recall_quality(orc):
    """hnsw.rs:806-854: 200 x 32, accurate configuration, the queries are the stored vectors 0, 10, ..., 190 themselves
    (so the true nearest neighbour is the query's own node), k = 1, ef = 100: recall >= 0.35.  The bar is that low for a
    reason the oracle reproduces: `prune_connections` (hnsw.rs:405-446 with :327) drops the id being inserted from a
    full neighbour list, so only the first ~m0 = 64 nodes stay reachable — of the twenty queries exactly those with an
    id below ~64 can find themselves (7 / 20 = 0.35)."""
    n = 200
    v = uniform(np.random.RandomState(42), n, 32)
    cfg = HnswConfig.accurate()
    g = orc.Hnsw(cfg._s, 32)
    levels = orc.draw_levels(7, n, cfg.ml, cfg.max_layers)
    for i in range(n):
        g.insert(v[i], int(levels[i]))
    qid = np.array([i * 10 % n for i in range(20)])
    ids, dist, cnt = g.search(v[qid], 1, 100)
    truth = np.array([np.argmin(orc.distance_batch(0, v[i], v)) for i in qid])
    assert np.array_equal(truth, qid)
    found = ids[:, 0] == truth
    assert (cnt == 1).all() and found.mean() >= 0.35
    assert found[qid < 60].all() and not found[qid >= 80].any()  # the quirk, visible: reachable = the early nodes


# ---- leann.rs ---------------------------------------------------------------------------------------------------------

def test_pruning_strategies():
    """leann.rs:1147-1168: Global is the default; every strategy validates with prune_ratio 0.3."""
    assert LeannConfig().pruning_strategy == PruningStrategy.Global
    for s in (PruningStrategy.Global, PruningStrategy.Local, PruningStrategy.Proportional):
        LeannConfig(pruning_strategy=s, prune_ratio=0.3).validate()


def test_in_memory_provider():
    """leann.rs:1221-1256 (`test_in_memory_provider`, `_batch`, `_invalid_id`) + `with_dimension` / `add` (:123-141)."""
    v = uniform(np.random.RandomState(42), 10, 8)
    p = InMemoryEmbeddingProvider(v)
    assert p.dimension() == 8
    for i in range(10):
        assert np.array_equal(p.compute_embedding(i), v[i])
    batch = p.compute_embeddings_batch([0, 2, 5])
    assert len(batch) == 3 and all(np.array_equal(batch[j], v[i]) for j, i in enumerate([0, 2, 5]))
    with pytest.raises(NodeNotFound):
        p.compute_embedding(999)
    e = InMemoryEmbeddingProvider.with_dimension(4)
    assert e.dimension() == 4 and e.add([1, 2, 3, 4]) == 0 and e.add([5, 6, 7, 8]) == 1
    assert np.array_equal(e.compute_embedding(1), np.array([5, 6, 7, 8], np.float32))
    with pytest.raises(DimensionMismatch):
        e.add([1, 2, 3])


def _oracle_leann(orc, n, dim, seed=42, **cfg_kw):
    cfg = LeannConfig(**cfg_kw)
    v = uniform(np.random.RandomState(seed), n, dim)
    levels = orc.draw_levels(seed, n, cfg.ml, cfg.max_layers)
    off, nbrs, entry, max_level = orc.leann_build(cfg._s, v, levels)
    return cfg, v, levels, off, nbrs, entry


@pytest.mark.parametrize("n", [5, 12, 29])
def test_build_increases_size_and_ids_are_valid(orc, n):
    """leann.rs:1469-1493 (`prop_build_increases_size`, `prop_search_returns_valid_ids`)."""
    cfg, v, _, off, nbrs, entry = _oracle_leann(orc, n, 8)
    assert off.size == n + 1 and int(off[-1]) == nbrs.size and (nbrs < n).all() and 0 <= entry < n
    for k in (1, 4, 9):
        ids, dist, cnt = orc.leann_search(cfg._s, v, off, nbrs, entry, v[:1], min(k, n), cfg.ef_search)
        assert 0 < cnt[0] <= min(k, n) and (ids[0, :cnt[0]] < n).all()


@pytest.mark.parametrize("n,dim", [(20, 8), (57, 33), (99, 63)])
def test_storage_valid(orc, n, dim):
    """leann.rs:1495-1512: positive storage, every vector in the index; the CSR accounting of leann.rs:296-301."""
    cfg, v, levels, off, nbrs, entry = _oracle_leann(orc, n, dim)
    g = CsrGraph()
    g.node_offsets, g.neighbors, g.levels, g.num_nodes = off, nbrs, levels, n
    g.degree_counts = np.diff(off.astype(np.int64)).astype(np.uint64)
    assert g.storage_bytes() == 8 * ((n + 1) + nbrs.size + n + n) > 0


@pytest.mark.parametrize("n,dim", [(10, 4), (50, 16), (100, 32)])
def test_various_sizes(orc, n, dim):
    """leann.rs:1514-1533: 5 results (or n) for the first vector as the query, which finds itself."""
    cfg, v, _, off, nbrs, entry = _oracle_leann(orc, n, dim)
    ids, dist, cnt = orc.leann_search(cfg._s, v, off, nbrs, entry, v[:1], min(5, n), cfg.ef_search)
    assert cnt[0] == min(5, n) and ids[0, 0] == 0 and dist[0, 0] < 0.01


@pytest.mark.parametrize("beam_width", [1, 2, 4])
def test_beam_widths(orc, beam_width):
    """leann.rs:1535-1554: `beam_width` is a stored parameter; the search (leann.rs:899-988) expands one candidate per
    step whatever its value, so the results are the same for every width."""
    base = None
    for bw in (1, beam_width):
        cfg, v, _, off, nbrs, entry = _oracle_leann(orc, 50, 16, beam_width=bw)
        ids, dist, cnt = orc.leann_search(cfg._s, v, off, nbrs, entry, v[:1], 5, cfg.ef_search)
        assert cnt[0] == 5
        base = base or (ids.copy(), dist.copy())
        assert np.array_equal(ids, base[0]) and np.array_equal(dist.view(np.uint32), base[1].view(np.uint32))


@pytest.mark.parametrize("prune_ratio", [0.0, 0.3, 0.5, 0.8])
def test_prune_ratios(orc, prune_ratio):
    """leann.rs:1556-1575: 5 results under every frontier pruning ratio."""
    cfg, v, _, off, nbrs, entry = _oracle_leann(orc, 50, 16, prune_ratio=prune_ratio)
    ids, dist, cnt = orc.leann_search(cfg._s, v, off, nbrs, entry, v[:1], 5, cfg.ef_search)
    assert cnt[0] == 5 and (np.diff(dist[0]) >= 0).all() and len(set(ids[0].tolist())) == 5


# ---- pq.rs ------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dim", [8, 12, 36, 60])
def test_encode_decode_dimensions(orc, dim):
    """pq.rs:738-760: 4 codes per vector, `dim` floats back."""
    v = uniform(np.random.RandomState(123), 20, dim)
    cb = orc.pq_train(1, v, 4, 16, 3, 42)
    assert cb.shape == (4, 16, dim // 4)
    codes = orc.pq_encode(1, cb, v[:1])
    assert codes.shape == (1, 4) and (codes < 16).all()
    assert orc.pq_decode(cb, codes).shape == (1, dim)


def test_kmeans_basic(orc):
    """pq.rs:811-829: two clusters around (1, 0) and (-1, 0) give one centroid on each side."""
    v = np.array([[1.0, 0.0], [1.1, 0.1], [0.9, -0.1], [-1.0, 0.0], [-1.1, 0.1], [-0.9, -0.1]], np.float32)
    cb = orc.pq_train(1, v, 1, 2, 10, 42)
    assert cb.shape == (1, 2, 2)
    xs = sorted(cb[0, :, 0].tolist())
    assert abs(xs[0] + 1.0) < 0.11 and abs(xs[1] - 1.0) < 0.11 and np.abs(cb[0, :, 1]).max() < 0.11


def test_codebook_host_side():
    """PQCodebook (pq.rs:66-112): construction, `get_centroid`, and the length check of `find_nearest` (pq.rs:87-92), which
    precedes any arithmetic; the values of pq.rs:787-809 themselves are known answers of tests/test_oracle_kat.py (oracle)
    and tests/test_pq_parity.py (GPU)."""
    from islands_b200 import PQCodebook

    cb = PQCodebook(4)
    assert cb.centroids == [] and cb.subvector_dim == 4
    cb.centroids = [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]]
    assert cb.get_centroid(1).tolist() == [0.0, 1.0, 0.0, 0.0] and cb.get_centroid(3) is None
    with pytest.raises(DimensionMismatch):
        cb.find_nearest([0.9, 0.1, 0.0])


# ---- search.rs --------------------------------------------------------------------------------------------------------

def test_multi_index_searcher_empty():
    """search.rs:425-430."""
    s = MultiIndexSearcher()
    assert s.num_indexes() == 0 and s.total_vectors() == 0
    assert s.search(np.full(8, 0.5, np.float32)) == []


def test_search_result_serialization():
    """search.rs:447-459: a result with metadata and text survives serde_json; field order of search.rs:56-67."""
    r = SearchResult(42, 0.5).with_metadata({"file": "test.rs"}).with_text("sample text")
    text = sj.search_result_to_json(r)
    assert text == '{"id":42,"score":0.5,"vector":null,"metadata":{"file":"test.rs"},"text":"sample text"}'
    p = sj.search_result_from_json(text)
    assert (p.id, p.score, p.text, p.metadata, p.vector) == (42, np.float32(0.5), "sample text", {"file": "test.rs"}, None)
    r = SearchResult(7, np.float32(0.1)).with_vector(np.array([0.25, -1.0, 0.1], np.float32))
    text = sj.search_result_to_json(r)
    assert text == '{"id":7,"score":0.1,"vector":[0.25,-1.0,0.1],"metadata":null,"text":null}'  # f32 digits, not f64's
    assert np.array_equal(sj.search_result_from_json(text).vector, r.vector)
    for bad in ("[]", '{"id":-1,"score":0.5}', '{"id":1}', '{"id":1,"score":"x"}', '{"id":1,"score":1,"vector":[true]}'):
        with pytest.raises(DeserializationError):
            sj.search_result_from_json(bad)


def test_search_config_serialization():
    """search.rs:461-480."""
    c = SearchConfig(top_k=20, ef=200, include_vectors=True, include_metadata=False, min_similarity=0.7)
    text = sj.search_config_to_json(c)
    assert text == '{"top_k":20,"ef":200,"include_vectors":true,"include_metadata":false,"min_similarity":0.7}'
    p = sj.search_config_from_json(text)
    assert (p.top_k, p.ef, p.include_vectors, p.include_metadata, p.min_similarity) == (20, 200, True, False, np.float32(0.7))
    d = sj.search_config_from_json(sj.search_config_to_json(SearchConfig()))
    assert (d.top_k, d.ef, d.include_vectors, d.include_metadata, d.min_similarity) == (10, 100, False, True, None)


# ---- the serde forms of the three index configurations ----------------------------------------------------------------

def test_config_json_forms():
    """`#[derive(Serialize, Deserialize)]` of LeannConfig (leann.rs:321-375), HnswConfig (hnsw.rs:13-28), PQConfig
    (pq.rs:12-22): declaration order, compact, enum variants by name, f32 fields with f32 digits."""
    text = sj.leann_config_to_json(LeannConfig())
    assert text == ('{"m":30,"m0":60,"ef_construction":128,"ml":' + sj.format_float(1.0 / np.log(30.0)) + ',"max_layers":16,'
                    '"metric":"cosine","ef_search":64,"beam_width":1,"prune_ratio":0.0,"pruning_strategy":"Global",'
                    '"high_degree_pruning":true,"hub_percentile":0.02,"is_compact":true,"is_recompute":true}')
    assert json.loads(text)["ml"] == 1.0 / np.log(30.0)  # shortest digits that round-trip
    c = LeannConfig.fast()
    c.metric, c.pruning_strategy, c.prune_seed = DistanceMetric.DotProduct, PruningStrategy.Proportional, 9
    back = sj.leann_config_from_json(sj.leann_config_to_json(c))
    for f in LeannConfig._fields:
        assert getattr(back, f) == getattr(c, f), f
    assert '"prune_seed":9' in sj.leann_config_to_json(c) and "prune_seed" not in text
    assert sj.hnsw_config_to_json(HnswConfig()) == ('{"m":16,"m0":32,"ef_construction":200,"ml":' + sj.format_float(1.0 / np.log(16.0))
                                                    + ',"metric":"cosine","max_layers":16}')
    h = sj.hnsw_config_from_json(sj.hnsw_config_to_json(HnswConfig(m=8, m0=24, metric=DistanceMetric.Manhattan)))
    assert (h.m, h.m0, h.ef_construction, h.metric, h.max_layers) == (8, 24, 200, DistanceMetric.Manhattan, 16)
    assert sj.pq_config_to_json(PQConfig()) == '{"num_subquantizers":8,"num_centroids":256,"training_iterations":25,"seed":null}'
    q = sj.pq_config_from_json('{"num_subquantizers":4,"num_centroids":16,"training_iterations":3,"seed":18446744073709551615}')
    assert (q.num_subquantizers, q.num_centroids, q.training_iterations, q.seed) == (4, 16, 3, 2 ** 64 - 1)
    for bad in ('{"m":30}', text.replace('"Global"', '"global"'), text.replace('"cosine"', '"Cosine"'), text.replace("true", "1", 1)):
        with pytest.raises(DeserializationError):
            sj.leann_config_from_json(bad)


def test_format_float_layout():
    """The layout rule of the `ryu` crate that serde_json prints floats with (restated; see the module header)."""
    f = sj.format_float
    assert [f(x) for x in (1.0, 0.5, 100.0, 12.34, 0.001234, 1e15, 1e16, 1e17, 1.5e300, 1e-5, 1e-6, 1.25e-7, -2.5, 0.0, -0.0)] == \
        ["1.0", "0.5", "100.0", "12.34", "0.001234", "1000000000000000.0", "1e16", "1e17", "1.5e300", "0.00001", "1e-6",
         "1.25e-7", "-2.5", "0.0", "-0.0"]
    assert [f(x, f32=True) for x in (0.02, 0.3, 0.7, 1e12, 1e13, 1e-6, 1e-7, 16777216.0)] == \
        ["0.02", "0.3", "0.7", "1000000000000.0", "1e13", "0.000001", "1e-7", "16777216.0"]
    assert f(float("nan")) == f(float("inf")) == "null"
    for x in (0.1, 1 / 3, 2.0 ** 70, 5e-324, 1.7976931348623157e308, 123456.789):
        assert float(f(x)) == x


# ---- bookkeeping: every unit test of the reference's src/core has a counterpart in this suite -------------------------

def test_every_reference_test_is_re_expressed():
    """tests/golden/reference_test_coverage.json (scripts/reference_test_coverage.py): the 131 tests of the reference's
    `#[cfg(test)]` modules for this path, each with the files of this suite that cite its lines.  Where the reference
    checkout is present the table is recomputed and must be current."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    table = json.load(open(os.path.join(root, "tests", "golden", "reference_test_coverage.json")))
    assert len(table) == 131
    for test, files in table.items():
        assert files, test
        for f in files:
            assert os.path.exists(os.path.join(root, f)), (test, f)
    if os.path.isdir("/root/reference/src/core"):
        spec = importlib.util.spec_from_file_location("reference_test_coverage", os.path.join(root, "scripts", "reference_test_coverage.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        tests, cites = mod.reference_tests(), mod.citations()
        now = {f"{f}.rs:{a} {name}": sorted({c[2] for c in cites[f] if not (c[1] < a or c[0] > e)})
               for f, ts in tests.items() for a, name, e in ts}
        assert now == table, "run python scripts/reference_test_coverage.py --write"
