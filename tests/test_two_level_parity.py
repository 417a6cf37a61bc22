"""Two-level search (PQ ADC traversal + exact rerank) vs the oracle's definition
(docs/leann-specification.md:223-269; no reference code => parity is oracle<->GPU)."""
import numpy as np
import pytest

from conftest import oracle_graph, uniform

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,ksub", [(8, 256), (4, 16), (16, 300), (48, 256)])
@pytest.mark.parametrize("ratio", [0.1, 0.5, 1.0])
def test_two_level_matches_oracle(gpu_lib, orc, m, ksub, ratio):
    from islands_b200 import LeannIndex, PQConfig, ProductQuantizer

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 3000, 96, seed=61)
    cb = orc.pq_train(1, v[:1000], m, ksub, 3, 7)
    codes = orc.pq_encode(1, cb, v)
    pq = ProductQuantizer(96, PQConfig(m, ksub, 3, 7))
    pq.set_codebooks(cb)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    idx.attach_pq(pq, codes)
    q = uniform(np.random.RandomState(62), 100, 96)
    for k, ef in [(10, 32), (10, 128)]:
        ids, dist, cnt, st = idx.search_two_level_batch(q, k, ef, ratio, stats=True)
        o_ids, o_dist, o_cnt, o_st = orc.leann_search_two_level(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef,
                                                                ratio, threads=8, stats=True)
        assert np.array_equal(cnt, o_cnt)
        assert np.array_equal(ids, o_ids)
        assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
        for f in ("n_hop", "n_edge", "n_dist", "n_adc", "n_rerank"):
            assert np.array_equal(getattr(st, f), o_st[f]), f
    if ratio == 1.0:
        # promoting everything evaluates every frontier node exactly: recall of the exact search
        e_ids, _, _ = idx.search_batch(q, 10, 128)
        t_ids, _, _ = idx.search_two_level_batch(q, 10, 128, 1.0)
        overlap = np.mean([len(set(e_ids[i]) & set(t_ids[i])) / 10 for i in range(100)])
        assert overlap > 0.9


def test_two_level_needs_pq(gpu_lib, orc):
    from islands_b200 import LeannIndex, PQError

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 3000, 96, seed=61)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    with pytest.raises(PQError):
        idx.search_two_level_batch(v[:2], 5, 32, 0.1)


@pytest.mark.parametrize("m,ksub", [(8, 256), (4, 16), (16, 300), (96, 256), (32, 256), (16, 64)])
def test_adc_traversal_rerank_matches_oracle(gpu_lib, orc, m, ksub):
    """PQ ADC traversal + exact rerank (include/islands_b200.h isl_index_search_adc_rerank)."""
    from islands_b200 import LeannIndex, PQConfig, ProductQuantizer

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 3000, 96, seed=61)
    cb = orc.pq_train(1, v[:1000], m, ksub, 3, 7)
    codes = orc.pq_encode(1, cb, v)
    pq = ProductQuantizer(96, PQConfig(m, ksub, 3, 7))
    pq.set_codebooks(cb)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    idx.attach_pq(pq, codes)
    q = np.concatenate([uniform(np.random.RandomState(63), 90, 96), v[:10]])
    # ef 16..500: result array in registers (2 / 4 / 8 / 10 / 16 entries per lane); 600: shared memory; 2500: global memory
    for k, ef in [(10, 16), (10, 100), (25, 200), (10, 300), (10, 500), (10, 600), (10, 2500)]:
        if ef == 2500 and m not in (32, 8):
            continue
        ids, dist, cnt, st = idx.search_adc_rerank_batch(q, k, ef, stats=True)
        o_ids, o_dist, o_cnt, o_st = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef,
                                                                 threads=8, stats=True)
        assert np.array_equal(cnt, o_cnt)
        assert np.array_equal(ids, o_ids)
        assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
        for f in ("n_hop", "n_edge", "n_dist", "n_adc", "n_rerank"):
            assert np.array_equal(getattr(st, f), o_st[f]), f
        assert (np.diff(dist, axis=1) >= 0).all()  # exact distances, ascending
        # without statistics the traversal (ef <= 256, m = 16 / 32) runs without the visited bitset: nodes may be
        # scored twice, the survivors and therefore the results must not change
        ids2, dist2, cnt2 = idx.search_adc_rerank_batch(q, k, ef)
        assert np.array_equal(cnt2, o_cnt) and np.array_equal(ids2, o_ids)
        assert np.array_equal(dist2.view(np.uint32), o_dist.view(np.uint32))


@pytest.mark.parametrize("m,ksub", [(32, 256), (8, 256)])
def test_adc_traversal_with_duplicate_neighbours(gpu_lib, orc, m, ksub):
    """Adjacency lists with repeated ids (from_csr accepts them): the first occurrence in LIST order
    decides, within one half of a 64-position pass and across its two halves."""
    from islands_b200 import LeannIndex, PQConfig, ProductQuantizer

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 3000, 96, seed=61)
    nbrs = nbrs.copy()
    rng = np.random.RandomState(5)
    for node in rng.choice(3000, 600, replace=False):
        s, e = int(off[node]), int(off[node + 1])
        deg = e - s
        if deg > 40:
            nbrs[s + 33:s + 39] = nbrs[s + 2:s + 8]      # second half repeats ids of the first half
            nbrs[s + 10] = nbrs[s + 9]                    # repeat inside the first half
            nbrs[s + 39] = nbrs[s + 38 - 1 + 1]           # repeat inside the second half (39 == 38)
            nbrs[s + 39] = nbrs[s + 38]
    cb = orc.pq_train(1, v[:1000], m, ksub, 3, 7)
    codes = orc.pq_encode(1, cb, v)
    pq = ProductQuantizer(96, PQConfig(m, ksub, 3, 7))
    pq.set_codebooks(cb)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    idx.attach_pq(pq, codes)
    q = uniform(np.random.RandomState(64), 100, 96)
    for k, ef in [(10, 64), (10, 150)]:
        ids, dist, cnt, st = idx.search_adc_rerank_batch(q, k, ef, stats=True)
        o_ids, o_dist, o_cnt, o_st = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef,
                                                                 threads=8, stats=True)
        assert np.array_equal(ids, o_ids) and np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
        for f in ("n_hop", "n_edge", "n_adc", "n_rerank"):
            assert np.array_equal(getattr(st, f), o_st[f]), f
        ids2, dist2, _ = idx.search_adc_rerank_batch(q, k, ef)  # no statistics: traversal without the visited bitset
        assert np.array_equal(ids2, o_ids) and np.array_equal(dist2.view(np.uint32), o_dist.view(np.uint32))
        e_ids, e_dist, _ = idx.search_batch(q, k, ef)
        x_ids, x_dist, _ = orc.leann_search(cfg._s, v, off, nbrs, entry, q, k, ef, threads=8)
        assert np.array_equal(e_ids, x_ids) and np.array_equal(e_dist.view(np.uint32), x_dist.view(np.uint32))


@pytest.mark.parametrize("m,ksub", [(32, 64), (16, 128)])
def test_adc_traversal_ties_with_and_without_visited_set(gpu_lib, orc, m, ksub):
    """A third of the base vectors are exact copies of others (identical PQ codes => exact ties of
    the table distance between different ids), small ef so that R evicts constantly.  The traversal with
    statistics (visited bitset) and the one without (no visited set: duplicates are recognised in R, the
    worst distance only decreases) must both return the oracle's ids and distances."""
    from islands_b200 import LeannIndex, PQConfig, ProductQuantizer

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 3000, 96, seed=71, dup=1000)
    cb = orc.pq_train(1, v[:1500], m, ksub, 3, 7)
    codes = orc.pq_encode(1, cb, v)
    assert np.array_equal(codes[:1000], codes[2000:])  # the copies carry the same codes
    pq = ProductQuantizer(96, PQConfig(m, ksub, 3, 7))
    pq.set_codebooks(cb)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    idx.attach_pq(pq, codes)
    q = np.concatenate([uniform(np.random.RandomState(72), 150, 96), v[:50]])
    for k, ef in [(5, 8), (10, 24), (10, 64), (20, 130), (10, 256), (10, 330), (10, 450), (10, 530)]:  # registers up to 512, then shared memory
        ids, dist, cnt, st = idx.search_adc_rerank_batch(q, k, ef, stats=True)
        o_ids, o_dist, o_cnt, o_st = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef,
                                                                 threads=8, stats=True)
        assert np.array_equal(cnt, o_cnt) and np.array_equal(ids, o_ids)
        assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
        for f in ("n_hop", "n_edge", "n_adc", "n_rerank"):
            assert np.array_equal(getattr(st, f), o_st[f]), f
        ids2, dist2, cnt2 = idx.search_adc_rerank_batch(q, k, ef)
        assert np.array_equal(cnt2, o_cnt) and np.array_equal(ids2, o_ids), (k, ef)
        assert np.array_equal(dist2.view(np.uint32), o_dist.view(np.uint32))


@pytest.mark.parametrize("m,ksub", [(32, 64), (8, 256)])
def test_adc_rerank_limit_matches_oracle(gpu_lib, orc, m, ksub):
    """isl_index_set_rerank_limit: the traversal keeps ef survivors, only the max(limit, k) with the best table
    distance (ties by id) get an exact distance.  Same definition in the oracle; register / shared-memory R,
    with and without statistics; limit 0 restores the full rerank."""
    from islands_b200 import LeannIndex, PQConfig, ProductQuantizer

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 3000, 96, seed=61)
    cb = orc.pq_train(1, v[:1000], m, ksub, 3, 7)
    codes = orc.pq_encode(1, cb, v)
    pq = ProductQuantizer(96, PQConfig(m, ksub, 3, 7))
    pq.set_codebooks(cb)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    idx.attach_pq(pq, codes)
    q = np.concatenate([uniform(np.random.RandomState(65), 90, 96), v[:10]])
    for k, ef, limit in [(10, 100, 32), (10, 200, 10), (25, 200, 5), (10, 300, 40), (10, 64, 500)]:
        idx.set_rerank_limit(limit)
        o_ids, o_dist, o_cnt, o_st = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef,
                                                                 threads=8, stats=True, rerank_limit=limit)
        ids, dist, cnt, st = idx.search_adc_rerank_batch(q, k, ef, stats=True)
        ids2, dist2, cnt2 = idx.search_adc_rerank_batch(q, k, ef)
        for a_ids, a_dist, a_cnt in ((ids, dist, cnt), (ids2, dist2, cnt2)):
            assert np.array_equal(a_cnt, o_cnt) and np.array_equal(a_ids, o_ids), (k, ef, limit)
            assert np.array_equal(a_dist.view(np.uint32), o_dist.view(np.uint32))
        for f in ("n_hop", "n_edge", "n_dist", "n_adc", "n_rerank"):
            assert np.array_equal(getattr(st, f), o_st[f]), f
        assert int(st.n_rerank.max()) <= max(limit, k)
    idx.set_rerank_limit(0)
    ids, dist, cnt = idx.search_adc_rerank_batch(q, 10, 100)
    o_ids, o_dist, _ = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, 10, 100, threads=8)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))


@pytest.mark.parametrize("n,m0,ef_list", [(2000, 100, [1, 10, 70, 256]), (40, 24, [1, 5, 64]), (3, 24, [1, 4]), (1, 24, [1, 3])])
def test_adc_traversal_edge_shapes(gpu_lib, orc, n, m0, ef_list):
    """Shapes around the vectorised hop of the lean traversal: lists longer than one 64-position pass (m0 = 100),
    graphs smaller than a pass / than k, a single node, ef = 1 and ef < k — with statistics (visited bitset) and
    without (bitset-free), both against the oracle."""
    from islands_b200 import LeannIndex, PQConfig, ProductQuantizer

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, n, 64, seed=81, m=m0 // 2, m0=m0, ef_construction=max(m0 + 8, 64))
    m, ksub = 16, min(32, n)
    cb = orc.pq_train(1, v, m, ksub, 3, 7)
    codes = orc.pq_encode(1, cb, v)
    pq = ProductQuantizer(64, PQConfig(m, ksub, 3, 7))
    pq.set_codebooks(cb)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    idx.attach_pq(pq, codes)
    q = np.concatenate([uniform(np.random.RandomState(82), 40, 64), v[:min(n, 8)]])
    for ef in ef_list:
        for k in (1, 10):
            o_ids, o_dist, o_cnt, o_st = orc.leann_search_adc_rerank(cfg._s, v, off, nbrs, entry, cb, codes, q, k, ef,
                                                                     threads=4, stats=True)
            ids, dist, cnt, st = idx.search_adc_rerank_batch(q, k, ef, stats=True)
            ids2, dist2, cnt2 = idx.search_adc_rerank_batch(q, k, ef)
            for a_ids, a_dist, a_cnt in ((ids, dist, cnt), (ids2, dist2, cnt2)):
                assert np.array_equal(a_cnt, o_cnt), (n, ef, k)
                assert np.array_equal(a_ids, o_ids), (n, ef, k)
                assert np.array_equal(a_dist.view(np.uint32), o_dist.view(np.uint32))
            for f in ("n_hop", "n_edge", "n_adc", "n_rerank"):
                assert np.array_equal(getattr(st, f), o_st[f]), (f, n, ef, k)
