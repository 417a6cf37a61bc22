"""K7: GPU graph construction vs the oracle's restatement of LeannIndex::build
(leann.rs:560-631): CSR arrays bit-exact for the sequential loop (batch=1) and for the batched
round model (same rounds on both sides)."""
import os

import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu


def _build_both(orc, v, cfg, levels, batch):
    from islands_b200 import InMemoryEmbeddingProvider, LeannIndex

    idx = LeannIndex(cfg)
    idx.build(InMemoryEmbeddingProvider(v), v.shape[0], levels=levels, batch=batch)
    g = idx.graph
    off, nbrs, entry, maxl = orc.leann_build(cfg._s, v, levels, batch=batch, threads=os.cpu_count() or 1)
    return idx, g, (off, nbrs, entry, maxl)


def _assert_same(g, ref):
    off, nbrs, entry, maxl = ref
    assert np.array_equal(g.node_offsets, off)
    assert np.array_equal(g.neighbors, nbrs)
    assert g.entry_point == entry and g.max_level == maxl
    assert np.array_equal(g.degree_counts, np.diff(off.astype(np.int64)).astype(np.uint64))


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_sequential_build_bit_exact(gpu_lib, orc, metric):
    from islands_b200 import LeannConfig

    v = uniform(np.random.RandomState(21 + metric), 700, 24)
    cfg = LeannConfig(metric=metric)
    levels = orc.draw_levels(4, 700, cfg.ml, cfg.max_layers)
    idx, g, ref = _build_both(orc, v, cfg, levels, 1)
    _assert_same(g, ref)


@pytest.mark.parametrize("batch", [7, 64, 1024])
@pytest.mark.parametrize("hub", [1, 0])
def test_batched_build_bit_exact(gpu_lib, orc, batch, hub):
    from islands_b200 import LeannConfig

    v = uniform(np.random.RandomState(31), 3000, 48)
    cfg = LeannConfig(high_degree_pruning=hub)
    levels = orc.draw_levels(5, 3000, cfg.ml, cfg.max_layers)
    idx, g, ref = _build_both(orc, v, cfg, levels, batch)
    _assert_same(g, ref)
    # the built index searches like the oracle on the same graph
    q = uniform(np.random.RandomState(32), 64, 48)
    ids, dist, cnt = idx.search_batch(q, 10, 64)
    o = orc.leann_search(cfg._s, v, ref[0], ref[1], ref[2], q, 10, 64)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist.view(np.uint32), o[1].view(np.uint32))


def test_build_small_m0_and_duplicates(gpu_lib, orc):
    from islands_b200 import LeannConfig

    v = uniform(np.random.RandomState(41), 900, 16)
    v[600:] = v[:300]  # duplicate vectors: distance ties in search, selection and pruning
    cfg = LeannConfig(m=4, m0=8, ef_construction=20)
    levels = orc.draw_levels(6, 900, cfg.ml, cfg.max_layers)
    for batch in (1, 32):
        idx, g, ref = _build_both(orc, v, cfg, levels, batch)
        _assert_same(g, ref)


def test_build_properties(gpu_lib, orc):
    """leann.rs:1269-1287, 1468-1511: len, dimension, valid ids, degree cap, default levels."""
    from islands_b200 import InMemoryEmbeddingProvider, LeannConfig, LeannIndex

    v = uniform(np.random.RandomState(51), 500, 32)
    idx = LeannIndex(LeannConfig())
    idx.build(InMemoryEmbeddingProvider(v), 500, seed=9, batch=16)  # levels drawn from the seed
    assert len(idx) == 500 and idx.dimension() == 32 and not idx.is_empty()
    g = idx.graph
    lv = orc.draw_levels(9, 500, idx.config.ml, idx.config.max_layers)
    assert np.array_equal(g.levels, lv)
    assert g.degree_counts.max() <= 60 and (g.neighbors < 500).all() and g.storage_bytes() == idx.storage_bytes()
    res = idx.search(v[0], 5)
    assert res[0][0] == 0 and res[0][1] < 0.01
    empty = LeannIndex(LeannConfig())
    empty.build(InMemoryEmbeddingProvider(v), 0)
    assert len(empty) == 0 and empty.search(v[0], 3) == []
    one = LeannIndex(LeannConfig())
    one.build(InMemoryEmbeddingProvider(v), 1)
    assert one.search(v[3], 3)[0][0] == 0 and len(one.search(v[3], 3)) == 1
