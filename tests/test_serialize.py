"""to_bytes / from_bytes: the product's bytes equal the independent CPU restatement
of the reference's bincode layout, and from_bytes(to_bytes(x)) behaves like x.  Parity unpinned by the
reference (no Rust toolchain here): the layout is pinned by hand-computed known answers only."""
import struct

import numpy as np
import pytest

from conftest import oracle_graph, uniform


def test_bincode_oracle_known_answers():
    """CPU: hand-computed bytes of the default LeannConfig and of a two-node CSR index."""
    from islands_b200 import LeannConfig
    from oracle import bincode_oracle as bo

    # bincode 1.x's own README example: `World(vec![Entity{x: 0.0, y: 4.0}, Entity{x: 10.0, y: 20.5}])` encodes to
    # "8 bytes for the length of the vector, 4 bytes per float" = 24 bytes: fixed-width little-endian, no varints
    world = bo._u64(2) + b"".join(bo._f32(x) for x in (0.0, 4.0, 10.0, 20.5))
    assert len(world) == 8 + 4 * 4 and world[:8] == bytes([2, 0, 0, 0, 0, 0, 0, 0])
    c = LeannConfig()
    b = bo.leann_config(c)
    assert len(b) == 8 * 3 + 8 + 8 + 4 + 8 + 8 + 4 + 4 + 1 + 4 + 1 + 1 == 75
    assert b[:8] == (30).to_bytes(8, "little") and b[8:16] == (60).to_bytes(8, "little")
    assert struct.unpack("<d", b[24:32])[0] == c.ml
    assert b[40:44] == b"\x00\x00\x00\x00"            # DistanceMetric::Cosine = variant 0
    assert b[-7:] == b"\x01" + struct.pack("<f", 0.02) + b"\x01\x01"   # high_degree_pruning, hub_percentile, compact, recompute
    idx = bo.leann_index(c, [0, 1, 2], [1, 0], [0, 0], 0, 0, 4)
    tail = idx[75:]
    expect = (struct.pack("<Q", 3) + struct.pack("<QQQ", 0, 1, 2) + struct.pack("<Q", 2) + struct.pack("<QQ", 1, 0)
              + struct.pack("<Q", 2) + struct.pack("<QQ", 0, 0) + b"\x01" + struct.pack("<Q", 0) + struct.pack("<Q", 0)
              + struct.pack("<Q", 2) + struct.pack("<Q", 2) + struct.pack("<QQ", 1, 1) + b"\x01" + struct.pack("<Q", 4))
    assert tail == expect
    assert bo.leann_index(c, [0], [], [], None, 0, None)[75:] == (struct.pack("<Q", 1) + struct.pack("<Q", 0) + struct.pack("<Q", 0)
                                                                   + struct.pack("<Q", 0) + b"\x00" + struct.pack("<QQ", 0, 0)
                                                                   + struct.pack("<Q", 0) + b"\x00")


@pytest.mark.gpu
def test_leann_index_bytes_match_oracle_and_round_trip(gpu_lib, orc):
    """to_bytes / from_bytes (leann.rs:1059-1066) and the reference's own round-trip tests (leann.rs:1346-1384: the
    restored index has the same length and answers searches), plus the bytes themselves against the bincode restatement."""
    from islands_b200 import LeannIndex, SerializationError
    from oracle import bincode_oracle as bo

    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 1000, 32, seed=11)
    idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
    data = idx.to_bytes()
    assert data == bo.leann_index(cfg, off, nbrs, levels, entry, int(levels[entry]), 32)
    back = LeannIndex.from_bytes(data, v)
    assert back.config.m == cfg.m and back.config.ef_construction == cfg.ef_construction
    g0, g1 = idx.graph, back.graph
    assert np.array_equal(g0.node_offsets, g1.node_offsets) and np.array_equal(g0.neighbors, g1.neighbors)
    assert np.array_equal(g0.levels, g1.levels) and g0.entry_point == g1.entry_point and g0.max_level == g1.max_level
    q = uniform(np.random.RandomState(3), 50, 32)
    a, b = idx.search_batch(q, 10, 64), back.search_batch(q, 10, 64)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert back.to_bytes() == data
    for bad in (data[:-1], data + b"\x00", data[:40]):
        with pytest.raises(SerializationError):
            LeannIndex.from_bytes(bad, v)
    empty = LeannIndex(cfg)
    empty.build(v, 0)
    assert empty.to_bytes() == bo.leann_index(cfg, [0], [], [], None, 0, None)


@pytest.mark.gpu
def test_pq_bytes_match_oracle_and_round_trip(gpu_lib):
    from islands_b200 import PQConfig, ProductQuantizer, SerializationError
    from oracle import bincode_oracle as bo

    rng = np.random.RandomState(0)
    v = uniform(rng, 600, 32)
    pq = ProductQuantizer(32, PQConfig(4, 16, 5, 7))
    untrained = pq.to_bytes()
    assert untrained == bo.product_quantizer(4, 16, 5, 7, None, 32, 1, False)
    pq.train(v)
    data = pq.to_bytes()
    assert data == bo.product_quantizer(4, 16, 5, 7, pq.codebooks(), 32, 1, True)
    back = ProductQuantizer.from_bytes(data)
    assert back.is_trained() and back.dimension == 32 and back.config.num_centroids == 16 and back.config.seed == 7
    assert np.array_equal(back.encode(v), pq.encode(v))
    assert back.to_bytes() == data
    assert not ProductQuantizer.from_bytes(untrained).is_trained()
    with pytest.raises(SerializationError):
        ProductQuantizer.from_bytes(data[:-5])


@pytest.mark.gpu
def test_hnsw_bytes_match_oracle_round_trip_and_keep_inserting(gpu_lib, orc):
    from islands_b200 import HnswConfig, HnswGraph, SerializationError
    from oracle import bincode_oracle as bo

    cfg = HnswConfig(m=8, m0=16, ef_construction=32, ml=0.8)
    n, d = 600, 24
    v = uniform(np.random.RandomState(5), n, d)
    lv = orc.draw_levels(9, n, cfg.ml, cfg.max_layers)
    og = orc.Hnsw(cfg._s, d)
    og.insert_batch(v[:400], lv[:400], batch=16, threads=8)
    g = HnswGraph(cfg)
    g.insert_batch(v[:400], lv[:400], batch=16)
    data = g.to_bytes()
    assert data == bo.hnsw_graph(cfg, v[:400], lv[:400], lambda i, layer: og.neighbors(i, layer), og.entry_point(), og.max_level())
    back = HnswGraph.from_bytes(data)
    assert len(back) == 400 and back.entry_point == g.entry_point and back.max_level == g.max_level
    assert back.config.m == 8 and back.config.ef_construction == 32
    assert back.to_bytes() == data
    q = uniform(np.random.RandomState(6), 40, d)
    a, b = g.search_batch(q, 10, 50), back.search_batch(q, 10, 50)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    # the deserialised graph keeps growing exactly like the original (cached edge distances rebuilt)
    og.insert_batch(v[400:], lv[400:], batch=16, threads=8)
    back.insert_batch(v[400:], lv[400:], batch=16)
    for layer in range(int(lv.max()) + 1):
        deg, nb = back.export_layer(layer)
        for i in range(n):
            ref = og.neighbors(i, layer)
            assert (deg[i] == -1) == (ref is None)
            if ref is not None:
                assert np.array_equal(nb[i, :deg[i]], ref), (i, layer)
    assert HnswGraph.from_bytes(HnswGraph(cfg).to_bytes()).is_empty()
    with pytest.raises(SerializationError):
        HnswGraph.from_bytes(data[:-3])


def test_from_bytes_survives_hostile_input():
    """CPU (the parsers are host code): truncated, extended and corrupted images of all three index types — length
    fields set to 2^64 - 1 and the like — come back as a status, never as a crash or a runaway allocation
    (scripts/fuzz_from_bytes.py: each batch in its own process under an address-space limit; 7000 cases clean)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_from_bytes.py"), "--cases", "500", "--seed", "11"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "500 cases, 0 failed batches" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
