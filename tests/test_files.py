"""Flat index files of docs/leann-specification.md:967-1027 (`.hnsw`, `.pq`, `.codes`): known-answer bytes built with
`struct` from the spec's field lists, round trips through memory maps, refusal of damaged files, and an oracle search
over a graph that went through a file.  Host-side format code: no GPU."""
import os
import struct

import numpy as np
import pytest

from conftest import oracle_graph, uniform
from islands_b200 import CsrGraph, LeannConfig
from islands_b200.core import SerializationError
from islands_b200.files import (SOURCE_REF_DTYPE, GraphFile, graph_file_from_csr, graph_file_from_hnsw, hubs_by_in_degree,
                                layer_from_export, read_codebook_file,
                                read_codes_file, read_graph_file, write_codebook_file, write_codes_file, write_graph_file)
from islands_b200.storage import DeserializationError


def small_graph():
    # 4 nodes; layer 0: 0->{1,2} 1->{0} 2->{0,3} 3->{2}; layer 1: only nodes 0 and 2 are present
    l0 = (np.array([0, 2, 3, 5, 6]), np.array([1, 2, 0, 0, 3, 2]))
    l1 = (np.array([0, 1, 1, 2, 2]), np.array([2, 0]))
    refs = np.zeros(4, SOURCE_REF_DTYPE)
    refs["source"] = [7, 7, 8, 9]
    refs["chunk_start"] = [0, 64, 0, 0]
    refs["chunk_end"] = [64, 128, 30, 12]
    return GraphFile(num_nodes=4, entry_point=2, metric=0, dimension=768, m=30, ef_construction=128,
                     layers=[l0, l1], hub_ids=np.array([0, 2]), source_refs=refs)


def test_graph_file_known_answer_bytes(tmp_path):
    """Byte image assembled by hand from the spec's field list (:971-995)."""
    path = tmp_path / "g.hnsw"
    written = write_graph_file(path, small_graph())
    data = path.read_bytes()
    assert written == len(data)
    header = (b"HNSW" + struct.pack("<I", 1) + struct.pack("<I", 4) + struct.pack("<B", 2) + struct.pack("<I", 2)
              + struct.pack("<B", 1)  # Cosine is 1 in the file (0 = L2)
              + struct.pack("<I", 768) + struct.pack("<H", 30) + struct.pack("<H", 128))
    assert len(header) == 26
    header += bytes(64 - len(header))
    l0_edges_at = 64 + 16 + 4 * 5
    layer0 = (struct.pack("<B3xIQ", 0, 4, l0_edges_at) + struct.pack("<5I", 0, 2, 3, 5, 6)
              + struct.pack("<6I", 1, 2, 0, 0, 3, 2))
    l1_at = l0_edges_at + 4 * 6
    layer1 = (struct.pack("<B3xIQ", 1, 4, l1_at + 16 + 4 * 5) + struct.pack("<5I", 0, 1, 1, 2, 2)
              + struct.pack("<2I", 2, 0))
    meta = struct.pack("<I", 2) + struct.pack("<2I", 0, 2)
    meta += struct.pack("<12I", 7, 0, 64, 7, 64, 128, 8, 0, 30, 9, 0, 12)
    assert data == header + layer0 + layer1 + meta
    # every array section starts on a 4-byte boundary (what makes the file mappable as typed arrays)
    assert l0_edges_at % 4 == 0 and l1_at % 4 == 0 and (len(header + layer0 + layer1) + 4) % 4 == 0


@pytest.mark.parametrize("mmap", [True, False])
def test_graph_file_round_trip(tmp_path, mmap):
    g = small_graph()
    path = tmp_path / "sub" / "g.hnsw"  # parent directories are created, like FileSystemStorage::save
    write_graph_file(path, g)
    r = read_graph_file(path, mmap=mmap)
    assert (r.num_nodes, r.entry_point, r.metric, r.dimension, r.m, r.ef_construction, r.num_layers) == (4, 2, 0, 768, 30, 128, 2)
    for (a, b), (c, d) in zip(g.layers, r.layers):
        assert np.array_equal(a, c) and np.array_equal(b, d)
        assert isinstance(c, np.memmap) == mmap and c.dtype == np.dtype("<u4")
    assert np.array_equal(r.hub_ids, [0, 2])
    assert np.array_equal(r.source_refs, g.source_refs)
    off, nb = r.to_csr()
    assert off.dtype == np.uint64 and nb.dtype == np.uint64 and off.tolist() == [0, 2, 3, 5, 6]
    assert r.levels().tolist() == [1, 0, 1, 0]
    if mmap:
        with pytest.raises(ValueError):
            r.layers[0][1][0] = 3  # the map is read-only


def test_graph_file_empty_graph_and_default_refs(tmp_path):
    path = tmp_path / "e.hnsw"
    write_graph_file(path, GraphFile(0, None, 1, 16, 8, 32, [(np.zeros(1, np.uint64), np.zeros(0, np.uint64))]))
    assert os.path.getsize(path) == 64 + 16 + 4 + 4  # header, layer header, row_ptr[1], num_hubs
    r = read_graph_file(path)
    assert r.num_nodes == 0 and r.entry_point is None and r.metric == 1 and r.layers[0][1].size == 0
    # without source_refs node i recomputes from token row i
    write_graph_file(path, GraphFile(3, 0, 0, 4, 2, 8, [(np.array([0, 1, 2, 3]), np.array([1, 2, 0]))]))
    assert read_graph_file(path).source_refs["source"].tolist() == [0, 1, 2]


def test_graph_file_refuses_what_it_cannot_represent(tmp_path):
    path = tmp_path / "x.hnsw"
    ok = small_graph()
    for change in (dict(entry_point=4), dict(m=70000), dict(metric=9), dict(hub_ids=np.array([4])),
                   dict(layers=[]), dict(layers=[(np.array([0, 2, 3, 5, 7]), ok.layers[0][1])]),
                   dict(layers=[(np.array([0, 3, 2, 5, 6]), ok.layers[0][1])]),
                   dict(layers=[(ok.layers[0][0], np.array([1, 2, 0, 0, 9, 2]))]),
                   dict(layers=[(ok.layers[0][0], np.array([1, 2, 0, 0, 2 ** 32, 2]))]),
                   dict(source_refs=np.zeros(3, SOURCE_REF_DTYPE))):
        g = small_graph()
        for k, v in change.items():
            setattr(g, k, v)
        with pytest.raises(SerializationError):
            write_graph_file(path, g)


def test_graph_file_damaged(tmp_path):
    path = tmp_path / "g.hnsw"
    write_graph_file(path, small_graph())
    good = path.read_bytes()
    bad = tmp_path / "bad.hnsw"

    def expect_error(data):
        bad.write_bytes(data)
        with pytest.raises(DeserializationError):
            read_graph_file(bad)

    for cut in (0, 10, 63, 70, 64 + 16 + 8, len(good) - 40, len(good) - 1):
        expect_error(good[:cut])
    expect_error(b"HNSX" + good[4:])
    expect_error(good[:4] + struct.pack("<I", 2) + good[8:])          # version
    expect_error(good[:13] + struct.pack("<I", 4) + good[17:])         # entry point == num_nodes
    expect_error(good[:17] + b"\x07" + good[18:])                      # metric code
    expect_error(good[:64] + b"\x01" + good[65:])                      # layer id
    expect_error(good[:72] + struct.pack("<Q", 0) + good[80:])         # edges_start
    # a DeserializationError is a SerializationError is a CoreError, like the reference's variants (error.rs:9-62)
    assert issubclass(DeserializationError, SerializationError)


def test_graph_file_from_csr_graph(tmp_path):
    csr = CsrGraph()
    csr.add_node([1, 2], 0)
    csr.add_node([0], 2)
    csr.add_node([0, 1], 0)
    cfg = LeannConfig(m=4, m0=8, ef_construction=16, metric=1)
    path = tmp_path / "c.hnsw"
    write_graph_file(path, graph_file_from_csr(csr, cfg, 32, hub_ids=hubs_by_in_degree(csr.node_offsets, csr.neighbors, 0.34)))
    r = read_graph_file(path)
    assert (r.num_nodes, r.entry_point, r.metric, r.dimension, r.m, r.ef_construction) == (3, 1, 1, 32, 4, 16)
    off, nb = r.to_csr()
    assert np.array_equal(off, csr.node_offsets) and np.array_equal(nb, csr.neighbors)
    assert r.hub_ids.tolist() == [0, 1]  # in-degrees 2, 2, 1: ceil(0.34 * 3) = 2 hubs, ties to the smaller id


def test_hnsw_layers_from_export(tmp_path):
    """`HnswGraph.export_layer` shape (degrees with -1 for absent nodes, padded neighbour rows) -> file layers, driven
    here by a stand-in object with the handle's interface (the real handle needs a device)."""
    pad = 0xFFFFFFFFFFFFFFFF
    deg0, nb0 = np.array([2, 1, 2, 1]), np.array([[1, 2, pad], [0, pad, pad], [0, 3, pad], [2, pad, pad]], np.uint64)
    deg1, nb1 = np.array([1, -1, 1, -1]), np.array([[2, pad], [pad, pad], [0, pad], [pad, pad]], np.uint64)
    rp, ed = layer_from_export(deg1, nb1)
    assert rp.tolist() == [0, 1, 1, 2, 2] and ed.tolist() == [2, 0]
    assert layer_from_export(np.zeros(0, np.int64), np.zeros((0, 4), np.uint64))[0].tolist() == [0]

    class Handle:  # the part of islands_b200.HnswGraph the converter uses
        config = type("Cfg", (), dict(m=2, m0=3, ef_construction=16, metric=1))()
        entry_point, max_level = 2, 1

        def __len__(self):
            return 4

        def dimension(self):
            return 8

        def export_layer(self, layer):
            return (deg0, nb0) if layer == 0 else (deg1, nb1)

    path = tmp_path / "h.hnsw"
    write_graph_file(path, graph_file_from_hnsw(Handle(), hub_ids=[0]))
    r = read_graph_file(path)
    assert (r.num_nodes, r.num_layers, r.entry_point, r.metric, r.dimension, r.m, r.ef_construction) == (4, 2, 2, 1, 8, 2, 16)
    assert r.layers[0][0].tolist() == [0, 2, 3, 5, 6] and r.layers[0][1].tolist() == [1, 2, 0, 0, 3, 2]
    assert r.layers[1][0].tolist() == [0, 1, 1, 2, 2] and r.layers[1][1].tolist() == [2, 0]
    assert r.levels().tolist() == [1, 0, 1, 0] and r.hub_ids.tolist() == [0]


def test_hubs_by_in_degree():
    row_ptr = np.array([0, 2, 4, 6, 7, 7])
    edges = np.array([3, 4, 3, 4, 3, 0, 0])  # in-degree: node 0: 2, 3: 3, 4: 2, others 0
    assert hubs_by_in_degree(row_ptr, edges, 0.2).tolist() == [3]
    assert hubs_by_in_degree(row_ptr, edges, 0.4).tolist() == [0, 3]
    assert hubs_by_in_degree(row_ptr, edges, 0.6).tolist() == [0, 3, 4]
    assert hubs_by_in_degree(row_ptr, edges, 5.0).tolist() == [0, 1, 2, 3, 4]
    assert hubs_by_in_degree(row_ptr, edges, 0.0).size == 0
    assert hubs_by_in_degree(np.array([0]), np.zeros(0, np.int64), 0.5).size == 0


def test_search_over_a_graph_read_from_a_file(tmp_path, orc):
    """The oracle's search (leann.rs:868-988) over the mapped arrays of a written file returns what it returns over the
    arrays the graph was built into."""
    cfg, v, levels, off, nbrs, entry = oracle_graph(orc, 600, 24, seed=11)
    csr = CsrGraph()
    csr.node_offsets, csr.neighbors, csr.entry_point, csr.num_nodes = off, nbrs, entry, 600
    path = tmp_path / "o.hnsw"
    write_graph_file(path, graph_file_from_csr(csr, cfg, 24))
    r = read_graph_file(path)
    f_off, f_nb = r.to_csr()
    q = uniform(np.random.RandomState(12), 40, 24)
    a = orc.leann_search(cfg._s, v, off, nbrs, entry, q, 10, 48)
    b = orc.leann_search(cfg._s, v, f_off, f_nb, r.entry_point, q, 10, 48)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[2], b[2])


def test_codebook_file(tmp_path):
    rng = np.random.RandomState(3)
    cb = uniform(rng, 4 * 16, 6).reshape(4, 16, 6)
    path = tmp_path / "q.pq"
    written = write_codebook_file(path, cb)
    data = path.read_bytes()
    assert written == len(data) == 32 + 4 * 4 * 16 * 6 + 4 * 4 * 16
    assert data[:32] == b"PQCB" + struct.pack("<IHHH", 1, 4, 16, 6) + bytes(18)
    assert data[32:32 + cb.nbytes] == cb.astype("<f4").tobytes()
    r_cb, r_norms = read_codebook_file(path)
    assert isinstance(r_cb, np.memmap) and np.array_equal(r_cb, cb) and r_norms.shape == (4, 16)
    # norms: strict left-to-right f32 fold, checked element by element against a scalar loop
    for s, c in ((0, 0), (3, 15), (2, 7)):
        acc = np.float32(0)
        for x in cb[s, c]:
            acc = np.float32(acc + np.float32(x * x))
        assert r_norms[s, c].view(np.uint32) == acc.view(np.uint32)
    write_codebook_file(path, cb, with_norms=False)
    r_cb, r_norms = read_codebook_file(path, mmap=False)
    assert r_norms is None and np.array_equal(r_cb, cb)
    good = path.read_bytes()
    for data in (good[:20], good[:-4], good + b"\0" * 8, b"PQCX" + good[4:], good[:4] + struct.pack("<I", 3) + good[8:],
                 good[:8] + struct.pack("<H", 0) + good[10:]):
        path.write_bytes(data)
        with pytest.raises(DeserializationError):
            read_codebook_file(path)
    with pytest.raises(SerializationError):
        write_codebook_file(path, np.zeros((4, 16), np.float32))


def test_codes_file(tmp_path):
    rng = np.random.RandomState(4)
    codes = rng.randint(0, 256, size=(1000, 32)).astype(np.uint16)  # what ProductQuantizer.encode returns
    path = tmp_path / "c.codes"
    assert write_codes_file(path, codes) == 16 + 1000 * 32
    data = path.read_bytes()
    assert data[:16] == b"PQCD" + struct.pack("<IIB", 1, 1000, 32) + bytes(3)
    assert data[16:] == codes.astype(np.uint8).tobytes()
    r = read_codes_file(path)
    assert isinstance(r, np.memmap) and r.dtype == np.uint8 and np.array_equal(r, codes)
    assert np.array_equal(read_codes_file(path, mmap=False).astype(np.uint16), codes)
    write_codes_file(path, np.zeros((0, 8), np.uint8))
    assert read_codes_file(path).shape == (0, 8)
    with pytest.raises(SerializationError):
        write_codes_file(path, np.full((2, 8), 256, np.uint16))  # two-byte codes (num_centroids > 256)
    with pytest.raises(SerializationError):
        write_codes_file(path, np.zeros(8, np.uint8))
    for data in (data[:10], data[:-1], data + b"\0", b"PQCB" + data[4:], data[:4] + struct.pack("<I", 0) + data[8:],
                 data[:12] + b"\0" + data[13:]):
        path.write_bytes(data)
        with pytest.raises(DeserializationError):
            read_codes_file(path)
