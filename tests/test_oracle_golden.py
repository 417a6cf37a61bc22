"""The oracle against its own frozen outputs (tests/golden/search_golden.npz, written by make_search_golden.py): graphs,
search results, counters, quantizer tables and codes, HNSW lists — every array bit for bit.  The reference holds no
golden results for these; this fixture holds the doubly-read state of the oracle still, so that the parity
target of the GPU tests cannot drift unnoticed.  CPU."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_reproduces_its_frozen_outputs(orc):
    spec = importlib.util.spec_from_file_location("make_search_golden", os.path.join(ROOT, "tests", "golden", "make_search_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.compute()
    golden = np.load(os.path.join(ROOT, "tests", "golden", "search_golden.npz"))
    assert sorted(now) == sorted(golden.files) and len(golden.files) >= 40
    for name in golden.files:
        assert now[name].dtype == golden[name].dtype and np.array_equal(now[name], golden[name]), name
    # a few values spelled out, so that the fixture is not only compared with itself
    assert golden["m0_entry"].tolist() == [int(golden["m0_entry"][0]), int(golden["m0_entry"][1])] and golden["m0_off"][0] == 0
    assert int(golden["m0_off"][-1]) == golden["m0_nbrs"].size and (np.diff(golden["m0_off"].astype(np.int64)) <= 60).all()
    for metric in range(4):
        dist = golden[f"m{metric}_dist"].view(np.float32)
        assert (golden[f"m{metric}_cnt"] == 10).all() and (np.diff(dist, axis=1) >= 0).all()
        assert (golden[f"m{metric}_ids"] < 500).all()
    assert (golden["pq_codes"] < 16).all() and golden["pq_codebooks"].shape == (4, 16, 8)


def test_second_reading_reproduces_the_frozen_search_results():
    """The pure-Python reading of leann.rs:899-988 (tests/test_oracle_second_reading.py) on the frozen graphs — no oracle
    in the loop: ids, distance bits and the work counter of the fixture come out of the second restatement too."""
    from islands_b200 import LeannConfig
    from test_oracle_second_reading import search

    golden = np.load(os.path.join(ROOT, "tests", "golden", "search_golden.npz"))
    rng = np.random.RandomState(2024)
    v = (rng.rand(500, 32).astype(np.float32) * 2 - 1).astype(np.float32)
    v[450:] = v[:50]
    q = (rng.rand(16, 32).astype(np.float32) * 2 - 1).astype(np.float32)
    for metric in (0, 1):
        cfg = LeannConfig(metric=metric)
        off, nbrs, entry = golden[f"m{metric}_off"], golden[f"m{metric}_nbrs"], int(golden[f"m{metric}_entry"][0])
        for qi in range(0, 16, 3):
            mine, computed = search(cfg, v, off, nbrs, entry, q[qi], 10, 48)
            assert [i for i, _ in mine] == golden[f"m{metric}_ids"][qi].tolist()
            assert [np.float32(d).view(np.uint32) for _, d in mine] == golden[f"m{metric}_dist"][qi].tolist()
            assert computed == int(golden[f"m{metric}_stats"][2, qi])


def _inputs():
    rng = np.random.RandomState(2024)
    v = (rng.rand(500, 32).astype(np.float32) * 2 - 1).astype(np.float32)
    v[450:] = v[:50]
    q = (rng.rand(16, 32).astype(np.float32) * 2 - 1).astype(np.float32)
    return v, q


def _check_search_against_golden(search_fn):
    """search_fn(cfg, vectors, offsets, neighbors, levels, entry, queries, k, ef) -> ids, dist, count, (n_hop, n_edge, n_dist)."""
    from islands_b200 import LeannConfig

    golden = np.load(os.path.join(ROOT, "tests", "golden", "search_golden.npz"))
    v, q = _inputs()
    for metric in range(4):
        cfg = LeannConfig(metric=metric)
        ids, dist, cnt, counters = search_fn(cfg, v, golden[f"m{metric}_off"], golden[f"m{metric}_nbrs"], golden[f"m{metric}_levels"],
                                             int(golden[f"m{metric}_entry"][0]), q, 10, 48)
        assert np.array_equal(cnt, golden[f"m{metric}_cnt"]), metric
        assert np.array_equal(ids, golden[f"m{metric}_ids"]), metric
        assert np.array_equal(np.ascontiguousarray(dist, np.float32).view(np.uint32), golden[f"m{metric}_dist"]), metric
        assert np.array_equal(np.stack([np.asarray(c, np.uint64) for c in counters]), golden[f"m{metric}_stats"]), metric


def test_oracle_search_against_the_frozen_results(orc):
    def oracle(cfg, v, off, nbrs, levels, entry, q, k, ef):
        ids, dist, cnt, st = orc.leann_search(cfg._s, v, off, nbrs, entry, q, k, ef, stats=True)
        return ids, dist, cnt, (st["n_hop"], st["n_edge"], st["n_dist"])

    _check_search_against_golden(oracle)



@pytest.mark.gpu
def test_gpu_search_against_the_frozen_results(gpu_lib):
    """The CUDA search on the frozen graphs against the frozen results — the same comparison the oracle passes above, with
    no oracle in the loop: ids, distance bits, counts and the three traversal counters for the four metrics."""
    from islands_b200 import LeannIndex

    def gpu(cfg, v, off, nbrs, levels, entry, q, k, ef):
        idx = LeannIndex.from_csr(cfg, v, off, nbrs, levels, entry)
        ids, dist, cnt, st = idx.search_batch(q, k, ef, stats=True)
        idx.free()
        return ids, dist, cnt, (st.n_hop, st.n_edge, st.n_dist)

    _check_search_against_golden(gpu)
