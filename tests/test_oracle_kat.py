"""Pins the CPU oracle against every known-answer value the reference's own tests hold for this
path.  Runs without a GPU."""
import json
import os

import numpy as np
import pytest

from conftest import uniform

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")
COS, L2, DOT, L1 = 0, 1, 2, 3


def test_distance_kats(orc):
    kats = json.load(open(GOLDEN))["distance"]
    for k in kats:
        got = orc.distance(k["metric"], k["a"], k["b"])
        if k.get("exact"):
            assert got == k["expect"], k
        else:
            assert abs(got - k["expect"]) < k["tol"], k


def test_distance_squared_kats(orc):
    assert abs(orc.distance_squared(L2, [0, 0], [3, 4]) - 25.0) < 1e-6   # distance.rs:354-363
    d = orc.distance(COS, [1, 0], [0, 1])
    assert abs(orc.distance_squared(COS, [1, 0], [0, 1]) - d * d) < 1e-6  # distance.rs:365-372


def test_batch_and_normalize(orc):
    d = orc.distance_batch(COS, [1, 0], [[1, 0], [0, 1], [-1, 0]])      # distance.rs:250-261
    assert abs(d[0]) < 1e-6 and abs(d[1] - 1) < 1e-6 and abs(d[2] - 2) < 1e-6
    v = orc.normalize([3.0, 4.0])                                        # distance.rs:231-240
    assert abs(np.sqrt((v * v).sum()) - 1) < 1e-6
    assert (orc.normalize([0.0, 0.0, 0.0]) == 0).all()                   # distance.rs:242-248


def test_distance_properties(orc):
    """distance.rs:264-328 proptests restated on seeded inputs."""
    rng = np.random.RandomState(0)
    for _ in range(200):
        a, b, c = (rng.uniform(-5, 5, 8).astype(np.float32) for _ in range(3))
        assert orc.distance(L2, a, b) >= 0 and orc.distance(L1, a, b) >= 0
        assert abs(orc.distance(L2, a, b) - orc.distance(L2, b, a)) < 1e-5
        assert abs(orc.distance(L2, a, a)) < 1e-6
        assert orc.distance(L2, a, c) <= orc.distance(L2, a, b) + orc.distance(L2, b, c) + 1e-5
        p, q = np.abs(a) + 0.1, np.abs(b) + 0.1
        assert 0 <= orc.distance(COS, p, q) <= 2
    # bitwise symmetry of every metric (relied on by the build's cached edge distances)
    for metric in (COS, L2, DOT, L1):
        for _ in range(200):
            a, b = uniform(rng, 1, 96)[0], uniform(rng, 1, 96)[0]
            assert np.float32(orc.distance(metric, a, b)).tobytes() == np.float32(orc.distance(metric, b, a)).tobytes()


def test_oracle_matches_numpy_sequential_fold(orc):
    """The fold order: a float32 running sum, one element at a time, products rounded separately."""
    rng = np.random.RandomState(1)
    a, b = uniform(rng, 1, 768)[0], uniform(rng, 1, 768)[0]
    dot = np.float32(0)
    na = np.float32(0)
    nb = np.float32(0)
    for x, y in zip(a, b):
        dot = np.float32(dot + np.float32(x * y))
        na = np.float32(na + np.float32(x * x))
        nb = np.float32(nb + np.float32(y * y))
    expect = np.float32(1) - np.float32(dot / np.float32(np.sqrt(np.float32(na * nb))))
    assert np.float32(orc.distance(COS, a, b)).tobytes() == np.float32(expect).tobytes()
    s = np.float32(0)
    for x, y in zip(a, b):
        diff = np.float32(x - y)
        s = np.float32(s + np.float32(diff * diff))
    assert np.float32(orc.distance(L2, a, b)).tobytes() == np.float32(np.sqrt(s)).tobytes()


def test_pq_kats(orc):
    kats = json.load(open(GOLDEN))["pq_find_nearest"]                     # pq.rs:787-809
    cb = np.array(kats["centroids"], np.float32)[None]                   # [m=1][ksub=3][dsub=4]
    for q, expect in kats["queries"]:
        assert orc.pq_encode(L2, cb, np.array([q], np.float32))[0, 0] == expect
    # LUT vs direct ADC agree to 1e-3 (pq.rs:639-669)
    rng = np.random.RandomState(2)
    v = uniform(rng, 300, 32)
    cbs = orc.pq_train(L2, v, 4, 16, 10, 42)
    codes = orc.pq_encode(L2, cbs, v[:10])
    q = uniform(rng, 1, 32)[0]
    t = orc.pq_build_tables(cbs, q)
    assert t.shape == (4, 16)
    np.testing.assert_allclose(orc.pq_table_distance(t, codes), orc.pq_asymmetric_distance(cbs, q, codes), atol=1e-3)
    # decode(encode(v)) has the right dimension and is no worse than a random centroid (pq.rs:573-607)
    dec = orc.pq_decode(cbs, codes)
    assert dec.shape == (10, 32)
    assert (orc.pq_asymmetric_distance(cbs, q, codes) >= 0).all()


def test_level_formula(orc):
    """leann.rs:549-554: floor(-ln(u) * ml) capped at max_layers-1."""
    ml = 1.0 / np.log(30.0)
    lib = orc.lib()
    assert lib.orc_level_from_uniform(0.9, ml, 16) == 0
    assert lib.orc_level_from_uniform(1.0 / 31.0, ml, 16) == 1
    assert lib.orc_level_from_uniform(1e-300, ml, 16) == 15
    lv = orc.draw_levels(5, 100000, ml, 16)
    assert 0.955 < (lv == 0).mean() < 0.975  # P(level 0) = 1 - 1/30


def test_search_properties_on_oracle(orc):
    """Properties the reference's LEANN tests pin (leann.rs:1269-1343, 1388-1433, 1437-1465)."""
    from islands_b200 import LeannConfig

    rng = np.random.RandomState(42)
    v = uniform(rng, 200, 32)
    cfg = LeannConfig.accurate()
    lv = orc.draw_levels(1, 200, cfg.ml, cfg.max_layers)
    off, nb, ep, _ = orc.leann_build(cfg._s, v, lv)
    assert off[-1] == nb.size and (np.diff(off.astype(np.int64)) <= cfg.m0).all()
    ids, dist, cnt = orc.leann_search(cfg._s, v, off, nb, ep, v[:1], 5, 128)
    assert ids[0, 0] == 0 and dist[0, 0] < 0.01
    q = uniform(rng, 20, 32)
    ids, dist, cnt = orc.leann_search(cfg._s, v, off, nb, ep, q, 20, 128)
    assert (cnt == 20).all() and (np.diff(dist, axis=1) >= 0).all() and (ids < 200).all()
    gt = np.array([np.argmin(orc.distance_batch(0, qq, v)) for qq in q])
    assert (ids[:, 0] == gt).mean() >= 0.35
    # frontier pruning still returns results (leann.rs:1437-1465)
    fast = LeannConfig.fast()
    ids, _, cnt = orc.leann_search(fast._s, v, off, nb, ep, q, 5, 32)
    assert (cnt > 0).all()


def test_build_batch1_equals_sequential_reference_loop(orc):
    """batch=1 of the round model is the reference's sequential insert loop by construction;
    larger batches keep the graph valid (degree cap, ids in range, no self loops)."""
    from islands_b200 import LeannConfig

    rng = np.random.RandomState(3)
    v = uniform(rng, 400, 16)
    cfg = LeannConfig()
    lv = orc.draw_levels(2, 400, cfg.ml, cfg.max_layers)
    a = orc.leann_build(cfg._s, v, lv, batch=1)
    b = orc.lib()  # same entry through the non-batched symbol
    off = np.zeros(401, np.uint64)
    nb = np.zeros(400 * 60, np.uint64)
    import ctypes as C
    ne, ep, ml = C.c_uint64(), C.c_int64(), C.c_uint64()
    rc = b.orc_leann_build(C.cast(C.byref(cfg._s), C.c_void_p), v.ctypes.data_as(orc.f32p), 400, 16,
                           lv.ctypes.data_as(orc.u64p), off.ctypes.data_as(orc.u64p), nb.ctypes.data_as(orc.u64p),
                           C.byref(ne), C.byref(ep), C.byref(ml))
    assert rc == 0 and np.array_equal(off, a[0]) and np.array_equal(nb[: ne.value], a[1]) and ep.value == a[2]
    off8, nb8, ep8, _ = orc.leann_build(cfg._s, v, lv, batch=8, threads=4)
    deg = np.diff(off8.astype(np.int64))
    assert deg.max() <= 60 and (nb8 < 400).all()
    src = np.repeat(np.arange(400), deg)
    assert (src != nb8).all()


def test_merge_kat(orc):
    ids = np.array([[[1, 2, 3]], [[10, 11, 0xFFFFFFFFFFFFFFFF]]], np.uint64)
    dist = np.array([[[0.1, 0.5, 0.9]], [[0.05, 0.5, np.inf]]], np.float32)
    oi, od, oc = orc.merge_topk(ids, dist, 3)
    assert oi.tolist() == [[10, 1, 2]] and oc.tolist() == [3]  # tie at 0.5: id 2 < id 11
    assert abs(orc.lib().orc_to_similarity(1.0) - 0.5) < 1e-7
