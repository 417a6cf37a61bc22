"""K1: batched distances through the C ABI vs the oracle (distance.rs:32-122): bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")


def test_reference_kats_on_gpu(gpu_lib):
    from islands_b200 import DistanceMetric

    for k in json.load(open(GOLDEN))["distance"]:
        got = DistanceMetric(k["metric"]).calculate(k["a"], k["b"])
        if k.get("exact"):
            assert got == k["expect"], k
        else:
            assert abs(got - k["expect"]) < k["tol"], k
    assert abs(DistanceMetric(1).calculate_squared([0, 0], [3, 4]) - 25.0) < 1e-6
    d = DistanceMetric(0).calculate([1, 0], [0, 1])
    assert abs(DistanceMetric(0).calculate_squared([1, 0], [0, 1]) - d * d) < 1e-6
    assert DistanceMetric(0).batch_calculate([1, 0], np.zeros((0, 2), np.float32)).size == 0


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [1, 3, 32, 96, 128, 250, 768, 1024])
def test_batch_bit_exact(gpu_lib, orc, metric, d):
    from islands_b200 import DistanceMetric

    rng = np.random.RandomState(d * 7 + metric)
    rows = uniform(rng, 777, d)
    rows[5] = 0.0  # zero vector: cosine must return exactly 1.0
    q = uniform(rng, 1, d)[0]
    got = DistanceMetric(metric).batch_calculate(q, rows)
    exp = orc.distance_batch(metric, q, rows)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    np.testing.assert_allclose(got, exp, rtol=1e-5)  # the stated bar
    if metric == 0:
        assert got[5] == 1.0
    sq = np.array([DistanceMetric(metric).calculate_squared(q, rows[i]) for i in range(3)], np.float32)
    exp_sq = np.array([orc.distance_squared(metric, q, rows[i]) for i in range(3)], np.float32)
    assert np.array_equal(sq.view(np.uint32), exp_sq.view(np.uint32))


def test_normalize_rows(gpu_lib, orc):
    from islands_b200 import normalize_vector

    rng = np.random.RandomState(3)
    rows = uniform(rng, 50, 70)
    rows[7] = 0.0
    got = normalize_vector(rows)
    exp = np.stack([orc.normalize(r) for r in rows])
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert (got[7] == 0).all()
    assert abs(np.linalg.norm(normalize_vector(np.array([3.0, 4.0]))) - 1.0) < 1e-6


def test_merge_topk(gpu_lib, orc):
    from islands_b200 import merge_topk

    rng = np.random.RandomState(4)
    parts, nq, k = 5, 300, 10
    dist = np.sort(rng.rand(parts, nq, k).astype(np.float32), axis=2)
    dist[:, :, 3] = dist[:, :, 2]  # ties inside and across parts
    ids = rng.permutation(parts * nq * k).astype(np.uint64).reshape(parts, nq, k)
    ids[2, :, 7:] = 0xFFFFFFFFFFFFFFFF  # a short shard
    dist[2, :, 7:] = np.inf
    got = merge_topk(ids, dist, k)
    exp = orc.merge_topk(ids, dist, k)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)
    # size-independent property: merging a single part only reorders exact ties by id
    one = merge_topk(ids[:1], dist[:1], k)
    assert np.array_equal(one[1], dist[0])
    order = np.lexsort((ids[0], dist[0]), axis=1)
    assert np.array_equal(one[0], np.take_along_axis(ids[0], order, axis=1))
