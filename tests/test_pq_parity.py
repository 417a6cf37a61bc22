"""K3/K4: product quantizer through the C ABI vs the oracle (pq.rs): codes and decoded vectors
bit-exact, tables and ADC distances bit-exact (bar 1e-5 relative)."""
import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu


def _pq(d, m, ksub, cb, metric=1):
    from islands_b200 import PQConfig, ProductQuantizer

    pq = ProductQuantizer(d, PQConfig(m, ksub, 5, 1)).with_metric(metric)
    pq.set_codebooks(cb)
    return pq


@pytest.mark.parametrize("d,m,ksub", [(32, 4, 16), (128, 8, 256), (768, 8, 256), (768, 96, 256), (60, 4, 300), (36, 4, 7)])
def test_encode_decode_tables_adc(gpu_lib, orc, d, m, ksub):
    rng = np.random.RandomState(d + m)
    cb = uniform(rng, m * ksub, d // m).reshape(m, ksub, d // m)
    cb[0, 1] = cb[0, 0]  # duplicate centroid: strict `<` must keep the lower index
    v = uniform(rng, 500, d)
    v[0, : d // m] = cb[0, 0]
    pq = _pq(d, m, ksub, cb)
    codes = pq.encode(v)
    exp_codes = orc.pq_encode(1, cb, v)
    assert np.array_equal(codes, exp_codes)
    assert codes[0, 0] == 0
    assert np.array_equal(pq.decode(codes).view(np.uint32), orc.pq_decode(cb, codes).view(np.uint32))
    q = uniform(rng, 1, d)[0]
    t = pq.build_distance_tables(q)
    et = orc.pq_build_tables(cb, q)
    assert np.array_equal(t.view(np.uint32), et.view(np.uint32))
    td = pq.table_distance(t, codes)
    assert np.array_equal(td.view(np.uint32), orc.pq_table_distance(et, codes).view(np.uint32))
    ad = pq.asymmetric_distance(q, codes)
    ea = orc.pq_asymmetric_distance(cb, q, codes)
    assert np.array_equal(ad.view(np.uint32), ea.view(np.uint32))
    np.testing.assert_allclose(td, ad, atol=1e-3)  # pq.rs:639-669
    # round trip property: encode(decode(codes)) == codes when centroids are distinct
    if ksub >= 16:
        cb2 = cb.copy()
        cb2[0, 1] += 0.5
        pq2 = _pq(d, m, ksub, cb2)
        c2 = pq2.encode(v)
        assert np.array_equal(pq2.encode(pq2.decode(c2)), c2)


@pytest.mark.parametrize("metric", [0, 2, 3])
def test_encode_other_metrics(gpu_lib, orc, metric):
    rng = np.random.RandomState(9 + metric)
    cb = uniform(rng, 4 * 32, 8).reshape(4, 32, 8)
    v = uniform(rng, 300, 32)
    pq = _pq(32, 4, 32, cb, metric)
    assert np.array_equal(pq.encode(v), orc.pq_encode(metric, cb, v))


def test_reference_pq_behaviour(gpu_lib, orc):
    """pq.rs:523-735 unit tests restated through the host mirror."""
    from islands_b200 import DimensionMismatch, EmptyCollection, InvalidConfig, PQConfig, PQError, ProductQuantizer

    pq = ProductQuantizer(128, PQConfig(8, 16, 10, 42))
    assert not pq.is_trained() and pq.num_subquantizers() == 8
    assert ProductQuantizer(128, PQConfig(8, 256)).compression_ratio() == 64.0  # pq.rs:671-677
    with pytest.raises(InvalidConfig):
        ProductQuantizer(100, PQConfig(8, 256))  # pq.rs:531-534
    with pytest.raises(PQError):
        pq.encode(np.zeros(128, np.float32))  # not trained (pq.rs:610-614)
    with pytest.raises(EmptyCollection):
        pq.train(np.zeros((0, 128), np.float32))  # pq.rs:559-563
    with pytest.raises(DimensionMismatch):
        pq.train(np.zeros((10, 64), np.float32))  # pq.rs:566-571
    v = uniform(np.random.RandomState(42), 300, 128)
    pq.train(v)
    assert pq.is_trained()
    codes = pq.encode(v[0])
    assert codes.shape == (8,) and (codes < 16).all()
    dec = pq.decode(codes)
    assert dec.shape == (128,)
    with pytest.raises(DimensionMismatch):
        pq.encode(np.zeros(64, np.float32))
    with pytest.raises(PQError):
        pq.decode(np.zeros(5, np.uint16))  # wrong code count (pq.rs:251-257)
    with pytest.raises(PQError):
        pq.decode(np.full(8, 99, np.uint16))  # invalid code (pq.rs:262-266)
    assert pq.asymmetric_distance(v[1], codes) >= 0


def test_train_matches_oracle_kmeans(gpu_lib, orc):
    """A seeded `train` draws from the reference's own generator on both sides — `StdRng::seed_from_u64` of rand 0.8.5
    (ChaCha12; pq.rs:190-193), restated in csrc/std_rng.h and, independently, in the oracle — so the codebooks are
    bit-identical, not merely statistically alike (tests/test_std_rng.py pins the generator itself)."""
    from islands_b200 import PQConfig, ProductQuantizer

    rng = np.random.RandomState(5)
    v = uniform(rng, 1500, 32)
    pq = ProductQuantizer(32, PQConfig(4, 16, 6, 123))
    pq.train(v)
    got = pq.codebooks()
    exp = orc.pq_train(1, v, 4, 16, 6, 123)
    assert got.shape == exp.shape
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # quantisation error is what training is for: it must beat random codebooks
    err = np.linalg.norm(pq.decode(pq.encode(v)) - v, axis=1).mean()
    rnd = uniform(rng, 4 * 16, 8).reshape(4, 16, 8)
    err_rnd = np.linalg.norm(orc.pq_decode(rnd, orc.pq_encode(1, rnd, v)) - v, axis=1).mean()
    assert err < err_rnd
    # fewer training vectors than centroids: k = min(k, n) (pq.rs:374)
    small = ProductQuantizer(32, PQConfig(4, 16, 2, 1))
    small.train(v[:5])
    assert small.codebooks().shape == (4, 5, 8)
