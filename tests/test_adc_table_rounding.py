"""The bfloat16 rounding of the ADC traversal's table entries (include/islands_b200.h, isl_index_search_adc_rerank):
the product's rule (csrc/common.cuh bf16_round_bits, through the isl_adc_table_round test hook — host code, no GPU needed)
and the oracle's independent restatement (oracle.cpp bf16_round) are both held against torch's float32 -> bfloat16
conversion (round to nearest even), a third implementation, on random values, every exponent, halfway cases, subnormals,
infinities and NaN."""
import ctypes as C

import numpy as np
import torch


def _cases():
    rng = np.random.RandomState(5)
    bits = [rng.randint(0, 2**32, size=200000, dtype=np.uint64).astype(np.uint32)]
    # every exponent with mantissas around the rounding boundary: ...7fff / 8000 / 8001 below an even and an odd kept bit
    exps = np.arange(0, 256, dtype=np.uint32) << 23
    for low in (0x00007fff, 0x00008000, 0x00008001, 0x00017fff, 0x00018000, 0x00018001, 0x007f7fff, 0x007f8000, 0x007fffff, 0):
        for sign in (0, 0x80000000):
            bits.append(exps | np.uint32(low) | np.uint32(sign))
    return np.concatenate(bits).view(np.float32)


def _torch_bf16(x):
    return torch.from_numpy(x.copy()).to(torch.bfloat16).to(torch.float32).numpy()


def _same(a, b):
    a, b = a.view(np.uint32), b.view(np.uint32)
    nan_a, nan_b = (a & 0x7fffffff) > 0x7f800000, (b & 0x7fffffff) > 0x7f800000
    return bool(np.array_equal(nan_a, nan_b) and np.array_equal(a[~nan_a], b[~nan_b]))


def test_oracle_rounding_matches_torch():
    from oracle import pyoracle as orc

    x = _cases()
    got = orc.adc_table_round(x)
    assert _same(got, _torch_bf16(x))
    assert not (got.view(np.uint32) & 0xffff).any()  # representable in 16 bits, NaN included


def test_product_rounding_matches_torch_and_oracle():
    from islands_b200 import _ffi
    from oracle import pyoracle as orc

    lib = _ffi.load()
    x = _cases()
    out = np.empty_like(x)
    f32p = C.POINTER(C.c_float)
    assert lib.isl_adc_table_round(x.ctypes.data_as(f32p), x.size, out.ctypes.data_as(f32p)) == 0
    assert _same(out, _torch_bf16(x))
    # product and oracle agree bit for bit, NaN payload included (both fold every NaN onto 0x7fc00000)
    assert np.array_equal(out.view(np.uint32), orc.adc_table_round(x).view(np.uint32))
