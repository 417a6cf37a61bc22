"""The generator behind seeded PQ training: rand 0.8.5 `StdRng::seed_from_u64` (ChaCha12Rng; pq.rs:190-193), a
third-party dependency of the reference (Cargo.lock rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4) restated
twice — islands_b200/csrc/std_rng.h (64-word BlockRng buffer) and oracle/oracle.cpp (continuous word stream).
Pinned here by the published known answers of the ChaCha block function (RFC 7539 §2.3.2 for 20 rounds; the
all-zero key vectors of draft-strombergson-chacha-test-vectors for 20 / 12 / 8 rounds) and by the agreement of the
two restatements on scripted draw sequences, including 64-bit draws that straddle a buffer refill.  Host code: no GPU."""
import ctypes as C
import struct

import numpy as np

from oracle import pyoracle as orc


def _block(key_words, counter, stream, rounds):
    key = (C.c_uint32 * 8)(*key_words)
    out = (C.c_uint32 * 16)()
    lib = orc.lib()
    lib.orc_chacha_block.restype = None
    lib.orc_chacha_block.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_uint32)]
    lib.orc_chacha_block(key, counter, stream, rounds, out)
    return struct.pack("<16I", *out).hex()


def test_chacha_block_known_answers():
    # RFC 7539 §2.3.2: key 00..1f, block counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00 (IETF layout: word 12 =
    # counter, words 13-15 = nonce; here words 12-13 = 64-bit counter, 14-15 = 64-bit stream)
    key = struct.unpack("<8I", bytes(range(32)))
    assert _block(key, 1 | (0x09000000 << 32), 0x4A000000, 20) == (
        "10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
        "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    zero = [0] * 8
    assert _block(zero, 0, 0, 20).startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    # ChaCha12 — the rounds StdRng uses — and ChaCha8, all-zero key and IV
    assert _block(zero, 0, 0, 12) == ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                                      "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    assert _block(zero, 0, 0, 8).startswith("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e")


def _draw(fn, seed, kinds, bound):
    kinds = np.ascontiguousarray(kinds, np.uint8)
    out = np.zeros(len(kinds), np.uint64)
    fn(C.c_uint64(seed), kinds.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(len(kinds)), C.c_uint64(bound),
       out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def test_two_restatements_agree_and_follow_the_stream_rules():
    from islands_b200 import _ffi

    prod = _ffi.load().isl_std_rng_draw
    olib = orc.lib()
    olib.orc_std_rng_draw.restype = None
    olib.orc_std_rng_draw.argtypes = [C.c_uint64, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    rng = np.random.RandomState(0)
    for seed in (0, 1, 42, 2 ** 63 + 12345, 2 ** 64 - 1):
        for bound in (1, 2, 3, 1000, 20000, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 7):
            kinds = rng.randint(0, 4, size=500)
            a, b = _draw(prod, seed, kinds, bound), _draw(olib.orc_std_rng_draw, seed, kinds, bound)
            assert np.array_equal(a, b), (seed, bound)
            assert (a[kinds == 3] < bound).all()
            f = a[kinds == 2].astype(np.uint32).view(np.float32)
            assert ((f >= 0) & (f < 1)).all()
    # word-stream rules: a u64 is two consecutive words, low first — also across the 64-word buffer boundary
    words = _draw(prod, 7, np.zeros(200, np.uint8), 1)
    mixed = _draw(prod, 7, np.array([0] * 63 + [1] + [0] * 10, np.uint8), 1)  # the u64 takes words 63 and 64
    assert np.array_equal(mixed[:63], words[:63])
    assert int(mixed[63]) == int(words[63]) | (int(words[64]) << 32)
    assert np.array_equal(mixed[64:], words[65:75])
    pairs = _draw(prod, 7, np.ones(50, np.uint8), 1)
    assert all(int(pairs[i]) == int(words[2 * i]) | (int(words[2 * i + 1]) << 32) for i in range(50))
    # gen::<f32>() = (word >> 8) * 2^-24
    fl = _draw(prod, 7, np.full(20, 2, np.uint8), 1).astype(np.uint32).view(np.float32)
    assert np.array_equal(fl, ((words[:20] >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)))
    # the seed expansion is PCG32 (XSH-RR): different seeds give unrelated streams, same seed the same stream
    assert not np.array_equal(_draw(prod, 1, np.zeros(8, np.uint8), 1), _draw(prod, 2, np.zeros(8, np.uint8), 1))
    assert np.array_equal(words[:8], _draw(prod, 7, np.zeros(8, np.uint8), 1))


def test_std_rng_construction_known_answer_of_the_rand_crate():
    """rand 0.8's own test of its standard generator (`rand/src/rngs/std.rs`, `test_stdrng_construction`): a StdRng built
    `from_seed([1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0, 0, ...])` returns 10719222850664546238 from `next_u64`, and a
    second StdRng built `from_rng` of the first (its seed is the next 32 bytes of the first one's stream) returns
    14064965282130556830.  Reproduced from the oracle's block function: this pins StdRng = ChaCha with TWELVE rounds on a
    non-zero key, the key layout (seed bytes as little-endian words), the zero counter / stream start and `next_u64` =
    two consecutive words, low word first — on published values, not on a restatement."""
    seed = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)
    first = bytes.fromhex(_block(struct.unpack("<8I", seed), 0, 0, 12))
    words = struct.unpack("<16I", first)
    assert words[0] | (words[1] << 32) == 10719222850664546238
    second = struct.unpack("<16I", bytes.fromhex(_block(words[2:10], 0, 0, 12)))  # from_rng: fill_bytes(32) = words 2..9
    assert second[0] | (second[1] << 32) == 14064965282130556830
    for rounds in (8, 20):  # the value is specific to twelve rounds
        w = struct.unpack("<16I", bytes.fromhex(_block(struct.unpack("<8I", seed), 0, 0, rounds)))
        assert w[0] | (w[1] << 32) != 10719222850664546238
