// std_rng.h — `rand::rngs::StdRng` of rand 0.8.5 (Cargo.lock: rand 0.8.5, rand_chacha 0.3.1, rand_core 0.6.4), the
// generator ProductQuantizer::train seeds with `StdRng::seed_from_u64(config.seed)` (src/core/pq.rs:190-193) and
// consumes as `gen::<usize>()`, `gen::<f32>()` and `SliceRandom::choose` (pq.rs:380, :404, :454).  The crates are
// third-party and not part of /root/reference; this restates their published algorithms:
//   * StdRng = ChaCha12Rng: ChaCha with 12 rounds, 256-bit key = the seed, 64-bit block counter in words 12-13,
//     64-bit stream id (0) in words 14-15; output words in block order.  The block function is pinned by the
//     RFC 7539 / draft-strombergson known answers (tests/test_std_rng.py).
//   * SeedableRng::seed_from_u64: the seed bytes are eight PCG32 (XSH-RR) outputs of the u64 state.
//   * BlockRng (64-word buffer = 4 blocks): next_u32 takes the next word; next_u64 takes two consecutive words,
//     low word first, also across a buffer refill.
//   * Standard f32: (next_u32 >> 8) * 2^-24.   usize on 64-bit targets: next_u64.
//   * gen_range(0..n) for n <= u32::MAX (SliceRandom::choose): widening-multiply rejection with
//     zone = (n << n.leading_zeros()) - 1.
// Host-side only (the training loop draws its few random numbers on the host, in the reference's order).
#pragma once

#include <cstdint>

namespace isl {

class StdRng {
 public:
  explicit StdRng(uint64_t seed_u64) {
    // rand_core::SeedableRng::seed_from_u64
    uint64_t state = seed_u64;
    for (int i = 0; i < 8; ++i) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
      const uint32_t rot = (uint32_t)(state >> 59);
      key_[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
    counter_ = 0;
    index_ = 64;  // empty buffer
  }

  uint32_t next_u32() {
    if (index_ >= 64) refill();
    return buf_[index_++];
  }
  uint64_t next_u64() {
    if (index_ < 63) {
      const uint64_t v = ((uint64_t)buf_[index_ + 1] << 32) | buf_[index_];
      index_ += 2;
      return v;
    }
    if (index_ >= 64) {
      refill();
      index_ = 2;
      return ((uint64_t)buf_[1] << 32) | buf_[0];
    }
    const uint64_t lo = buf_[63];  // the value straddles two buffers
    refill();
    index_ = 1;
    return ((uint64_t)buf_[0] << 32) | lo;
  }
  float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }  // Standard, 24 bits, [0, 1)
  // rng.gen_range(0..n), n in [1, 2^32): UniformInt<u32>::sample_single
  uint32_t gen_range_u32(uint32_t n) {
    const uint32_t zone = (n << __builtin_clz(n)) - 1u;
    for (;;) {
      const uint64_t m = (uint64_t)next_u32() * n;
      if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
    }
  }
  // SliceRandom::choose index for a slice of `len` elements (len >= 1)
  uint64_t choose_index(uint64_t len) {
    if (len <= 0xffffffffull) return gen_range_u32((uint32_t)len);
    // UniformInt<usize>::sample_single on 64-bit targets
    const uint64_t zone = (len << __builtin_clzll(len)) - 1ull;
    for (;;) {
      const unsigned __int128 m = (unsigned __int128)next_u64() * len;
      if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
    }
  }

  static void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                         (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t s[16];
    for (int i = 0; i < 16; ++i) s[i] = init[i];
    auto rotl = [](uint32_t x, int n) { return (x << n) | (x >> (32 - n)); };
    auto qr = [&](int a, int b, int c, int d) {
      s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
      s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
      s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
      s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    };
    for (int r = 0; r < rounds; r += 2) {
      qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
      qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = s[i] + init[i];
  }

 private:
  void refill() {  // four consecutive blocks
    for (int b = 0; b < 4; ++b) chacha_block(key_, counter_ + b, 0, 12, buf_ + 16 * b);
    counter_ += 4;
    index_ = 0;
  }
  uint32_t key_[8];
  uint64_t counter_;
  uint32_t buf_[64];
  uint32_t index_;
};

}  // namespace isl
