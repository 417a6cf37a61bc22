// bincode.h — the byte layout of the reference's to_bytes / from_bytes (leann.rs:1059-1066,
// pq.rs:351-358, hnsw.rs:507-514): `bincode::serialize`, i.e. the bincode 1.x default encoding of the
// serde data model — little-endian fixed-width integers, usize as u64, f32 / f64 raw, bool as one
// byte, Vec<T> / HashMap as a u64 length followed by the items, Option<T> as a one-byte tag (0 / 1)
// followed by T, a unit enum variant as its u32 index; struct fields in declaration order.
#pragma once

#include <cstdint>
#include <cstring>
#include <vector>

namespace isl {

struct ByteWriter {
  std::vector<uint8_t> buf;
  template <class T>
  void raw(T v) {
    const size_t o = buf.size();
    buf.resize(o + sizeof(T));
    std::memcpy(buf.data() + o, &v, sizeof(T));
  }
  void u8(uint8_t v) { raw(v); }
  void u32(uint32_t v) { raw(v); }
  void u64(uint64_t v) { raw(v); }
  void f32(float v) { raw(v); }
  void f64(double v) { raw(v); }
  void boolean(bool v) { u8(v ? 1 : 0); }
  void opt_u64(bool some, uint64_t v) {
    u8(some ? 1 : 0);
    if (some) u64(v);
  }
  void vec_u64(const uint64_t* p, uint64_t n) {
    u64(n);
    const size_t o = buf.size();
    buf.resize(o + n * 8);
    if (n) std::memcpy(buf.data() + o, p, n * 8);
  }
  void vec_f32(const float* p, uint64_t n) {
    u64(n);
    const size_t o = buf.size();
    buf.resize(o + n * 4);
    if (n) std::memcpy(buf.data() + o, p, n * 4);
  }
};

struct ByteReader {
  const uint8_t* p;
  uint64_t len, pos = 0;
  bool ok = true;
  ByteReader(const uint8_t* bytes, uint64_t n) : p(bytes), len(n) {}
  template <class T>
  T raw() {
    T v{};
    if (!ok || len - pos < sizeof(T)) {
      ok = false;
      return v;
    }
    std::memcpy(&v, p + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
  uint8_t u8() { return raw<uint8_t>(); }
  uint32_t u32() { return raw<uint32_t>(); }
  uint64_t u64() { return raw<uint64_t>(); }
  float f32() { return raw<float>(); }
  double f64() { return raw<double>(); }
  bool boolean() {
    const uint8_t b = u8();
    if (b > 1) ok = false;
    return b == 1;
  }
  bool opt_u64(uint64_t* v) {
    const uint8_t tag = u8();
    if (tag > 1) ok = false;
    if (tag == 1) *v = u64();
    return tag == 1;
  }
  // length prefix of a sequence whose items take at least `min_item` bytes each
  uint64_t seq_len(uint64_t min_item) {
    const uint64_t n = u64();
    if (!ok || (min_item && n > (len - pos) / min_item)) {
      ok = false;
      return 0;
    }
    return n;
  }
  void vec_u64(std::vector<uint64_t>* out) {
    const uint64_t n = seq_len(8);
    out->resize(ok ? n : 0);
    if (ok && n) {
      std::memcpy(out->data(), p + pos, n * 8);
      pos += n * 8;
    }
  }
  void vec_f32(std::vector<float>* out) {
    const uint64_t n = seq_len(4);
    out->resize(ok ? n : 0);
    if (ok && n) {
      std::memcpy(out->data(), p + pos, n * 4);
      pos += n * 4;
    }
  }
  bool done() const { return ok && pos == len; }
};

}  // namespace isl
