// common.cuh — shared host/device helpers for the islands_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/islands_b200.h"

namespace isl {

// ---- host-side error plumbing ---------------------------------------------------------
void set_last_error(const std::string& msg);
isl_status fail(isl_status st, const std::string& msg);
isl_status cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launch_count;

#define ISL_CUDA_TRY(expr)                                     \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::isl::cuda_fail(_e, #expr); \
  } while (0)

#define ISL_TRY(expr)                   \
  do {                                  \
    isl_status _s = (expr);             \
    if (_s != ISL_OK) return _s;        \
  } while (0)

// Nothing may unwind through the C ABI (the callers are Rust, ctypes, C): every `isl_status` entry point is a
// function-try-block closed by ISL_ABI_GUARD, so a std::bad_alloc / std::length_error out of a host container (an
// absurd size argument, an exhausted host) comes back as a status with a message instead of aborting the process.
isl_status exception_status() noexcept;
#define ISL_ABI_GUARD \
  catch (...) { return ::isl::exception_status(); }

inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// RAII device buffer (cudaMalloc / cudaFree).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// ---- device-side ordering helpers -----------------------------------------------------
// (OrderedFloat<f32>, u64) tuple order of the reference's heaps (leann.rs:701-702, 907-908):
// NaN is greater than every number and equal to itself; ties fall through to the id.
__host__ __device__ __forceinline__ bool of_lt(float a, float b) {
  bool an = a != a, bn = b != b;
  if (an || bn) return !an && bn;
  return a < b;
}
__host__ __device__ __forceinline__ bool key_lt(float d1, uint32_t id1, float d2, uint32_t id2) {
  if (of_lt(d1, d2)) return true;
  if (of_lt(d2, d1)) return false;
  return id1 < id2;
}
__host__ __device__ __forceinline__ bool key_lt64(float d1, uint64_t id1, float d2, uint64_t id2) {
  if (of_lt(d1, d2)) return true;
  if (of_lt(d2, d1)) return false;
  return id1 < id2;
}

// Table entries of the ADC TRAVERSAL (isl_index_search_adc_rerank / _adc_recompute; include/islands_b200.h): every entry of
// build_distance_tables (pq.rs:307-338) is rounded to bfloat16 (round to nearest even on the bit pattern, NaN -> quiet NaN)
// before the traversal folds it; the per-query table is then 2 bytes per entry in shared memory, which doubles the resident
// queries per SM.  The fold itself stays the f32 left fold + sqrt of pq.rs:341-348.  oracle.cpp restates the same rule.
__host__ __device__ __forceinline__ uint32_t bf16_round_bits(uint32_t u) {
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc00000u;
  return (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;
}

// The seeded draw stream of PruningStrategy::Proportional (include/islands_b200.h, isl_pruning_strategy).
__host__ __device__ __forceinline__ uint64_t splitmix_mix(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ float prune_draw(uint64_t seed, uint64_t query, uint64_t draw) {
  const uint64_t h = splitmix_mix(splitmix_mix(seed + 0x9E3779B97F4A7C15ull * (query + 1)) + 0x9E3779B97F4A7C15ull * (draw + 1));
  return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte asynchronous global->shared copy (LDGSTS), bypassing L1 (.cg): rows are streamed once.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
#endif

}  // namespace isl
