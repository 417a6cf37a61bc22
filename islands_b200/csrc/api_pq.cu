// api_pq.cu — C ABI of the product quantizer (src/core/pq.rs:116-359).
#include <algorithm>
#include <cstring>
#include <memory>

#include "api_common.h"

using namespace isl;

namespace isl {
// Uploads h_codebooks ([m][ksub][dsub]) into the padded device layout [m][ksub][ld_sub].
isl_status pq_upload_codebooks(isl_pq* pq) {
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  const size_t rows = (size_t)m * pq->ksub;
  ISL_CUDA_TRY(pq->d_codebooks.alloc(std::max<size_t>(rows * pq->ld_sub, 4)));
  ISL_CUDA_TRY(cudaMemsetAsync(pq->d_codebooks.p, 0, pq->d_codebooks.bytes(), pq->stream));
  if (rows)
    ISL_CUDA_TRY(cudaMemcpy2DAsync(pq->d_codebooks.p, (size_t)pq->ld_sub * 4, pq->h_codebooks.data(),
                                   (size_t)pq->dsub * 4, (size_t)pq->dsub * 4, rows,
                                   cudaMemcpyHostToDevice, pq->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(pq->stream));
  return ISL_OK;
}

static isl_status pq_require_trained(const isl_pq* pq) {
  if (!pq) return fail(ISL_INVALID_ARGUMENT, "pq is null");
  if (!pq->trained) return fail(ISL_PQ_ERROR, "Quantizer not trained");  // pq.rs:222-224
  return ISL_OK;
}

static isl_status dim_check(const isl_pq* pq, uint32_t dim) {
  if (dim != pq->dim)
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(pq->dim) + ", got " +
                                      std::to_string(dim));
  return ISL_OK;
}

}  // namespace isl

#include "std_rng.h"

extern "C" {

isl_status isl_pq_new(uint32_t dimension, const isl_pq_config* cfg, isl_pq** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  ISL_TRY(isl_pq_config_validate(cfg, dimension));  // pq.rs:133-134
  std::unique_ptr<isl_pq> pq(new isl_pq());
  pq->cfg = *cfg;
  pq->dim = dimension;
  pq->dsub = (uint32_t)(dimension / cfg->num_subquantizers);  // pq.rs:136
  pq->ld_sub = std::max<uint32_t>(4, round_up(pq->dsub, 4));
  pq->ksub = 0;
  ISL_TRY(current_device(&pq->device, &pq->sms));
  ISL_CUDA_TRY(cudaStreamCreateWithFlags(&pq->stream, cudaStreamNonBlocking));
  *out = pq.release();
  return ISL_OK;
} ISL_ABI_GUARD

void isl_pq_free(isl_pq* pq) {
  if (!pq) return;
  DeviceGuard g(pq->device);
  if (pq->stream) cudaStreamDestroy(pq->stream);
  delete pq;
}

isl_status isl_pq_set_metric(isl_pq* pq, int32_t metric) try {
  if (!pq) return fail(ISL_INVALID_ARGUMENT, "pq is null");
  if (metric < 0 || metric > 3) return fail(ISL_INVALID_CONFIG, "unknown metric");
  pq->metric = metric;
  return ISL_OK;
} ISL_ABI_GUARD
int32_t isl_pq_is_trained(const isl_pq* pq) { return pq && pq->trained ? 1 : 0; }
uint64_t isl_pq_num_subquantizers(const isl_pq* pq) { return pq ? pq->cfg.num_subquantizers : 0; }
float isl_pq_compression_ratio(const isl_pq* pq) {
  if (!pq) return 0.0f;  // pq.rs:168-172
  return (float)(pq->dim * 4ull) / (float)isl_pq_config_bytes_per_vector(&pq->cfg);
}

isl_status isl_pq_set_codebooks(isl_pq* pq, const float* codebooks, uint64_t num_centroids) try {
  if (!pq || !codebooks) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (num_centroids == 0 || num_centroids > 65536)
    return fail(ISL_INVALID_CONFIG, "num_centroids must be in range [1, 65536]");
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  pq->ksub = (uint32_t)num_centroids;
  const size_t total = (size_t)pq->cfg.num_subquantizers * num_centroids * pq->dsub;
  pq->h_codebooks.assign(codebooks, codebooks + total);
  ISL_TRY(pq_upload_codebooks(pq));
  pq->trained = true;
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_pq_get_codebooks(const isl_pq* pq, float* out, uint64_t* out_num_centroids) try {
  ISL_TRY(pq_require_trained(pq));
  if (out_num_centroids) *out_num_centroids = pq->ksub;
  if (out) std::memcpy(out, pq->h_codebooks.data(), pq->h_codebooks.size() * 4);
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_pq_encode(const isl_pq* pq, const float* vectors, uint64_t n, uint32_t dim,
                         uint16_t* out_codes) try {
  ISL_TRY(pq_require_trained(pq));
  ISL_TRY(dim_check(pq, dim));  // pq.rs:225-230
  if (n == 0) return ISL_OK;
  if (!vectors || !out_codes) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  DevBuf<float> dv;
  DevBuf<uint16_t> dc;
  ISL_CUDA_TRY(dv.alloc(n * dim));
  ISL_CUDA_TRY(dc.alloc(n * m));
  ISL_CUDA_TRY(cudaMemcpyAsync(dv.p, vectors, n * dim * 4, cudaMemcpyHostToDevice, pq->stream));
  ISL_TRY(launch_pq_encode(pq->dev(), dv.p, dim, n, dc.p, pq->sms, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_codes, dc.p, n * m * 2, cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(pq->stream));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_pq_decode(const isl_pq* pq, const uint16_t* codes, uint64_t n, uint64_t codes_per_vector,
                         float* out) try {
  ISL_TRY(pq_require_trained(pq));
  if (codes_per_vector != pq->cfg.num_subquantizers)  // pq.rs:251-257
    return fail(ISL_PQ_ERROR, "Expected " + std::to_string(pq->cfg.num_subquantizers) + " codes, got " +
                                  std::to_string(codes_per_vector));
  if (n == 0) return ISL_OK;
  if (!codes || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  DevBuf<uint16_t> dc;
  DevBuf<float> dout;
  DevBuf<unsigned int> flag;
  ISL_CUDA_TRY(dc.alloc(n * m));
  ISL_CUDA_TRY(dout.alloc(n * pq->dim));
  ISL_CUDA_TRY(flag.alloc(1));
  ISL_CUDA_TRY(cudaMemsetAsync(flag.p, 0, 4, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(dc.p, codes, n * m * 2, cudaMemcpyHostToDevice, pq->stream));
  ISL_TRY(launch_pq_decode(pq->dev(), dc.p, n, dout.p, flag.p, pq->stream));
  unsigned int h = 0;
  ISL_CUDA_TRY(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(out, dout.p, n * pq->dim * 4, cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(pq->stream));
  if (h) return fail(ISL_PQ_ERROR, "Invalid code (>= number of centroids)");  // pq.rs:264-266
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_pq_build_tables(const isl_pq* pq, const float* query, uint32_t dim, float* out_tables) try {
  if (!pq) return fail(ISL_INVALID_ARGUMENT, "pq is null");
  ISL_TRY(dim_check(pq, dim));  // pq.rs:308-313
  if (!query || !out_tables) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (pq->ksub == 0) return ISL_OK;  // untrained: empty tables (pq.rs:317-335 iterates no centroids)
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  DevBuf<float> dq, dt;
  ISL_CUDA_TRY(dq.alloc(dim));
  ISL_CUDA_TRY(dt.alloc((size_t)m * pq->ksub));
  ISL_CUDA_TRY(cudaMemcpyAsync(dq.p, query, (size_t)dim * 4, cudaMemcpyHostToDevice, pq->stream));
  ISL_TRY(launch_pq_tables(pq->dev(), dq.p, dim, 1, dt.p, pq->sms, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_tables, dt.p, dt.bytes(), cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(pq->stream));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_pq_table_distance(const isl_pq* pq, const float* tables, const uint16_t* codes,
                                 uint64_t n, float* out) try {
  ISL_TRY(pq_require_trained(pq));
  if (n == 0) return ISL_OK;
  if (!tables || !codes || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  DevBuf<float> dt, dout;
  DevBuf<uint16_t> dc;
  ISL_CUDA_TRY(dt.alloc((size_t)m * pq->ksub));
  ISL_CUDA_TRY(dc.alloc(n * m));
  ISL_CUDA_TRY(dout.alloc(n));
  ISL_CUDA_TRY(cudaMemcpyAsync(dt.p, tables, dt.bytes(), cudaMemcpyHostToDevice, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(dc.p, codes, n * m * 2, cudaMemcpyHostToDevice, pq->stream));
  ISL_TRY(launch_pq_table_distance(pq->dev(), dt.p, dc.p, n, dout.p, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(pq->stream));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_pq_asymmetric_distance(const isl_pq* pq, const float* query, uint32_t dim,
                                      const uint16_t* codes, uint64_t n, float* out) try {
  if (!pq) return fail(ISL_INVALID_ARGUMENT, "pq is null");
  ISL_TRY(dim_check(pq, dim));  // pq.rs:276-281
  if (n == 0) return ISL_OK;
  if (!query || !codes || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (pq->ksub == 0) return fail(ISL_PQ_ERROR, "Invalid code: quantizer holds no centroids");
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  DevBuf<float> dq, dout;
  DevBuf<uint16_t> dc;
  DevBuf<unsigned int> flag;
  ISL_CUDA_TRY(dq.alloc(dim));
  ISL_CUDA_TRY(dc.alloc(n * m));
  ISL_CUDA_TRY(dout.alloc(n));
  ISL_CUDA_TRY(flag.alloc(1));
  ISL_CUDA_TRY(cudaMemsetAsync(flag.p, 0, 4, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(dq.p, query, (size_t)dim * 4, cudaMemcpyHostToDevice, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(dc.p, codes, n * m * 2, cudaMemcpyHostToDevice, pq->stream));
  ISL_TRY(launch_pq_asymmetric(pq->dev(), dq.p, dc.p, n, dout.p, flag.p, pq->stream));
  unsigned int h = 0;
  ISL_CUDA_TRY(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, pq->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(pq->stream));
  if (h) return fail(ISL_PQ_ERROR, "Invalid code (>= number of centroids)");  // pq.rs:290-292
  return ISL_OK;
} ISL_ABI_GUARD

// The rounding of the ADC traversal's table entries (common.cuh), exposed so that it can be checked on its own.
isl_status isl_adc_table_round(const float* in, uint64_t count, float* out) try {
  if ((!in || !out) && count) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  for (uint64_t i = 0; i < count; ++i) {
    uint32_t u;
    memcpy(&u, in + i, 4);
    u = isl::bf16_round_bits(u);
    memcpy(out + i, &u, 4);
  }
  return ISL_OK;
} ISL_ABI_GUARD

// The generator isl_pq_train draws from (std_rng.h = rand 0.8.5 StdRng::seed_from_u64), exposed so that the
// restatement can be checked on its own: kind 0 = next_u32, 1 = next_u64, 2 = gen::<f32>() bits, 3 = choose(bound).
isl_status isl_std_rng_draw(uint64_t seed, const uint8_t* kinds, uint64_t count, uint64_t bound, uint64_t* out) try {
  if ((!kinds || !out) && count) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  isl::StdRng rng(seed);
  for (uint64_t i = 0; i < count; ++i) {
    switch (kinds[i]) {
      case 0: out[i] = rng.next_u32(); break;
      case 1: out[i] = rng.next_u64(); break;
      case 2: {
        const float f = rng.next_f32();
        uint32_t b;
        memcpy(&b, &f, 4);
        out[i] = b;
        break;
      }
      case 3:
        if (bound == 0) return fail(ISL_INVALID_ARGUMENT, "choose needs a non-empty range");
        out[i] = rng.choose_index(bound);
        break;
      default: return fail(ISL_INVALID_ARGUMENT, "unknown draw kind");
    }
  }
  return ISL_OK;
} ISL_ABI_GUARD

// Test hook: throws inside a guarded entry point, to show that nothing unwinds through the ABI (ISL_ABI_GUARD).
// kind 0 = std::bad_alloc, 1 = std::length_error (what a vector of an absurd size throws), 2 = a non-std exception.
isl_status isl_test_raise(int32_t kind) try {
  if (kind == 0) throw std::bad_alloc();
  if (kind == 1) {
    std::vector<float> v;
    v.resize(v.max_size() + 1);
  }
  if (kind == 2) throw 42;
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
