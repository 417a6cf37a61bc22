// pq.cu — product-quantizer kernels: K4 encode / decode, K3 table build and ADC lookups.
// Reference: src/core/pq.rs:86-106 (find_nearest), :221-271 (encode/decode), :275-348 (ADC).
#include <algorithm>

#include "dist_pass.cuh"
#include "kernels.h"

namespace isl {

namespace {
constexpr int kCH = 64;
constexpr int kStages = 3;

// Distance between a sub-vector held at xs (own shared-memory row of the lane) and a centroid
// at cs (shared memory, same address for all lanes => broadcast), folded exactly as
// DistanceMetric::calculate does (distance.rs:71-122) with a = sub-vector, b = centroid.
__device__ __forceinline__ float sub_distance(int32_t metric, const float* xs, const float* cs,
                                              uint32_t dsub) {
  if (metric == ISL_METRIC_EUCLIDEAN) {
    float s = 0.0f;
    for (uint32_t t = 0; t < dsub; ++t) s = acc_step<ACC_L2>(s, xs[t], cs[t]);
    return __fsqrt_rn(s);
  }
  if (metric == ISL_METRIC_MANHATTAN) {
    float s = 0.0f;
    for (uint32_t t = 0; t < dsub; ++t) s = acc_step<ACC_L1>(s, xs[t], cs[t]);
    return s;
  }
  if (metric == ISL_METRIC_DOT) {
    float s = 0.0f;
    for (uint32_t t = 0; t < dsub; ++t) s = acc_step<ACC_DOT>(s, xs[t], cs[t]);
    return -s;
  }
  float dot = 0.0f, na = 0.0f, nb = 0.0f;
  for (uint32_t t = 0; t < dsub; ++t) {
    const float x = xs[t], y = cs[t];
    dot = __fadd_rn(dot, __fmul_rn(x, y));
    na = __fadd_rn(na, __fmul_rn(x, x));
    nb = __fadd_rn(nb, __fmul_rn(y, y));
  }
  return finalize_distance(ISL_METRIC_COSINE, dot, na, nb);
}

// CTA = 128 lanes = 128 vectors of one subspace j.  The sub-vectors are staged once in shared
// memory (row stride odd => conflict-free per-lane reads); the subspace codebook is streamed
// through shared memory in tiles and read as a broadcast.  Each lane scans centroids in
// ascending order and keeps the first strict minimum (pq.rs:94-103).
__global__ void __launch_bounds__(128)
pq_encode_kernel(PqDev pq, const float* __restrict__ vectors, uint32_t ld, uint64_t n,
                 uint16_t* __restrict__ codes, uint32_t xstride, uint32_t cb_off, uint32_t tile_c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);           // [128][xstride]
  float* cb = xs + cb_off;                                  // [tile_c][ld_sub], 16B aligned
  const uint32_t j = blockIdx.y;
  const uint64_t tiles = (n + 127) / 128;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint64_t base = tile * 128;
    const uint32_t cnt = (uint32_t)min((uint64_t)128, n - base);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt * pq.dsub; i += 128) {
      const uint32_t r = i / pq.dsub, t = i % pq.dsub;
      xs[r * xstride + t] = vectors[(base + r) * ld + (uint64_t)j * pq.dsub + t];
    }
    float best = 3.402823466e+38f;  // f32::MAX (pq.rs:95)
    uint32_t best_c = 0;
    for (uint32_t c0 = 0; c0 < pq.ksub; c0 += tile_c) {
      const uint32_t tc = min(tile_c, pq.ksub - c0);
      __syncthreads();
      {
        const float4* src = reinterpret_cast<const float4*>(pq.codebooks + ((size_t)j * pq.ksub + c0) * pq.ld_sub);
        float4* dst = reinterpret_cast<float4*>(cb);
        for (uint32_t i = threadIdx.x; i < tc * pq.ld_sub / 4; i += 128) dst[i] = src[i];
      }
      __syncthreads();
      if (threadIdx.x < cnt) {
        const float* x = xs + threadIdx.x * xstride;
        for (uint32_t c = 0; c < tc; ++c) {
          const float dist = sub_distance(pq.metric, x, cb + c * pq.ld_sub, pq.dsub);
          if (dist < best) {
            best = dist;
            best_c = c0 + c;
          }
        }
      }
    }
    if (threadIdx.x < cnt) codes[(base + threadIdx.x) * pq.m + j] = (uint16_t)best_c;
  }
}

__global__ void pq_decode_kernel(PqDev pq, const uint16_t* __restrict__ codes, uint64_t n,
                                 float* __restrict__ out, unsigned int* flag) {
  const uint64_t d = (uint64_t)pq.m * pq.dsub;
  const uint64_t total = n * d;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / d;
    const uint32_t col = (uint32_t)(i % d);
    const uint32_t j = col / pq.dsub, t = col % pq.dsub;
    const uint32_t c = codes[r * pq.m + j];
    if (c >= pq.ksub) {
      atomicExch(flag, 1u);
      continue;
    }
    out[i] = pq.codebooks[((size_t)j * pq.ksub + c) * pq.ld_sub + t];
  }
}

// One warp per (query, subspace, 32 centroids): lane per centroid row, reference-order fold.
__global__ void __launch_bounds__(32)
pq_tables_kernel(PqDev pq, const float* __restrict__ queries, uint32_t q_ld, uint64_t nq,
                 float* __restrict__ tables) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using G = StageGeom<kCH>;
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* q_smem = stage + kStages * G::STAGE_FLOATS;  // [ld_sub]
  uint32_t* ids = reinterpret_cast<uint32_t*>(q_smem + pq.ld_sub);
  const uint32_t lane = lane_id();
  const uint32_t ctiles = (pq.ksub + 31) / 32;
  const uint64_t work = nq * pq.m * ctiles;
  for (uint64_t w = blockIdx.x; w < work; w += gridDim.x) {
    const uint32_t ct = (uint32_t)(w % ctiles);
    const uint32_t j = (uint32_t)((w / ctiles) % pq.m);
    const uint64_t qi = w / ((uint64_t)ctiles * pq.m);
    __syncwarp();
    for (uint32_t t = lane; t < pq.ld_sub; t += 32)
      q_smem[t] = t < pq.dsub ? queries[qi * q_ld + (uint64_t)j * pq.dsub + t] : 0.0f;
    const uint32_t c0 = ct * 32;
    const uint32_t cnt = min(32u, pq.ksub - c0);
    ids[lane] = c0 + lane;
    __syncwarp();
    const float acc = warp_rows_fold<ACC_L2, kCH, kStages>(
        pq.codebooks + (size_t)j * pq.ksub * pq.ld_sub, pq.ld_sub, pq.dsub, ids, cnt, q_smem, stage);
    if (lane < cnt) tables[(qi * pq.m + j) * pq.ksub + c0 + lane] = acc;
  }
}

__global__ void pq_table_distance_kernel(PqDev pq, const float* __restrict__ tables,
                                         const uint16_t* __restrict__ codes, uint64_t n,
                                         float* __restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (uint32_t j = 0; j < pq.m; ++j) {
      uint32_t c = codes[i * pq.m + j];
      if (c >= pq.ksub) c = pq.ksub - 1;  // reference would panic on an out-of-range index
      s = __fadd_rn(s, tables[(size_t)j * pq.ksub + c]);
    }
    out[i] = __fsqrt_rn(s);
  }
}

__global__ void pq_asymmetric_kernel(PqDev pq, const float* __restrict__ query,
                                     const uint16_t* __restrict__ codes, uint64_t n,
                                     float* __restrict__ out, unsigned int* flag) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    float total = 0.0f;
    bool bad = false;
    for (uint32_t j = 0; j < pq.m; ++j) {
      const uint32_t c = codes[i * pq.m + j];
      if (c >= pq.ksub) {  // get_centroid -> None -> PQError (pq.rs:290-292)
        bad = true;
        break;
      }
      const float* ce = pq.codebooks + ((size_t)j * pq.ksub + c) * pq.ld_sub;
      const float* qs = query + (size_t)j * pq.dsub;
      float sub = 0.0f;
      for (uint32_t t = 0; t < pq.dsub; ++t) sub = acc_step<ACC_L2>(sub, qs[t], ce[t]);
      total = __fadd_rn(total, sub);
    }
    if (bad) {
      atomicExch(flag, 1u);
      out[i] = __int_as_float(0x7fc00000);
    } else {
      out[i] = __fsqrt_rn(total);
    }
  }
}

inline uint32_t grid_1d(uint64_t total, int threads) {
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((total + threads - 1) / threads, 148 * 16));
}
}  // namespace

isl_status launch_pq_encode(const PqDev& pq, const float* d_vectors, uint32_t ld, uint64_t n,
                            uint16_t* d_codes, int sms, cudaStream_t st) {
  if (n == 0) return ISL_OK;
  const uint32_t xstride = pq.dsub | 1u;  // odd stride: lane r reads bank (r*xstride + t) % 32, all distinct
  const uint32_t cb_off = (128u * xstride + 3u) & ~3u;
  const uint32_t tile_c =
      std::max<uint32_t>(1, std::min<uint32_t>(pq.ksub, (96u * 1024u) / (pq.ld_sub * 4u)));
  const size_t smem = ((size_t)cb_off + (size_t)tile_c * pq.ld_sub) * 4;
  if (smem > 227 * 1024) return fail(ISL_PQ_ERROR, "pq encode: sub-vector dimension too large for shared memory");
  ISL_CUDA_TRY(cudaFuncSetAttribute(pq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* the maximum, always: never lowered under a concurrent launch */));
  const uint64_t tiles = (n + 127) / 128;
  const uint32_t gx = (uint32_t)std::min<uint64_t>(tiles, std::max<uint64_t>(1, (uint64_t)(2 * sms) / pq.m + 1));
  dim3 grid(gx, pq.m);
  pq_encode_kernel<<<grid, 128, smem, st>>>(pq, d_vectors, ld, n, d_codes, xstride, cb_off, tile_c);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_pq_decode(const PqDev& pq, const uint16_t* d_codes, uint64_t n, float* d_out,
                            unsigned int* d_flag, cudaStream_t st) {
  if (n == 0) return ISL_OK;
  pq_decode_kernel<<<grid_1d(n * pq.m * pq.dsub, 256), 256, 0, st>>>(pq, d_codes, n, d_out, d_flag);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_pq_tables(const PqDev& pq, const float* d_queries, uint32_t q_ld, uint64_t nq,
                            float* d_tables, int sms, cudaStream_t st) {
  if (nq == 0) return ISL_OK;
  const size_t smem = (size_t)kStages * StageGeom<kCH>::STAGE_FLOATS * 4 + (size_t)pq.ld_sub * 4 + 32 * 4;
  ISL_CUDA_TRY(cudaFuncSetAttribute(pq_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* the maximum, always: never lowered under a concurrent launch */));
  int per_sm = 0;
  ISL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pq_tables_kernel, 32, smem));
  if (per_sm < 1) return fail(ISL_CUDA_ERROR, "pq tables kernel does not fit on an SM");
  const uint64_t work = nq * pq.m * ((pq.ksub + 31) / 32);
  const uint32_t grid = (uint32_t)std::min<uint64_t>(work, (uint64_t)per_sm * sms);
  pq_tables_kernel<<<grid, 32, smem, st>>>(pq, d_queries, q_ld, nq, d_tables);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_pq_table_distance(const PqDev& pq, const float* d_tables, const uint16_t* d_codes,
                                    uint64_t n, float* d_out, cudaStream_t st) {
  if (n == 0) return ISL_OK;
  pq_table_distance_kernel<<<grid_1d(n, 256), 256, 0, st>>>(pq, d_tables, d_codes, n, d_out);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_pq_asymmetric(const PqDev& pq, const float* d_query, const uint16_t* d_codes,
                                uint64_t n, float* d_out, unsigned int* d_flag, cudaStream_t st) {
  if (n == 0) return ISL_OK;
  pq_asymmetric_kernel<<<grid_1d(n, 256), 256, 0, st>>>(pq, d_query, d_codes, n, d_out, d_flag);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

}  // namespace isl
