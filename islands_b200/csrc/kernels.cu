// kernels.cu — standalone kernels: K1 batched distances, row norms, normalisation, K8 merge.
#include "kernels.h"

#include <algorithm>

#include "dist_pass.cuh"

namespace isl {

namespace {
constexpr int kCH = 64;
constexpr int kStages = 3;

// One warp per tile of 32 consecutive rows; persistent over tiles.
// MODE 0: distance(query,row) for `metric`;  MODE 1: Σ y*y per row.
template <int ACC, int MODE>
__global__ void __launch_bounds__(32)
rows_fold_kernel(int32_t metric, int squared, const float* __restrict__ query,
                 const float* __restrict__ rows, uint64_t n_rows, uint32_t d, uint32_t ld,
                 float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using G = StageGeom<kCH>;
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* q_smem = stage + kStages * G::STAGE_FLOATS;
  uint32_t* ids = reinterpret_cast<uint32_t*>(q_smem + ld);
  const uint32_t lane = lane_id();
  if (MODE == 0) {
    const float4* src = reinterpret_cast<const float4*>(query);
    float4* dst = reinterpret_cast<float4*>(q_smem);
    for (uint32_t i = lane; i < ld / 4; i += 32) dst[i] = src[i];
  } else {
    for (uint32_t i = lane; i < ld; i += 32) q_smem[i] = 0.0f;
  }
  __syncwarp();
  float na = 0.0f;
  if (MODE == 0 && metric == ISL_METRIC_COSINE) na = smem_sqnorm_fold(q_smem, d);
  const uint64_t tiles = (n_rows + 31) / 32;
  for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint64_t base = t * 32;
    const uint32_t cnt = (uint32_t)min((uint64_t)32, n_rows - base);
    // Row ids are relative to the tile base so that 32-bit ids suffice for any n_rows.
    ids[lane] = lane;
    __syncwarp();
    const float* tile_rows = rows + base * ld;
    float nb = 0.0f;
    float acc;
    if (MODE == 0 && ACC == ACC_DOT)
      acc = warp_rows_fold<ACC, kCH, kStages, true>(tile_rows, ld, d, ids, cnt, q_smem, stage, &nb);
    else
      acc = warp_rows_fold<ACC, kCH, kStages, false>(tile_rows, ld, d, ids, cnt, q_smem, stage);
    if (lane < cnt) {
      float r = acc;
      if (MODE == 0) {
        // calculate_squared (distance.rs:54-66): Euclidean returns the fold itself, the other
        // metrics return d * d.
        if (squared && metric == ISL_METRIC_EUCLIDEAN) {
          r = acc;
        } else {
          r = finalize_distance(metric, acc, na, nb);
          if (squared) r = __fmul_rn(r, r);
        }
      }
      out[base + lane] = r;
    }
    __syncwarp();
  }
}

template <int ACC, int MODE>
isl_status launch_rows_fold(int32_t metric, int squared, const float* q, const float* rows,
                            uint64_t n_rows, uint32_t d, uint32_t ld, float* out, int sms,
                            cudaStream_t st) {
  if (n_rows == 0) return ISL_OK;
  auto kern = rows_fold_kernel<ACC, MODE>;
  const size_t smem = (size_t)kStages * StageGeom<kCH>::STAGE_FLOATS * 4 + (size_t)ld * 4 + 32 * 4;
  if (smem > 227 * 1024) return fail(ISL_INVALID_ARGUMENT, "distance: dimension too large for shared memory");
  ISL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* the maximum, always: never lowered under a concurrent launch */));
  int per_sm = 0;
  ISL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) return fail(ISL_CUDA_ERROR, "distance kernel does not fit on an SM");
  const uint64_t tiles = (n_rows + 31) / 32;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)per_sm * sms);
  kern<<<grid, 32, smem, st>>>(metric, squared, q, rows, n_rows, d, ld, out);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

// x / ‖x‖ per row with the norm folded in reference order (distance.rs:125-132).
__global__ void normalize_rows_kernel(float* rows, const float* __restrict__ sq, uint64_t n_rows,
                                      uint32_t d, uint32_t ld) {
  const uint64_t total = n_rows * (uint64_t)d;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / d;
    const uint32_t c = (uint32_t)(i % d);
    const float norm = __fsqrt_rn(sq[r]);
    if (norm > 0.0f) rows[r * ld + c] = __fdiv_rn(rows[r * ld + c], norm);
  }
}

__global__ void pad_rows_kernel(const float* __restrict__ src, uint32_t d, float* __restrict__ dst,
                                uint32_t ld, uint64_t n) {
  const uint64_t total = n * (uint64_t)ld;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / ld;
    const uint32_t c = (uint32_t)(i % ld);
    dst[i] = c < d ? src[r * d + c] : 0.0f;
  }
}

__global__ void narrow_ids_kernel(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst,
                                  uint64_t count, uint64_t limit, unsigned int* flag) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t v = src[i];
    if (v >= limit) atomicExch(flag, 1u);
    dst[i] = (uint32_t)v;
  }
}

__global__ void pad_adjacency_kernel(const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ nbrs,
                                     uint64_t n, uint32_t stride, uint32_t* __restrict__ out) {
  const uint64_t total = n * stride;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t u = i / stride;
    const uint32_t j = (uint32_t)(i % stride);
    const uint64_t s = offsets[u], e = offsets[u + 1];
    out[i] = j < e - s ? nbrs[s + j] : 0xffffffffu;
  }
}

// One warp per node: sets *flag when a neighbour list names the same id twice.  A graph built by
// LeannIndex::build never does (a node is linked to a neighbour once), a CSR handed to from_csr may;
// the search keeps only the first occurrence (leann.rs:933-937), and can skip that test when no list
// has one.
__global__ void __launch_bounds__(256)
list_duplicates_kernel(const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ nbrs, uint64_t n,
                       unsigned int* __restrict__ flag) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t u = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
    const uint64_t s = offsets[u];
    const uint32_t deg = (uint32_t)(offsets[u + 1] - s);
    bool dup = false;
    for (uint32_t b = 0; b < deg; b += 32) {
      const bool valid = b + lane < deg;
      const uint32_t v = valid ? nbrs[s + b + lane] : 0xffffffffu - lane;  // padding lanes never match a real id pair
      // every warp-wide operation below is executed by all 32 lanes (b, c, t and deg are warp-uniform)
      const uint32_t same = __match_any_sync(0xffffffffu, v);
      const uint32_t live = __ballot_sync(0xffffffffu, valid);
      if (valid && __popc(same & live) > 1) dup = true;
      for (uint32_t c = 0; c < b; c += 32) {  // against the earlier chunks of the list (all full)
        const uint32_t w = nbrs[s + c + lane];
        for (uint32_t t = 0; t < 32; ++t) {
          const uint32_t wt = __shfl_sync(0xffffffffu, w, t);
          if (valid && wt == v) dup = true;
        }
      }
    }
    if (__any_sync(0xffffffffu, dup) && lane == 0) atomicExch(flag, 1u);
  }
}

// One warp per query.  The parts*k candidates are staged in shared memory, then the k smallest
// keys are extracted one at a time with a warp arg-min; positions (not keys) are retired, so
// duplicate (dist,id) pairs coming from different parts are kept, as a concat+sort would.
__global__ void __launch_bounds__(128)
merge_topk_kernel(const uint64_t* __restrict__ ids, const float* __restrict__ dist, uint32_t parts,
                  uint64_t nq, uint32_t k, uint64_t* __restrict__ out_ids,
                  float* __restrict__ out_dist, uint32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t total = parts * k;
  uint64_t* s_id = reinterpret_cast<uint64_t*>(smem_raw) + (size_t)warp * total;
  float* s_d = reinterpret_cast<float*>(reinterpret_cast<uint64_t*>(smem_raw) + (size_t)(blockDim.x >> 5) * total) +
               (size_t)warp * total;
  const uint64_t qi = blockIdx.x * (uint64_t)(blockDim.x >> 5) + warp;
  if (qi >= nq) return;
  for (uint32_t i = lane; i < total; i += 32) {
    const uint32_t p = i / k, j = i % k;
    const uint64_t off = ((uint64_t)p * nq + qi) * k + j;
    s_id[i] = ids[off];
    s_d[i] = dist[off];
  }
  __syncwarp();
  uint32_t produced = 0;
  for (uint32_t r = 0; r < k; ++r) {
    float bd = 0.0f;
    uint64_t bid = ISL_INVALID_ID;
    uint32_t bpos = 0xffffffffu;
    for (uint32_t i = lane; i < total; i += 32) {
      const uint64_t id = s_id[i];
      if (id == ISL_INVALID_ID) continue;
      const float dd = s_d[i];
      if (bpos == 0xffffffffu || key_lt64(dd, id, bd, bid) || (!key_lt64(bd, bid, dd, id) && i < bpos)) {
        bd = dd;
        bid = id;
        bpos = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, off);
      const uint64_t oid = __shfl_xor_sync(0xffffffffu, bid, off);
      const uint32_t opos = __shfl_xor_sync(0xffffffffu, bpos, off);
      if (opos != 0xffffffffu &&
          (bpos == 0xffffffffu || key_lt64(od, oid, bd, bid) || (!key_lt64(bd, bid, od, oid) && opos < bpos))) {
        bd = od;
        bid = oid;
        bpos = opos;
      }
    }
    if (bpos == 0xffffffffu) break;
    if (lane == 0) {
      out_ids[qi * k + r] = bid;
      out_dist[qi * k + r] = bd;
      s_id[bpos] = ISL_INVALID_ID;
    }
    produced++;
    __syncwarp();
  }
  for (uint32_t r = produced + lane; r < k; r += 32) {
    out_ids[qi * k + r] = ISL_INVALID_ID;
    out_dist[qi * k + r] = __int_as_float(0x7f800000);
  }
  if (lane == 0 && out_count) out_count[qi] = produced;
}

inline uint32_t grid_1d(uint64_t total, int threads) {
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((total + threads - 1) / threads, 148 * 16));
}
}  // namespace

isl_status launch_distance_batch(int32_t metric, bool squared, const float* d_query,
                                 const float* d_rows, uint64_t n_rows, uint32_t d, uint32_t ld,
                                 float* d_out, int sms, cudaStream_t st) {
  const int sq = squared ? 1 : 0;
  switch (acc_kind_of_metric(metric)) {
    case ACC_DOT: return launch_rows_fold<ACC_DOT, 0>(metric, sq, d_query, d_rows, n_rows, d, ld, d_out, sms, st);
    case ACC_L2: return launch_rows_fold<ACC_L2, 0>(metric, sq, d_query, d_rows, n_rows, d, ld, d_out, sms, st);
    default: return launch_rows_fold<ACC_L1, 0>(metric, sq, d_query, d_rows, n_rows, d, ld, d_out, sms, st);
  }
}

isl_status launch_row_sqnorms(const float* d_rows, uint64_t n_rows, uint32_t d, uint32_t ld,
                              float* d_out, int sms, cudaStream_t st) {
  return launch_rows_fold<ACC_SQNORM, 1>(0, 0, nullptr, d_rows, n_rows, d, ld, d_out, sms, st);
}

isl_status launch_normalize_rows(float* d_rows, uint64_t n_rows, uint32_t d, uint32_t ld, int sms,
                                 cudaStream_t st) {
  if (n_rows == 0) return ISL_OK;
  DevBuf<float> sq;
  ISL_CUDA_TRY(sq.alloc(n_rows));
  ISL_TRY(launch_row_sqnorms(d_rows, n_rows, d, ld, sq.p, sms, st));
  normalize_rows_kernel<<<grid_1d(n_rows * d, 256), 256, 0, st>>>(d_rows, sq.p, n_rows, d, ld);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  ISL_CUDA_TRY(cudaStreamSynchronize(st));  // sq is freed on return
  return ISL_OK;
}

isl_status launch_pad_rows(const float* d_src, uint32_t d, float* d_dst, uint32_t ld, uint64_t n,
                           cudaStream_t st) {
  if (n == 0) return ISL_OK;
  pad_rows_kernel<<<grid_1d(n * ld, 256), 256, 0, st>>>(d_src, d, d_dst, ld, n);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_narrow_ids(const uint64_t* d_src, uint32_t* d_dst, uint64_t count, uint64_t limit,
                             unsigned int* d_flag, cudaStream_t st) {
  if (count == 0) return ISL_OK;
  narrow_ids_kernel<<<grid_1d(count, 256), 256, 0, st>>>(d_src, d_dst, count, limit, d_flag);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_pad_adjacency(const uint64_t* d_offsets, const uint32_t* d_nbrs, uint64_t n, uint32_t stride,
                                uint32_t* d_out, cudaStream_t st) {
  if (n == 0) return ISL_OK;
  pad_adjacency_kernel<<<grid_1d(n * stride, 256), 256, 0, st>>>(d_offsets, d_nbrs, n, stride, d_out);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

__global__ void degree_counts_kernel(const uint64_t* __restrict__ offsets, uint64_t n, uint32_t* __restrict__ deg) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    deg[i] = (uint32_t)(offsets[i + 1] - offsets[i]);
}

isl_status launch_degree_counts(const uint64_t* d_offsets, uint64_t n, uint32_t* d_deg, cudaStream_t st) {
  if (n == 0) return ISL_OK;
  degree_counts_kernel<<<grid_1d(n, 256), 256, 0, st>>>(d_offsets, n, d_deg);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_list_duplicates(const uint64_t* d_offsets, const uint32_t* d_nbrs, uint64_t n, unsigned int* d_flag,
                                  cudaStream_t st) {
  if (n == 0) return ISL_OK;
  list_duplicates_kernel<<<grid_1d(n * 32, 256), 256, 0, st>>>(d_offsets, d_nbrs, n, d_flag);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_merge_topk(const uint64_t* d_ids, const float* d_dist, uint32_t parts, uint64_t nq,
                             uint32_t k, uint64_t* d_out_ids, float* d_out_dist,
                             uint32_t* d_out_count, cudaStream_t st) {
  if (nq == 0 || k == 0) return ISL_OK;
  const uint32_t warps = 4;
  const size_t smem = (size_t)warps * parts * k * 12;
  if (smem > 200 * 1024) return fail(ISL_INVALID_ARGUMENT, "merge: parts*k too large for shared memory");
  ISL_CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* the maximum, always: never lowered under a concurrent launch */));
  const uint32_t grid = (uint32_t)((nq + warps - 1) / warps);
  merge_topk_kernel<<<grid, warps * 32, smem, st>>>(d_ids, d_dist, parts, nq, k, d_out_ids, d_out_dist,
                                                     d_out_count);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

}  // namespace isl
