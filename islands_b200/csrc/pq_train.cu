// pq_train.cu — ProductQuantizer::train on the GPU (src/core/pq.rs:175-218, kmeans :362-463).
//
// Same algorithm and the same f32 operation order as the reference: distance-weighted seeding
// (weights are distances, not squared), Lloyd iterations with strict `<` assignment, centroid
// sums accumulated in vector order and divided by the count, empty clusters reseeded with a
// random training vector.  The random numbers come from the reference's own generator, restated in
// std_rng.h: `StdRng::seed_from_u64(seed)` (rand 0.8.5 = ChaCha12), consumed in the reference's order
// (one generator across the subspaces; `gen::<usize>() % n`, `gen::<f32>()`, `choose`), so a seeded
// training follows the reference's stream; the oracle restates the same generator independently and
// codebooks match it bit for bit.
#include <cub/cub.cuh>

#include <algorithm>
#include <random>

#include "api_common.h"
#include "dist_pass.cuh"
#include "std_rng.h"

namespace isl {
namespace {

__device__ __forceinline__ float metric_fold(int32_t metric, const float* a, const float* b, uint32_t d) {
  float dot = 0.0f, na = 0.0f, nb = 0.0f, s = 0.0f;
  if (metric == ISL_METRIC_COSINE) {
    for (uint32_t t = 0; t < d; ++t) {
      const float x = a[t], y = b[t];
      dot = __fadd_rn(dot, __fmul_rn(x, y));
      na = __fadd_rn(na, __fmul_rn(x, x));
      nb = __fadd_rn(nb, __fmul_rn(y, y));
    }
    return finalize_distance(metric, dot, na, nb);
  }
  if (metric == ISL_METRIC_EUCLIDEAN) {
    for (uint32_t t = 0; t < d; ++t) s = acc_step<ACC_L2>(s, a[t], b[t]);
  } else if (metric == ISL_METRIC_MANHATTAN) {
    for (uint32_t t = 0; t < d; ++t) s = acc_step<ACC_L1>(s, a[t], b[t]);
  } else {
    for (uint32_t t = 0; t < d; ++t) s = acc_step<ACC_DOT>(s, a[t], b[t]);
  }
  return finalize_distance(metric, s, 0.0f, 0.0f);
}

// mind[i] = min(mind[i], metric(v_i, centroid))   (pq.rs:385-393, running form of the fold)
__global__ void update_min_dist_kernel(int32_t metric, const float* __restrict__ vectors, uint32_t ld,
                                       uint32_t col0, uint32_t dsub, uint64_t n,
                                       const float* __restrict__ centroid, float* __restrict__ mind) {
  extern __shared__ float s_c[];
  for (uint32_t t = threadIdx.x; t < dsub; t += blockDim.x) s_c[t] = centroid[t];
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float dist = metric_fold(metric, vectors + i * ld + col0, s_c, dsub);
    mind[i] = fminf(mind[i], dist);
  }
}

__global__ void fill_kernel(float* p, uint64_t n, float v) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Sequential parts of the seeding step, one warp: lanes stage chunks into shared memory, lane 0
// folds them left to right.  total = Σ d_i (pq.rs:396); then the first i whose running sum of
// d_i/total reaches the threshold (pq.rs:398-413).
__global__ void __launch_bounds__(32) weighted_pick_kernel(const float* __restrict__ mind, uint64_t n,
                                                           float threshold, uint64_t* __restrict__ picked) {
  __shared__ float buf[1024];
  const uint32_t lane = threadIdx.x;
  float total = 0.0f;
  for (uint64_t base = 0; base < n; base += 1024) {
    const uint32_t cnt = (uint32_t)min((uint64_t)1024, n - base);
    for (uint32_t i = lane; i < cnt; i += 32) buf[i] = mind[base + i];
    __syncwarp();
    if (lane == 0)
      for (uint32_t i = 0; i < cnt; ++i) total = __fadd_rn(total, buf[i]);
    __syncwarp();
  }
  total = __shfl_sync(0xffffffffu, total, 0);
  const bool norm = total > 0.0f;
  float cumsum = 0.0f;
  uint64_t sel = 0;
  int done = 0;
  for (uint64_t base = 0; base < n && !done; base += 1024) {
    const uint32_t cnt = (uint32_t)min((uint64_t)1024, n - base);
    for (uint32_t i = lane; i < cnt; i += 32) buf[i] = mind[base + i];
    __syncwarp();
    if (lane == 0) {
      for (uint32_t i = 0; i < cnt; ++i) {
        const float w = norm ? __fdiv_rn(buf[i], total) : buf[i];
        cumsum = __fadd_rn(cumsum, w);
        if (cumsum >= threshold) {
          sel = base + i;
          done = 1;
          break;
        }
      }
    }
    done = __shfl_sync(0xffffffffu, done, 0);
    __syncwarp();
  }
  if (lane == 0) *picked = sel;
}

__global__ void copy_row_kernel(const float* __restrict__ vectors, uint32_t ld, uint32_t col0, uint32_t dsub,
                                const uint64_t* __restrict__ row_ptr, uint64_t row_imm, float* __restrict__ dst,
                                uint32_t ld_sub) {
  const uint64_t row = row_ptr ? *row_ptr : row_imm;
  for (uint32_t t = threadIdx.x; t < ld_sub; t += blockDim.x)
    dst[t] = t < dsub ? vectors[row * ld + col0 + t] : 0.0f;
}

__global__ void histogram_kernel(const uint16_t* __restrict__ codes, uint64_t n, uint32_t* __restrict__ counts) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(counts + codes[i], 1u);
}

__global__ void iota_kernel(uint32_t* p, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

// new_centroid[c][t] = (Σ over members in ascending vector index of v[t]) / count (pq.rs:436-451).
// One thread per (cluster, component); members of a cluster are contiguous in `order`.
__global__ void centroid_update_kernel(const float* __restrict__ vectors, uint32_t ld, uint32_t col0, uint32_t dsub,
                                       uint32_t ld_sub, const uint32_t* __restrict__ order,
                                       const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ counts,
                                       uint32_t k, float* __restrict__ centroids) {
  const uint32_t total = k * dsub;
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
    const uint32_t c = x / dsub, t = x % dsub;
    const uint32_t cnt = counts[c];
    if (cnt == 0) continue;  // reseeded by the host (pq.rs:452-456)
    const uint32_t s0 = seg_start[c];
    float sum = 0.0f;
    for (uint32_t i = 0; i < cnt; ++i) sum = __fadd_rn(sum, vectors[(size_t)order[s0 + i] * ld + col0 + t]);
    centroids[(size_t)c * ld_sub + t] = __fdiv_rn(sum, (float)cnt);
  }
}

inline uint32_t grid_1d(uint64_t total, int threads) {
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((total + threads - 1) / threads, 148 * 8));
}

}  // namespace
}  // namespace isl

using namespace isl;

extern "C" isl_status isl_pq_train(isl_pq* pq, const float* vectors, uint64_t n, uint32_t dim) {
  if (!pq) return fail(ISL_INVALID_ARGUMENT, "pq is null");
  if (n == 0) return fail(ISL_EMPTY_COLLECTION, "empty collection");  // pq.rs:176-178
  if (dim != pq->dim)                                                  // pq.rs:181-188
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(pq->dim) + ", got " +
                                      std::to_string(dim));
  if (!vectors) return fail(ISL_INVALID_ARGUMENT, "vectors is null");
  if (n >= (1ull << 32)) return fail(ISL_INVALID_ARGUMENT, "too many training vectors");
  DeviceGuard g(pq->device);
  std::lock_guard<std::mutex> lock(pq->mu);
  cudaStream_t st = pq->stream;
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  const uint32_t dsub = pq->dsub, ld_sub = pq->ld_sub;
  const uint32_t k = (uint32_t)std::min<uint64_t>(pq->cfg.num_centroids, n);  // pq.rs:374
  const uint32_t iters = (uint32_t)pq->cfg.training_iterations;
  uint64_t seed = pq->cfg.has_seed ? pq->cfg.seed : ((uint64_t)std::random_device{}() << 32) ^ std::random_device{}();
  StdRng rng(seed);  // StdRng::seed_from_u64 (pq.rs:190-193), one generator across subspaces, in order (:196-214)

  DevBuf<float> dv, mind, cent;
  DevBuf<uint16_t> codes, codes_sorted;
  DevBuf<uint32_t> order, order_sorted, counts, seg_start;
  DevBuf<uint64_t> picked;
  DevBuf<uint8_t> cub_tmp;
  ISL_CUDA_TRY(dv.alloc(n * dim));
  ISL_CUDA_TRY(mind.alloc(n));
  ISL_CUDA_TRY(cent.alloc((size_t)m * k * ld_sub));
  ISL_CUDA_TRY(codes.alloc(n));
  ISL_CUDA_TRY(codes_sorted.alloc(n));
  ISL_CUDA_TRY(order.alloc(n));
  ISL_CUDA_TRY(order_sorted.alloc(n));
  ISL_CUDA_TRY(counts.alloc(k));
  ISL_CUDA_TRY(seg_start.alloc(k));
  ISL_CUDA_TRY(picked.alloc(1));
  size_t cub_bytes = 0;
  ISL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, codes.p, codes_sorted.p, order.p, order_sorted.p,
                                               (int)n, 0, 16, st));
  ISL_CUDA_TRY(cub_tmp.alloc(cub_bytes + 16));
  ISL_CUDA_TRY(cudaMemcpyAsync(dv.p, vectors, n * dim * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemsetAsync(cent.p, 0, cent.bytes(), st));
  iota_kernel<<<grid_1d(n, 256), 256, 0, st>>>(order.p, n);
  count_launch();
  std::vector<uint32_t> h_counts(k), h_start(k);

  for (uint32_t j = 0; j < m; ++j) {
    const uint32_t col0 = j * dsub;
    float* cj = cent.p + (size_t)j * k * ld_sub;
    // ---- seeding (pq.rs:376-415) ---------------------------------------------------------
    const uint64_t first = rng.next_u64() % n;  // rng.gen::<usize>() % vectors.len() (pq.rs:380)
    copy_row_kernel<<<1, 128, 0, st>>>(dv.p, dim, col0, dsub, nullptr, first, cj, ld_sub);
    fill_kernel<<<grid_1d(n, 256), 256, 0, st>>>(mind.p, n, 3.402823466e+38f);
    count_launch(2);
    for (uint32_t c = 1; c < k; ++c) {
      update_min_dist_kernel<<<grid_1d(n, 128), 128, dsub * 4, st>>>(pq->metric, dv.p, dim, col0, dsub, n,
                                                                     cj + (size_t)(c - 1) * ld_sub, mind.p);
      const float threshold = rng.next_f32();  // rng.gen::<f32>() (pq.rs:404)
      weighted_pick_kernel<<<1, 32, 0, st>>>(mind.p, n, threshold, picked.p);
      copy_row_kernel<<<1, 128, 0, st>>>(dv.p, dim, col0, dsub, picked.p, 0, cj + (size_t)c * ld_sub, ld_sub);
      count_launch(3);
    }
    ISL_CUDA_TRY(cudaGetLastError());
    // ---- Lloyd iterations (pq.rs:420-460) -------------------------------------------------
    PqDev one;
    one.codebooks = cj;
    one.m = 1;
    one.ksub = k;
    one.dsub = dsub;
    one.ld_sub = ld_sub;
    one.metric = pq->metric;
    for (uint32_t it = 0; it < iters; ++it) {
      ISL_TRY(launch_pq_encode(one, dv.p + col0, dim, n, codes.p, pq->sms, st));
      ISL_CUDA_TRY(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
      histogram_kernel<<<grid_1d(n, 256), 256, 0, st>>>(codes.p, n, counts.p);
      size_t tb = cub_bytes;
      ISL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tb, codes.p, codes_sorted.p, order.p, order_sorted.p,
                                                   (int)n, 0, 16, st));
      count_launch(3);
      ISL_CUDA_TRY(cudaMemcpyAsync(h_counts.data(), counts.p, k * 4, cudaMemcpyDeviceToHost, st));
      ISL_CUDA_TRY(cudaStreamSynchronize(st));
      uint32_t run = 0;
      for (uint32_t c = 0; c < k; ++c) {
        h_start[c] = run;
        run += h_counts[c];
      }
      ISL_CUDA_TRY(cudaMemcpyAsync(seg_start.p, h_start.data(), k * 4, cudaMemcpyHostToDevice, st));
      centroid_update_kernel<<<grid_1d((uint64_t)k * dsub, 128), 128, 0, st>>>(
          dv.p, dim, col0, dsub, ld_sub, order_sorted.p, seg_start.p, counts.p, k, cj);
      count_launch();
      for (uint32_t c = 0; c < k; ++c) {
        if (h_counts[c] == 0) {  // vectors.choose(rng) (pq.rs:452-456)
          const uint64_t row = rng.choose_index(n);
          copy_row_kernel<<<1, 128, 0, st>>>(dv.p, dim, col0, dsub, nullptr, row, cj + (size_t)c * ld_sub, ld_sub);
          count_launch();
        }
      }
      ISL_CUDA_TRY(cudaGetLastError());
    }
  }
  // codebooks back to the host layout [m][k][dsub]
  pq->ksub = k;
  pq->h_codebooks.resize((size_t)m * k * dsub);
  ISL_CUDA_TRY(cudaMemcpy2DAsync(pq->h_codebooks.data(), (size_t)dsub * 4, cent.p, (size_t)ld_sub * 4,
                                 (size_t)dsub * 4, (size_t)m * k, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  pq->d_codebooks = std::move(cent);
  pq->trained = true;  // pq.rs:216
  return ISL_OK;
}
