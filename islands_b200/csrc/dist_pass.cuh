// dist_pass.cuh — reference-order distance evaluation, one candidate row per lane.
//
// The reference accumulates every distance as a strict left-to-right f32 fold with separately
// rounded multiply and add (src/core/distance.rs:71-122; no FMA, no reassociation).  IDs are
// only bit-exact if every distance that feeds a comparison is formed in that order, so the
// fold cannot be split across lanes.  Instead each lane of a warp owns ONE candidate row and
// walks it sequentially, while the warp cooperatively streams the 32 rows HBM -> shared
// memory in CH-float slices with 16-byte cp.async copies (coalesced: one row slice is CH*4
// contiguous bytes) through a STAGES-deep ring.  Rows sit in shared memory with a stride of
// CH+4 floats so that the per-lane 128-bit reads are bank-conflict free; the query slice is
// read as a broadcast.
#pragma once

#include "common.cuh"

namespace isl {

enum : int {
  ACC_DOT = 0,    // acc += x*y            (cosine numerator, dot product)
  ACC_L2 = 1,     // acc += (x-y)*(x-y)    (euclidean, PQ tables)
  ACC_L1 = 2,     // acc += |x-y|          (manhattan)
  ACC_SQNORM = 3  // acc += y*y            (per-row squared norm, cosine denominator)
};

__host__ __device__ constexpr int acc_kind_of_metric(int metric) {
  return metric == ISL_METRIC_EUCLIDEAN ? ACC_L2 : (metric == ISL_METRIC_MANHATTAN ? ACC_L1 : ACC_DOT);
}

template <int ACC>
__device__ __forceinline__ float acc_step(float acc, float x, float y) {
  if (ACC == ACC_DOT) return __fadd_rn(acc, __fmul_rn(x, y));
  if (ACC == ACC_L2) {
    float diff = __fsub_rn(x, y);
    return __fadd_rn(acc, __fmul_rn(diff, diff));
  }
  if (ACC == ACC_L1) return __fadd_rn(acc, fabsf(__fsub_rn(x, y)));
  return __fadd_rn(acc, __fmul_rn(y, y));
}

// Final step of DistanceMetric::calculate given the folded accumulator (distance.rs:82-87,
// :93, :114, :121).  na / nb are the query / row squared norms folded in the same order.
__device__ __forceinline__ float finalize_distance(int metric, float acc, float na, float nb) {
  if (metric == ISL_METRIC_COSINE) {
    float norm = __fsqrt_rn(__fmul_rn(na, nb));
    if (norm == 0.0f) return 1.0f;
    return __fsub_rn(1.0f, __fdiv_rn(acc, norm));
  }
  if (metric == ISL_METRIC_EUCLIDEAN) return __fsqrt_rn(acc);
  if (metric == ISL_METRIC_DOT) return -acc;
  return acc;
}

template <int CH>
struct StageGeom {
  static constexpr int ROW_STRIDE = CH + 4;         // floats; (CH+4)/4 odd => conflict-free LDS.128
  static constexpr int STAGE_FLOATS = 32 * ROW_STRIDE;
  static constexpr int VEC_PER_ROW = CH / 4;
};

// Sequential fold of lane `lane`'s row against the query, rows addressed through row_ids[]
// (shared memory, cnt <= 32 entries).  Returns the accumulator of the lane's row (undefined
// for lanes >= cnt).  All 32 lanes must call.  `ld` is the row stride in floats (multiple of 4,
// rows 16-byte aligned); `d` the logical dimension.
// WITH_NB additionally folds Σ y*y of the row (returned in *nb_out), which is what
// cosine_distance does in its single loop (distance.rs:76-80) when no precomputed norm exists.
template <int ACC, int CH, int STAGES, bool WITH_NB = false>
__device__ __forceinline__ float warp_rows_fold(const float* __restrict__ vectors, uint32_t ld,
                                                uint32_t d, const uint32_t* row_ids, uint32_t cnt,
                                                const float* q_smem, float* stage,
                                                float* nb_out = nullptr) {
  using G = StageGeom<CH>;
  const uint32_t lane = lane_id();
  const uint32_t nchunks = (d + CH - 1) / CH;
  const uint32_t pieces = cnt * G::VEC_PER_ROW;

  auto issue = [&](uint32_t c) {
    if (c < nchunks) {
      float* buf = stage + (c % STAGES) * G::STAGE_FLOATS;
      const uint32_t col0 = c * CH;
      for (uint32_t idx = lane; idx < pieces; idx += 32) {
        const uint32_t row = idx / G::VEC_PER_ROW;
        const uint32_t v = idx % G::VEC_PER_ROW;
        const uint32_t col = col0 + v * 4;
        if (col < ld)
          cp_async16(buf + row * G::ROW_STRIDE + v * 4,
                     vectors + (size_t)row_ids[row] * ld + col);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  float acc = 0.0f;
  float nb = 0.0f;
  for (uint32_t c = 0; c < nchunks; ++c) {
    issue(c + STAGES - 1);
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    if (lane < cnt) {
      const float4* row = reinterpret_cast<const float4*>(stage + (c % STAGES) * G::STAGE_FLOATS +
                                                          lane * G::ROW_STRIDE);
      const float4* qq = reinterpret_cast<const float4*>(q_smem + c * CH);
      if (c * CH + CH <= d) {
#pragma unroll
        for (int v = 0; v < G::VEC_PER_ROW; ++v) {
          const float4 y = row[v];
          const float4 x = qq[v];
          acc = acc_step<ACC>(acc, x.x, y.x);
          acc = acc_step<ACC>(acc, x.y, y.y);
          acc = acc_step<ACC>(acc, x.z, y.z);
          acc = acc_step<ACC>(acc, x.w, y.w);
          if (WITH_NB) {
            nb = acc_step<ACC_SQNORM>(nb, 0.0f, y.x);
            nb = acc_step<ACC_SQNORM>(nb, 0.0f, y.y);
            nb = acc_step<ACC_SQNORM>(nb, 0.0f, y.z);
            nb = acc_step<ACC_SQNORM>(nb, 0.0f, y.w);
          }
        }
      } else {
        const float* rf = reinterpret_cast<const float*>(row);
        const float* qf = reinterpret_cast<const float*>(qq);
        const uint32_t rem = d - c * CH;
        for (uint32_t t = 0; t < rem; ++t) {
          acc = acc_step<ACC>(acc, qf[t], rf[t]);
          if (WITH_NB) nb = acc_step<ACC_SQNORM>(nb, 0.0f, rf[t]);
        }
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  if (WITH_NB) *nb_out = nb;
  return acc;
}

// Σ x*x of a vector held in shared memory, folded left to right (distance.rs:78).  Every lane
// computes the same value (broadcast reads), so no shuffle is needed afterwards.
__device__ __forceinline__ float smem_sqnorm_fold(const float* q_smem, uint32_t d) {
  float s = 0.0f;
  for (uint32_t i = 0; i < d; ++i) {
    float x = q_smem[i];
    s = __fadd_rn(s, __fmul_rn(x, x));
  }
  return s;
}

}  // namespace isl
