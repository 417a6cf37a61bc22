// row_stream.cuh — the streaming engine of the search kernel: gathers candidate rows from HBM
// with TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a shared-memory ring and folds them in
// reference order, one row per lane (see dist_pass.cuh for why the fold cannot be split).
//
// Rows are consumed in groups of 32 (lane = row).  Each lane issues ONE bulk copy per ring slot
// for its own row slice (256 B .. 2 KB contiguous), completion is tracked by an mbarrier per
// slot (expect_tx / complete_tx), so staging costs a handful of instructions per slot instead
// of one LDGSTS per 16 bytes.  All groups of one expansion share the ring: the copies of the
// next group are in flight while the current group is being admitted.  A group with few rows
// gets proportionally wider slices (64 floats for > 16 rows ... 512 floats for <= 4 rows) so
// every slot carries ~8 KB however full the group is; the bytes in flight — and with them the
// HBM rate — stay constant.
#pragma once

#include "dist_pass.cuh"

namespace isl {

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared (1-D, no tensor map): 16-byte aligned, size multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// log2(floats per slice): CH floats for a full group, doubling as the group thins out (at most 8x)
template <int CH>
__device__ __forceinline__ uint32_t group_slice_shift(uint32_t cnt) {
  constexpr uint32_t B = CH == 64 ? 6u : (CH == 128 ? 7u : 8u);
  return cnt > 16 ? B : (cnt > 8 ? B + 1 : (cnt > 4 ? B + 2 : B + 3));
}

template <int STAGES>
struct RowRing {
  float* stage;        // STAGES * StageGeom<64>::STAGE_FLOATS floats
  uint64_t* bars;      // STAGES mbarriers
  uint32_t phase_bits; // bit s = parity the next wait on slot s must observe
  uint32_t islot;      // next slot to fill
  uint32_t cslot;      // next slot to consume
};

// Folds rows row_ids[0 .. total) against the query in shared memory and calls
// on_group(base, cnt, acc) after each group of <= 32 rows (acc = the lane's folded accumulator).
template <int ACC, int STAGES, int CH = 64, class OnGroup>
__device__ __forceinline__ void stream_rows_fold(RowRing<STAGES>& ring, const float* __restrict__ vectors, uint32_t ld,
                                                 uint32_t d, const uint32_t* row_ids, uint32_t total,
                                                 const float* q_smem, OnGroup&& on_group,
                                                 const uint32_t* __restrict__ row_of_id = nullptr) {
  using G = StageGeom<CH>;
  const uint32_t lane = lane_id();
  const uint32_t ngroups = (total + 31) >> 5;
  uint32_t ig = 0, ic = 0;                 // next slice to issue: group, chunk
  const float* my_row = nullptr;           // this lane's row of the group being issued
  uint32_t icnt = 0, ish = 6;
  auto load_issue_group = [&]() {
    if (ig < ngroups) {
      icnt = min(32u, total - (ig << 5));
      ish = group_slice_shift<CH>(icnt);
      my_row = nullptr;
      if (lane < icnt) {
        const uint32_t id = row_ids[(ig << 5) + lane];
        my_row = vectors + (size_t)(row_of_id ? __ldg(row_of_id + id) : id) * ld;
      }
    }
  };
  load_issue_group();
  auto issue_next = [&]() {
    if (ig < ngroups) {
      const uint32_t col0 = ic << ish;
      const uint32_t bytes = min(1u << ish, ld - col0) << 2;
      uint64_t* bar = ring.bars + ring.islot;
      if (lane == 0) mbar_arrive_expect_tx(bar, bytes * icnt);
      if (my_row)
        bulk_g2s(ring.stage + ring.islot * G::STAGE_FLOATS + lane * ((1u << ish) + 4), my_row + col0, bytes, bar);
      if (++ic == ((d + (1u << ish) - 1) >> ish)) {
        ic = 0;
        ++ig;
        load_issue_group();
      }
      ring.islot = (ring.islot + 1 == STAGES) ? 0 : ring.islot + 1;
    }
  };
  // Every slot is filled before the first wait, and a slot is refilled the moment its slice has been folded — BEFORE the
  // group's admission runs, so the first slice of the next group is in flight while the warp replays the admissions (with
  // a single slot nothing was in flight during on_group: one exposed DRAM latency per group, and the whole insert phase of
  // a rerank group).
#pragma unroll
  for (int s = 0; s < STAGES; ++s) issue_next();
  for (uint32_t g = 0; g < ngroups; ++g) {
    const uint32_t cnt = min(32u, total - (g << 5));
    const uint32_t sh = group_slice_shift<CH>(cnt);
    const uint32_t stride = (1u << sh) + 4;
    const uint32_t nch = (d + (1u << sh) - 1) >> sh;
    float acc = 0.0f;
    for (uint32_t c = 0; c < nch; ++c) {
      mbar_wait(ring.bars + ring.cslot, (ring.phase_bits >> ring.cslot) & 1u);
      ring.phase_bits ^= 1u << ring.cslot;
      if (lane < cnt) {
        const uint32_t col0 = c << sh;
        const uint32_t len = min(1u << sh, d - col0);
        const float4* row = reinterpret_cast<const float4*>(ring.stage + ring.cslot * G::STAGE_FLOATS + lane * stride);
        const float4* qq = reinterpret_cast<const float4*>(q_smem + col0);
        if (len == CH) {
#pragma unroll
          for (int v = 0; v < CH / 4; ++v) {
            const float4 y = row[v];
            const float4 x = qq[v];
            acc = acc_step<ACC>(acc, x.x, y.x);
            acc = acc_step<ACC>(acc, x.y, y.y);
            acc = acc_step<ACC>(acc, x.z, y.z);
            acc = acc_step<ACC>(acc, x.w, y.w);
          }
        } else {
          const uint32_t nvec = len >> 2;
#pragma unroll 8
          for (uint32_t v = 0; v < nvec; ++v) {
            const float4 y = row[v];
            const float4 x = qq[v];
            acc = acc_step<ACC>(acc, x.x, y.x);
            acc = acc_step<ACC>(acc, x.y, y.y);
            acc = acc_step<ACC>(acc, x.z, y.z);
            acc = acc_step<ACC>(acc, x.w, y.w);
          }
          const float* rf = reinterpret_cast<const float*>(row);
          const float* qf = reinterpret_cast<const float*>(qq);
          for (uint32_t t = nvec << 2; t < len; ++t) acc = acc_step<ACC>(acc, qf[t], rf[t]);
        }
      }
      __syncwarp();  // every lane is done with the slot before it is refilled
      ring.cslot = (ring.cslot + 1 == STAGES) ? 0 : ring.cslot + 1;
      issue_next();  // into the slot that was just consumed (islot follows cslot around the ring)
    }
    on_group(g << 5, cnt, acc);
  }
}

}  // namespace isl
