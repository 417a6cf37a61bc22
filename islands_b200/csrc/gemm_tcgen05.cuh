// gemm_tcgen05.cuh — the dense contraction of the recompute encoder:
//   out[M][N] = act( A[M][K] · W[N][K]^T + bias[N] ) (+ residual[M][N]),  bf16 operands, f32 accumulate.
//
// sm_100a only.  One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer  : cp.async.bulk.tensor.2d (128B-swizzled K-major tiles) into a 4-stage ring
//   warp 1   MMA issuer    : one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                            (128 x BN x 16 per instruction), accumulators in TMEM
//   warp 2   TMEM allocator: 2 x BN columns, so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 4-11 epilogue    : tcgen05.ld (32 lanes x 32 columns per warp), bias / GELU / residual in
//                            registers, bf16 (or f32) stores
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace isl {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int THREADS = 384;     // warps 0-3: TMA, MMA, TMEM alloc, idle; warps 4-11: epilogue
constexpr int EPI_WARPS = 8;

enum : int { EPI_NONE = 0, EPI_GELU = 1 };

struct Params {
  int M, N, K;
  const float* bias;               // [N] or null
  const __nv_bfloat16* residual;   // [M][N] or null (added after the activation)
  __nv_bfloat16* out_bf16;         // [M][N] or null
  float* out_f32;                  // [M][N] or null
  int epilogue;                    // EPI_*
};

template <int BN>
constexpr size_t smem_bytes() {
  return (size_t)(BN == 256 ? 3 : 4) * (BM * BK * 2 + BN * BK * 2) + (size_t)8 * 32 * (BN + 16) /* epilogue staging */ +
         1024 /* alignment slack */ + 256 /* barriers */;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GEMM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra GEMM_DONE;\n"
      "bra GEMM_WAIT;\n"
      "GEMM_DONE:\n"
      "}\n" ::"r"(s32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          s32(smem_dst)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; both operands K-major
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// 32 consecutive f32 columns of this warp's 32 TMEM lanes -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// BERT "gelu": x * 0.5 * (1 + erf(x / sqrt(2))).  erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7,
// far below the bf16 rounding of the stored activation): one reciprocal, one exp, five FMAs.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, __expf(-z * z), 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}


// Epilogue of one warp: 32 accumulator rows (TMEM lanes of this warp's quarter) x COLS columns
// starting at column `col0` of the output.  tcgen05.ld hands each lane one ROW, so direct stores
// would scatter every instruction over 32 rows (half-written 32-byte sectors — measured as 1.6x
// the L2 traffic of the whole GEMM).  Instead the warp transposes through its own shared-memory
// staging block: the residual rows come in and the finished rows go out with fully coalesced
// 16-byte accesses; bias / GELU / residual are applied in f32 registers in between (one rounding).
// Row stride COLS*2 + 16 bytes keeps the per-lane 16-byte accesses bank-conflict free.
template <int COLS>
__device__ __forceinline__ void epilogue_warp(const Params& p, uint32_t taddr, int row0, int col0, uint8_t* stage, int lane) {
  constexpr int ROW_BYTES = COLS * 2;
  constexpr int STRIDE = ROW_BYTES + 16;
  constexpr int LANES_PER_ROW = ROW_BYTES / 16;       // 16-byte pieces per row
  constexpr int ROWS_PER_PASS = 32 / LANES_PER_ROW;
  const int sub_row = lane / LANES_PER_ROW, piece = lane % LANES_PER_ROW;
  const bool staged = p.out_bf16 != nullptr;
  if (staged && p.residual) {
#pragma unroll 4
    for (int r = sub_row; r < 32; r += ROWS_PER_PASS) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < p.M)
        v = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)(row0 + r) * p.N + col0) + piece);
      *reinterpret_cast<uint4*>(stage + r * STRIDE + piece * 16) = v;
    }
    __syncwarp();
  }
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
#pragma unroll 1
  for (int c0 = 0; c0 < COLS; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + c0, v);
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
    if (p.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i);
        f[4 * i] += b.x;
        f[4 * i + 1] += b.y;
        f[4 * i + 2] += b.z;
        f[4 * i + 3] += b.w;
      }
    }
    if (p.epilogue == EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
    }
    uint4* srow = reinterpret_cast<uint4*>(stage + lane * STRIDE + c0 * 2);
    if (p.residual) {
      uint4 r4[4];
      if (staged) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r4[i] = srow[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          r4[i] = row_ok ? __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.N + col0 + c0) + i)
                         : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&r4[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = __bfloat1622float2(rb[j]);
          f[8 * i + 2 * j] += x.x;
          f[8 * i + 2 * j + 1] += x.y;
        }
      }
    }
    if (staged) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
        srow[i] = o;
      }
    }
    if (p.out_f32 && row_ok) {  // test / debug output: direct stores
      float4* o4 = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.N + col0 + c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
  }
  if (staged) {
    __syncwarp();
#pragma unroll 4
    for (int r = sub_row; r < 32; r += ROWS_PER_PASS) {
      if (row0 + r < p.M)
        reinterpret_cast<uint4*>(p.out_bf16 + (size_t)(row0 + r) * p.N + col0)[piece] =
            *reinterpret_cast<const uint4*>(stage + r * STRIDE + piece * 16);
    }
    __syncwarp();
  }
}

template <int BN>
__host__ __device__ constexpr size_t epi_stage_bytes() { return (size_t)EPI_WARPS * 32 * (BN + 16); }  // per warp: 32 rows x (BN/2 cols bf16 + pad)
template <int BN>
__host__ __device__ constexpr int stages_for() { return BN == 256 ? 3 : 4; }

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // power of two >= 32
  constexpr int STAGES = stages_for<BN>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = tiles + (size_t)STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + epi_stage_bytes<BN>());
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM, num_n = p.N / BN, num_k = p.K / BK;
  const int tiles_total = num_m * num_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_b) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, EPI_WARPS);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        const int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          mbar_expect_tx(full + stage, STAGE_BYTES);
          uint8_t* sa = tiles + (size_t)stage * STAGE_BYTES;
          tma_load_2d(sa, &tm_a, kb * BK, m_blk * BM, full + stage);
          tma_load_2d(sa + A_BYTES, &tm_b, kb * BK, n_blk * BN, full + stage);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        mbar_wait(tempty + acc, acc_phase ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_addr = s32(tiles + (size_t)stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = smem_desc_sw128(b_addr + k * UMMA_K * 2);
            tc_mma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty + stage);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(tfull + acc);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q + 32) = tile rows; the two warps
    // of a quarter split the tile's columns ----
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
      const int m_blk = t / num_n, n_blk = t % num_n;
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      epilogue_warp<BN / 2>(p, tmem_base + acc * BN + half * (BN / 2) + ((uint32_t)(q * 32) << 16), m_blk * BM + q * 32,
                            n_blk * BN + half * (BN / 2), epi_stage + (size_t)(warp - 4) * 32 * (BN + 16), lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster — the two SMs of one TPC — share one
// 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile, so a
// k-block costs 32 KB of L2 -> shared traffic per SM instead of 48 KB for the same tensor work
// (the single-CTA kernel is bound by exactly that feed, see profiles/).  The leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256); both CTAs' TMA loads signal the leader's full barrier,
// completion is multicast to both CTAs' empty / accumulator-full barriers, and the epilogue
// warps of both CTAs release the accumulator on the leader's barrier.
constexpr int PAIR_BN = 256;
constexpr int PAIR_STAGES = 4;
constexpr uint32_t PAIR_A_BYTES = BM * BK * 2, PAIR_B_BYTES = (PAIR_BN / 2) * BK * 2;
constexpr uint32_t PAIR_STAGE_BYTES = PAIR_A_BYTES + PAIR_B_BYTES;
constexpr size_t pair_smem_bytes() { return (size_t)PAIR_STAGES * PAIR_STAGE_BYTES + (size_t)8 * 32 * (PAIR_BN + 16) + 1024 + 256; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's barrier (peer bit cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(s32(smem_dst)),
      "l"(map), "r"(s32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {  // arrives on `bar` in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   s32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the same barrier of CTA rank 0
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(s32(bar))
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  constexpr int BN = PAIR_BN;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = tiles + (size_t)PAIR_STAGES * PAIR_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + epi_stage_bytes<BN>());
  uint64_t* full = bars;                          // [STAGES]  (used in the leader)
  uint64_t* empty = bars + PAIR_STAGES;           // [STAGES]
  uint64_t* tfull = bars + 2 * PAIR_STAGES;       // [2]
  uint64_t* tempty = bars + 2 * PAIR_STAGES + 2;  // [2]       (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PAIR_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_m = (p.M + 2 * BM - 1) / (2 * BM), num_n = p.N / BN, num_k = p.K / BK;
  const int tiles_total = num_m * num_n;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_b) : "memory");
    for (int s = 0; s < PAIR_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, 2 * EPI_WARPS);  // the epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer (both CTAs) ----
      uint32_t stage = 0, phase = 0;
      for (int t = cluster_id; t < tiles_total; t += num_clusters) {
        const int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          if (rank == 0) mbar_expect_tx(full + stage, 2 * PAIR_STAGE_BYTES);  // both CTAs' bytes land here
          uint8_t* sa = tiles + (size_t)stage * PAIR_STAGE_BYTES;
          tma_load_2d_pair(sa, &tm_a, kb * BK, m_blk * 2 * BM + (int)rank * BM, full + stage);
          tma_load_2d_pair(sa + PAIR_A_BYTES, &tm_b, kb * BK, n_blk * BN + (int)rank * (BN / 2), full + stage);
          if (++stage == PAIR_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---- MMA issuer (leader CTA only) ----
      constexpr uint32_t idesc =
          (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = cluster_id; t < tiles_total; t += num_clusters) {
        mbar_wait(tempty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_addr = s32(tiles + (size_t)stage * PAIR_STAGE_BYTES);
          const uint32_t b_addr = a_addr + PAIR_A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = smem_desc_sw128(b_addr + k * UMMA_K * 2);
            tc_mma_bf16_pair(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_pair(empty + stage);
          if (++stage == PAIR_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit_pair(tfull + acc);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue (both CTAs): this CTA's TMEM holds its 128 rows of the 256-row tile ----
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = cluster_id; t < tiles_total; t += num_clusters) {
      const int m_blk = t / num_n, n_blk = t % num_n;
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      epilogue_warp<BN / 2>(p, tmem_base + acc * BN + half * (BN / 2) + ((uint32_t)(q * 32) << 16),
                            m_blk * 2 * BM + (int)rank * BM + q * 32, n_blk * BN + half * (BN / 2),
                            epi_stage + (size_t)(warp - 4) * 32 * (BN + 16), lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty + acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync_all();  // neither CTA may leave while the other still reads its shared memory or signals its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace gemm
}  // namespace isl
