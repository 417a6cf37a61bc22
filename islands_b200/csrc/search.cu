// search.cu — launch planning for the batched best-first search kernel (search_core.cuh).
#include "search.h"

#include <algorithm>
#include <mutex>

#include "adc_traverse.cuh"
#include "search_core.cuh"

namespace isl {

// the exact traversal with R as a register bag: one translation unit per accumulation kind (search_bag.inc)
isl_status plan_exact_bag_dot(int nr, uint32_t ld, uint32_t u_cap, int sms, SearchPlan* plan);
isl_status plan_exact_bag_l2(int nr, uint32_t ld, uint32_t u_cap, int sms, SearchPlan* plan);
isl_status plan_exact_bag_l1(int nr, uint32_t ld, uint32_t u_cap, int sms, SearchPlan* plan);
isl_status launch_exact_bag_dot(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st);
isl_status launch_exact_bag_l2(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st);
isl_status launch_exact_bag_l1(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st);

namespace {
#ifndef ISL_CH
#define ISL_CH 128
#endif
constexpr int kCH = ISL_CH;  // floats per row slice of a full group (row_stream.cuh)
#ifndef ISL_STAGES
#define ISL_STAGES 1
#endif
constexpr int kStages = ISL_STAGES;
// R (8 B per entry) stays in shared memory up to this many entries; above, it lives in an
// L2-resident global buffer so that enough warps stay resident per SM.
constexpr uint32_t kEfSmemMax = 2048;
#ifndef ISL_BAG_EF_MAX
#define ISL_BAG_EF_MAX 1024
#endif
constexpr uint32_t kBagEfMax = ISL_BAG_EF_MAX;  // exact traversal: register bag up to this ef (0 disables it: dev builds)
constexpr uint32_t kLutSmemMaxFloats = 8192;  // 32 KB of PQ tables per query in shared memory
constexpr uint32_t kAqSmemMaxEntries = 2048;  // 16 KB approximate queue in shared memory
constexpr int kMaxDynSmem = 227 * 1024;

// The dynamic shared-memory limit is a process-wide attribute of a kernel instantiation: it is raised ONCE to the
// hardware maximum (call_once per instantiation and device), never to the size of one call — two threads planning
// different searches must not lower it under each other's launches.
template <class K>
isl_status opt_in_smem(K kern, std::once_flag* flags, cudaError_t* results) {
  int dev = 0;
  ISL_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(ISL_CUDA_ERROR, "search: device ordinal out of range");
  std::call_once(flags[dev], [&] {
    cudaFuncAttributes fa{};
    results[dev] = cudaFuncGetAttributes(&fa, kern);  // static shared memory (the tie-list header) counts against the 227 KB
    if (results[dev] == cudaSuccess)
      results[dev] = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem - (int)fa.sharedSizeBytes);
  });
  if (results[dev] != cudaSuccess) return cuda_fail(results[dev], "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  return ISL_OK;
}

template <int ACC, bool R_SMEM, int TWO>
isl_status plan_one(uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan) {
  auto kern = leann_search_kernel<ACC, kCH, kStages, R_SMEM, TWO>;
  const size_t smem = search_smem_bytes<kCH, kStages>(ld, R_SMEM ? ef : 0, u_cap, plan->lut_smem_floats,
                                                      plan->aq_smem_entries);
  if (smem > 226 * 1024)
    return fail(ISL_INVALID_ARGUMENT, "search: dimension / ef need more than 226 KB of shared memory per warp");
  static std::once_flag once[64];
  static cudaError_t once_result[64];
  ISL_TRY(opt_in_smem(kern, once, once_result));
  int per_sm = 0;
  ISL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) return fail(ISL_CUDA_ERROR, "search: kernel does not fit on an SM");
  plan->smem = smem;
  plan->ctas_per_sm = per_sm;
  plan->grid = (uint32_t)(per_sm * sms);
  plan->r_in_smem = R_SMEM;
  return ISL_OK;
}

template <int ACC, bool R_SMEM, int TWO>
isl_status launch_one(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st) {
  leann_search_kernel<ACC, kCH, kStages, R_SMEM, TWO><<<grid, 32, plan.smem, st>>>(args);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

template <bool R_SMEM>
isl_status plan_lean(uint32_t ef, uint32_t u_cap, uint32_t pq_m, int sms, SearchPlan* plan) {
  auto kern = leann_search_kernel<ACC_DOT, kCH, kStages, R_SMEM, 3, 0>;
  const size_t smem = search_smem_bytes<kCH, kStages>(0, R_SMEM ? ef : 0, u_cap, plan->lut_smem_floats, 0, true) +
                      (R_SMEM ? search_smem_bytes_idc() : 0);
  if (smem > 226 * 1024) return fail(ISL_INVALID_ARGUMENT, "search: ef needs more than 226 KB of shared memory per warp");
  plan->novis_ok = R_SMEM && plan->lut_smem_floats != 0 && (pq_m == 16 || pq_m == 32);  // and n < kIdcMaxNodes (checked by the caller)
  static std::once_flag once[64];
  static cudaError_t once_result[64];
  ISL_TRY(opt_in_smem(kern, once, once_result));
  int per_sm = 0;
  ISL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) return fail(ISL_CUDA_ERROR, "search: kernel does not fit on an SM");
  plan->smem = smem;
  plan->ctas_per_sm = per_sm;
  plan->grid = (uint32_t)(per_sm * sms);
  plan->r_in_smem = R_SMEM;
  return ISL_OK;
}

template <bool R_SMEM>
isl_status launch_lean(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st) {
  leann_search_kernel<ACC_DOT, kCH, kStages, R_SMEM, 3, 0><<<grid, 32, plan.smem, st>>>(args);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

// The register-bag traversal (adc_traverse.cuh): one-byte codes, m = 16 / 32, table in shared memory, ef <= 32 * NR.
template <int NR, int KS>
isl_status plan_bag(uint32_t ef, int sms, SearchPlan* plan) {
  auto kern = adc_traverse_kernel<NR, KS>;
  const size_t smem = adc_traverse_smem_bytes(plan->lut_smem_floats, ef);
  static std::once_flag once[64];
  static cudaError_t once_result[64];
  ISL_TRY(opt_in_smem(kern, once, once_result));
  int per_sm = 0;
  ISL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) return fail(ISL_CUDA_ERROR, "search: kernel does not fit on an SM");
  plan->novis_ok = true;  // and n < kIdcMaxNodes (checked by the caller)
  plan->smem = smem;
  plan->ctas_per_sm = per_sm;
  plan->grid = (uint32_t)(per_sm * sms);
  plan->r_in_smem = true;
  return ISL_OK;
}

template <int NR, int KS>
isl_status launch_bag(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st) {
  adc_traverse_kernel<NR, KS><<<grid, 32, plan.smem, st>>>(args);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

template <int TWO>
isl_status plan_dispatch(int acc, bool r_smem, uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan) {
#define ISL_PLAN(A)                                                                \
  case A:                                                                          \
    return r_smem ? plan_one<A, true, TWO>(ld, ef, u_cap, sms, plan)               \
                  : plan_one<A, false, TWO>(ld, ef, u_cap, sms, plan);
  switch (acc) {
    ISL_PLAN(ACC_DOT)
    ISL_PLAN(ACC_L2)
    ISL_PLAN(ACC_L1)
  }
#undef ISL_PLAN
  return fail(ISL_INVALID_CONFIG, "search: unknown metric");
}

template <int TWO>
isl_status launch_dispatch(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st) {
#define ISL_LAUNCH(A)                                                              \
  case A:                                                                          \
    return plan.r_in_smem ? launch_one<A, true, TWO>(plan, args, grid, st)         \
                          : launch_one<A, false, TWO>(plan, args, grid, st);
  switch (plan.acc) {
    ISL_LAUNCH(ACC_DOT)
    ISL_LAUNCH(ACC_L2)
    ISL_LAUNCH(ACC_L1)
  }
#undef ISL_LAUNCH
  return fail(ISL_INVALID_CONFIG, "search: unknown metric");
}
}  // namespace

isl_status plan_search(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan) {
  plan->acc = acc_kind_of_metric(metric);
  plan->two_level = false;
  plan->lut_smem_floats = 0;
  plan->aq_smem_entries = 0;
  plan->aq_cap = 0;
  plan->mode = 0;
  plan->nr = 0;
  // up to ef = 1024 the result set is an unsorted bag in registers (search_core.cuh), NR = 4 / 8 / 16 / 32 entries per lane
  if (ef <= kBagEfMax) {
    plan->nr = ef <= 128 ? 4 : (ef <= 256 ? 8 : (ef <= 512 ? 16 : 32));
    switch (plan->acc) {
      case ACC_DOT: return plan_exact_bag_dot(plan->nr, ld, u_cap, sms, plan);
      case ACC_L2: return plan_exact_bag_l2(plan->nr, ld, u_cap, sms, plan);
      case ACC_L1: return plan_exact_bag_l1(plan->nr, ld, u_cap, sms, plan);
    }
    return fail(ISL_INVALID_CONFIG, "search: unknown metric");
  }
  return plan_dispatch<0>(plan->acc, ef <= kEfSmemMax, ld, ef, u_cap, sms, plan);
}

isl_status plan_search_two_level(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, uint32_t pq_m,
                                 uint32_t pq_ksub, uint32_t aq_cap, int sms, SearchPlan* plan) {
  plan->acc = acc_kind_of_metric(metric);
  plan->two_level = true;
  const uint32_t lut_floats = (pq_m * pq_ksub + 1u) & ~1u;  // even: keeps the queue 8-byte aligned
  plan->lut_smem_floats = lut_floats <= kLutSmemMaxFloats ? lut_floats : 0;
  plan->aq_cap = aq_cap;
  plan->aq_smem_entries = aq_cap <= kAqSmemMaxEntries ? aq_cap : 0;
  plan->mode = 1;
  return plan_dispatch<1>(plan->acc, ef <= kEfSmemMax, ld, ef, u_cap, sms, plan);
}

isl_status plan_search_adc(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, uint32_t pq_m,
                           uint32_t pq_ksub, int sms, SearchPlan* plan) {
  plan->acc = acc_kind_of_metric(metric);
  plan->two_level = true;
  plan->mode = 2;
  const uint32_t lut_floats = (pq_m * pq_ksub + 1u) & ~1u;
  plan->lut_smem_floats = lut_floats <= kLutSmemMaxFloats ? lut_floats : 0;
  plan->aq_cap = 0;
  plan->aq_smem_entries = 0;
  return plan_dispatch<2>(plan->acc, ef <= kEfSmemMax, ld, ef, u_cap, sms, plan);
}

isl_status plan_search_adc_traverse(uint32_t ef, uint32_t u_cap, uint32_t pq_m, uint32_t pq_ksub, bool one_byte_codes, int sms,
                                    SearchPlan* plan) {
  plan->acc = ACC_DOT;
  plan->two_level = true;
  plan->mode = 3;
  const uint32_t lut_floats = (pq_m * pq_ksub + 1u) & ~1u;
  plan->lut_smem_floats = lut_floats <= kLutSmemMaxFloats ? lut_floats : 0;
  plan->aq_cap = 0;
  plan->aq_smem_entries = 0;
  // register bag: NR = ceil(ef / 32) entries per lane (rounded up to an instantiated size), up to ef = 512
  plan->nr = 0;
  if (one_byte_codes && (pq_m == 16 || pq_m == 32) && plan->lut_smem_floats && ef <= 512) {
    const int need = (int)((ef + 31) / 32);
    const int sizes[] = {2, 4, 6, 8, 12, 16};
    for (int sz : sizes)
      if (need <= sz) {
        plan->nr = sz;
        break;
      }
  }
  plan->ks = pq_ksub == 128 ? 128 : 0;  // the specialised table fold of adc_traverse.cuh
#define ISL_BAG(N) \
  case N:          \
    return plan->ks == 128 ? plan_bag<N, 128>(ef, sms, plan) : plan_bag<N, 0>(ef, sms, plan);
  switch (plan->nr) {
    ISL_BAG(2)
    ISL_BAG(4)
    ISL_BAG(6)
    ISL_BAG(8)
    ISL_BAG(12)
    ISL_BAG(16)
  }
#undef ISL_BAG
  return ef <= kEfSmemMax ? plan_lean<true>(ef, u_cap, pq_m, sms, plan) : plan_lean<false>(ef, u_cap, pq_m, sms, plan);
}

isl_status plan_search_rerank(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan) {
  plan->acc = acc_kind_of_metric(metric);
  plan->two_level = true;
  plan->mode = 2;
  plan->lut_smem_floats = 0;  // phase 2 never reads the PQ tables
  plan->aq_cap = 0;
  plan->aq_smem_entries = 0;
  return plan_dispatch<2>(plan->acc, ef <= kEfSmemMax, ld, ef, u_cap, sms, plan);
}

isl_status launch_search(const SearchPlan& plan, const SearchArgs& args, cudaStream_t st) {
  // Never launch more warps than queries: idle slots would only clear their bitsets.
  const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(plan.grid, args.nq));
  if (plan.mode == 3) {
#define ISL_BAG(N) \
  case N:          \
    return plan.ks == 128 ? launch_bag<N, 128>(plan, args, grid, st) : launch_bag<N, 0>(plan, args, grid, st);
    switch (plan.nr) {
      ISL_BAG(2)
      ISL_BAG(4)
      ISL_BAG(6)
      ISL_BAG(8)
      ISL_BAG(12)
      ISL_BAG(16)
    }
#undef ISL_BAG
    return plan.r_in_smem ? launch_lean<true>(plan, args, grid, st) : launch_lean<false>(plan, args, grid, st);
  }
  if (plan.mode == 2) return launch_dispatch<2>(plan, args, grid, st);
  if (plan.mode == 0 && plan.nr > 0) {
    switch (plan.acc) {
      case ACC_DOT: return launch_exact_bag_dot(plan, args, grid, st);
      case ACC_L2: return launch_exact_bag_l2(plan, args, grid, st);
      case ACC_L1: return launch_exact_bag_l1(plan, args, grid, st);
    }
    return fail(ISL_INVALID_CONFIG, "search: unknown metric");
  }
  return plan.mode == 1 ? launch_dispatch<1>(plan, args, grid, st) : launch_dispatch<0>(plan, args, grid, st);
}

}  // namespace isl
