// search.cu — launch planning for the batched best-first search kernel (search_core.cuh).
#include "search.h"

#include <algorithm>

#include "search_core.cuh"

namespace isl {

namespace {
constexpr int kCH = 64;
constexpr int kStages = 3;
// R (8 B per entry) stays in shared memory up to this many entries; above, it lives in an
// L2-resident global buffer so that enough warps stay resident per SM.
constexpr uint32_t kEfSmemMax = 2048;

template <int ACC, bool R_SMEM>
isl_status plan_one(uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan) {
  auto kern = leann_search_kernel<ACC, kCH, kStages, R_SMEM>;
  const size_t smem = search_smem_bytes<kCH, kStages>(ld, R_SMEM ? ef : 0, u_cap);
  if (smem > 227 * 1024)
    return fail(ISL_INVALID_ARGUMENT, "search: dimension / ef need more than 227 KB of shared memory per warp");
  ISL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  ISL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) return fail(ISL_CUDA_ERROR, "search: kernel does not fit on an SM");
  plan->smem = smem;
  plan->ctas_per_sm = per_sm;
  plan->grid = (uint32_t)(per_sm * sms);
  plan->r_in_smem = R_SMEM;
  return ISL_OK;
}

template <int ACC, bool R_SMEM>
isl_status launch_one(const SearchPlan& plan, const SearchArgs& args, uint32_t grid, cudaStream_t st) {
  leann_search_kernel<ACC, kCH, kStages, R_SMEM><<<grid, 32, plan.smem, st>>>(args);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}
}  // namespace

isl_status plan_search(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan) {
  const bool r_smem = ef <= kEfSmemMax;
  const int acc = acc_kind_of_metric(metric);
  plan->acc = acc;
#define ISL_PLAN(A)                                                          \
  case A:                                                                    \
    return r_smem ? plan_one<A, true>(ld, ef, u_cap, sms, plan)              \
                  : plan_one<A, false>(ld, ef, u_cap, sms, plan);
  switch (acc) {
    ISL_PLAN(ACC_DOT)
    ISL_PLAN(ACC_L2)
    ISL_PLAN(ACC_L1)
  }
#undef ISL_PLAN
  return fail(ISL_INVALID_CONFIG, "search: unknown metric");
}

isl_status launch_search(const SearchPlan& plan, const SearchArgs& args, cudaStream_t st) {
  // Never launch more warps than queries: idle slots would only clear their bitsets.
  const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(plan.grid, args.nq));
#define ISL_LAUNCH(A)                                                        \
  case A:                                                                    \
    return plan.r_in_smem ? launch_one<A, true>(plan, args, grid, st)        \
                          : launch_one<A, false>(plan, args, grid, st);
  switch (plan.acc) {
    ISL_LAUNCH(ACC_DOT)
    ISL_LAUNCH(ACC_L2)
    ISL_LAUNCH(ACC_L1)
  }
#undef ISL_LAUNCH
  return fail(ISL_INVALID_CONFIG, "search: unknown metric");
}

}  // namespace isl
