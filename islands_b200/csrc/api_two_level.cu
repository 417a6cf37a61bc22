// api_two_level.cu — two-level search: PQ ADC traversal + exact rerank.
// The reference has no code for this composition (src/core/leann.rs:54-56 says so); it is
// defined from docs/leann-specification.md:223-269 and the PQ primitives pq.rs:307-348, and
// restated on the CPU in oracle/orc_leann_search_two_level.  Parity is oracle<->GPU only.
#include <algorithm>
#include <cmath>

#include "api_common.h"

using namespace isl;

extern "C" {

isl_status isl_index_attach_pq(isl_index* idx, const isl_pq* pq, const uint16_t* codes) try {
  if (!idx || !pq) return fail(ISL_INVALID_ARGUMENT, "null handle");
  if (!pq->trained) return fail(ISL_PQ_ERROR, "Quantizer not trained");
  if (pq->dim != idx->dim && idx->n)
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                      std::to_string(pq->dim));
  if (idx->n && !codes) return fail(ISL_INVALID_ARGUMENT, "codes is null");
  DeviceGuard g(idx->device);
  std::unique_lock<std::shared_mutex> lock(idx->mu);
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  const uint64_t total = idx->n * m;
  for (uint64_t i = 0; i < total; ++i)
    if (codes[i] >= pq->ksub) return fail(ISL_PQ_ERROR, "Invalid code " + std::to_string(codes[i]));
  idx->codes8.release();
  idx->codes16.release();
  if (pq->ksub <= 256) {  // PQConfig::bytes_per_vector (pq.rs:58-64): one byte per code
    std::vector<uint8_t> c8(total);
    for (uint64_t i = 0; i < total; ++i) c8[i] = (uint8_t)codes[i];
    ISL_CUDA_TRY(idx->codes8.alloc(std::max<uint64_t>(total, 1)));
    if (total) ISL_CUDA_TRY(cudaMemcpy(idx->codes8.p, c8.data(), total, cudaMemcpyHostToDevice));
  } else {
    ISL_CUDA_TRY(idx->codes16.alloc(std::max<uint64_t>(total, 1)));
    if (total) ISL_CUDA_TRY(cudaMemcpy(idx->codes16.p, codes, total * 2, cudaMemcpyHostToDevice));
  }
  idx->pq = pq;
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"

namespace isl {
// The PQ searches on a leased scratch; the caller holds the device guard and the handle's shared lock and has run
// search_checks.  shard != null (mode 2 only): the rerank launch also writes the shard-exchange records, the host
// copies of the results and the final synchronisation are left to the caller (api_shard.cu).
isl_status pq_search_on_scratch(int mode, const isl_index* idx, SearchScratch* sc, const float* queries, uint64_t nq, uint32_t k,
                                uint32_t ef, float rerank_ratio, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                isl_search_stats* stats, const ShardOut* shard) {
  if (!idx->pq) return fail(ISL_PQ_ERROR, "no product quantizer attached (isl_index_attach_pq)");
  if (mode == 1 && (!(rerank_ratio > 0.0f) || rerank_ratio > 1.0f))
    return fail(ISL_INVALID_ARGUMENT, "rerank_ratio must be in (0, 1]");
  if (!shard && (!out_ids || !out_dist)) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  const isl_pq* pq = idx->pq;
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  const uint32_t lut_floats = m * pq->ksub;

  if (mode == 2) {
    // Two launches: the lean ADC traversal (MODE 3: no staging ring or query vector in shared memory,
    // result array in registers up to ef = 512) hands its ef survivors to the exact rerank (MODE 2,
    // phase 2).  Each half gets the occupancy its own shared-memory footprint allows.
    const uint32_t maxdeg2 = std::max<uint32_t>(idx->max_degree, 1);
    // the vectorised hop of the lean kernel (one-byte codes, m = 16 / 32, table in shared memory) keeps its list in
    // registers: only the generic hop needs room for a whole neighbour list in shared memory
    const bool vector_hop = idx->codes8.p && (m == 16 || m == 32) && lut_floats <= 8192;
    const uint32_t u_cap_t = vector_hop ? 32u : std::max<uint32_t>(32, round_up(maxdeg2 + 1, 32));
    const uint32_t u_cap_r = std::max<uint32_t>(32, round_up(ef, 32));
    SearchPlan pt, pr;
    ISL_TRY(plan_search_adc_traverse(ef, u_cap_t, m, pq->ksub, idx->codes8.p != nullptr, idx->sms, &pt));
    ISL_TRY(plan_search_rerank(idx->cfg.metric, idx->ld, ef, u_cap_r, idx->sms, &pr));
    const uint32_t vis_words2 = round_up((uint32_t)((idx->n + 31) / 32), 4);
    const uint32_t slots_t = (uint32_t)std::min<uint64_t>(pt.grid, nq), slots_r = (uint32_t)std::min<uint64_t>(pr.grid, nq);
    const bool novis = !stats && pt.novis_ok && idx->codes8.p && idx->n < kIdcMaxNodes;
    if (!novis) ISL_TRY(ensure(sc->visited, (size_t)slots_t * vis_words2));  // the bitset-free traversal needs no scratch
    if (!pt.r_in_smem || !pr.r_in_smem) ISL_TRY(ensure(sc->r_global, (size_t)std::max(slots_t, slots_r) * ef));
    ISL_TRY(ensure(sc->ties_global, (size_t)std::max(slots_t, slots_r) * ef));
    ISL_TRY(ensure(sc->aux_f32, (size_t)nq * lut_floats + 2));
    ISL_TRY(ensure(sc->q_stage, nq * idx->ld));
    ISL_TRY(ensure(sc->out_ids, nq * k));
    ISL_TRY(ensure(sc->out_dist, nq * k));
    ISL_TRY(ensure(sc->out_count, nq));
    ISL_TRY(ensure(sc->out_stats, nq));
    ISL_TRY(ensure(sc->rc_surv, nq * (size_t)ef));
    ISL_TRY(ensure(sc->rc_surv_cnt, nq));
    cudaStream_t st = sc->stream;
    if (idx->ld != idx->dim) ISL_CUDA_TRY(cudaMemsetAsync(sc->q_stage.p, 0, nq * idx->ld * 4, st));
    ISL_CUDA_TRY(cudaMemcpy2DAsync(sc->q_stage.p, (size_t)idx->ld * 4, queries, (size_t)idx->dim * 4,
                                   (size_t)idx->dim * 4, nq, cudaMemcpyHostToDevice, st));
    ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, 4 * sizeof(unsigned int), st));
    ISL_CUDA_TRY(cudaEventRecord(sc->ev0, st));
    const bool fused_lut = pt.lut_smem_floats != 0;  // tables built per query inside the traversal kernel
    if (!fused_lut) ISL_TRY(launch_pq_tables(pq->dev(), sc->q_stage.p, idx->ld, nq, sc->aux_f32.p, idx->sms, st));
    SearchArgs a{};
    a.vectors = idx->vectors.p;
    a.sqnorms = idx->sqnorms.p;
    a.ld = idx->ld;
    a.d = idx->dim;
    a.n = (uint32_t)idx->n;
    search_args_set_graph(idx, &a);
    a.queries = sc->q_stage.p;
    a.q_ld = idx->ld;
    a.nq = (uint32_t)nq;
    a.entry = (uint32_t)idx->entry;
    a.k = k;
    a.ef = ef;
    a.metric = idx->cfg.metric;
    a.visited = sc->visited.p;
    a.vis_words = vis_words2;
    a.r_global = sc->r_global.p;
    a.ties_global = sc->ties_global.p;
    a.u_cap = u_cap_t;
    a.out_ids = sc->out_ids.p;
    a.out_dist = sc->out_dist.p;
    a.out_count = sc->out_count.p;
    a.stats = stats ? sc->out_stats.p : nullptr;
    // without statistics the traversal runs without the visited bitset (same survivors; search_core.cuh)
    a.novis = novis ? 1u : 0u;
    a.work_counter = sc->counters.p;
    a.error_flag = sc->counters.p + 1;
    a.luts = fused_lut ? nullptr : sc->aux_f32.p;
    a.pq_codebooks = pq->d_codebooks.p;
    a.pq_dsub = pq->dsub;
    a.pq_ld_sub = pq->ld_sub;
    a.codes8 = idx->codes8.p;
    a.codes16 = idx->codes8.p ? nullptr : idx->codes16.p;
    a.pq_m = m;
    a.pq_ksub = pq->ksub;
    a.lut_smem_floats = pt.lut_smem_floats;
    a.phase = 1;
    a.rerank_limit = idx->rerank_limit ? std::max(idx->rerank_limit, k) : 0u;
    a.surv_ids = sc->rc_surv.p;
    a.surv_cnt = sc->rc_surv_cnt.p;
    ISL_TRY(launch_search(pt, a, st));
    ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, sizeof(unsigned int), st));  // work counter of the second launch
    a.u_cap = u_cap_r;
    a.lut_smem_floats = 0;
    a.phase = 2;
    if (shard) a.shard = *shard;
    ISL_TRY(launch_search(pr, a, st));
    ISL_CUDA_TRY(cudaEventRecord(sc->ev1, st));
    if (shard) return ISL_OK;  // records written; exchange, merge and the final synchronisation follow in api_shard.cu
    ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, sc->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
    ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, sc->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, sc->out_count.p, nq * 4, cudaMemcpyDeviceToHost, st));
    if (stats)
      ISL_CUDA_TRY(cudaMemcpyAsync(stats, sc->out_stats.p, nq * sizeof(isl_search_stats), cudaMemcpyDeviceToHost, st));
    return search_finish(idx, sc, 3);
  }

  // |AQ| <= max_degree / a at all times (it gains at most max_degree entries per expansion and
  // then loses ceil(a * |AQ|)); one spare entry per insertion batch keeps the bound simple.
  const uint32_t maxdeg = std::max<uint32_t>(idx->max_degree, 1);
  uint32_t aq_cap = 0;
  uint32_t u_cap = std::max<uint32_t>(32, round_up(maxdeg + 1, 32));
  SearchPlan plan;
  if (mode == 1) {
    const uint64_t aq_cap64 = (uint64_t)std::ceil((double)maxdeg / (double)rerank_ratio) + maxdeg + 2;
    if (aq_cap64 > (1u << 22)) return fail(ISL_INVALID_ARGUMENT, "rerank_ratio too small for this graph degree");
    aq_cap = (uint32_t)aq_cap64;
    ISL_TRY(plan_search_two_level(idx->cfg.metric, idx->ld, ef, u_cap, m, pq->ksub, aq_cap, idx->sms, &plan));
  } else {
    u_cap = std::max<uint32_t>(u_cap, round_up(ef, 32));  // the rerank list holds the ef survivors
    ISL_TRY(plan_search_adc(idx->cfg.metric, idx->ld, ef, u_cap, m, pq->ksub, idx->sms, &plan));
  }
  const uint32_t vis_words = round_up((uint32_t)((idx->n + 31) / 32), 4);
  const uint32_t slots = (uint32_t)std::min<uint64_t>(plan.grid, nq);
  ISL_TRY(ensure(sc->visited, (size_t)slots * vis_words));
  if (!plan.r_in_smem) ISL_TRY(ensure(sc->r_global, (size_t)slots * ef));
  ISL_TRY(ensure(sc->ties_global, (size_t)slots * ef));
  if (mode == 1 && !plan.aq_smem_entries) ISL_TRY(ensure(sc->aux_u2, (size_t)slots * aq_cap));
  ISL_TRY(ensure(sc->aux_f32, (size_t)nq * lut_floats + 2));
  ISL_TRY(ensure(sc->q_stage, nq * idx->ld));
  ISL_TRY(ensure(sc->out_ids, nq * k));
  ISL_TRY(ensure(sc->out_dist, nq * k));
  ISL_TRY(ensure(sc->out_count, nq));
  if (stats) ISL_TRY(ensure(sc->out_stats, nq));
  cudaStream_t st = sc->stream;
  if (idx->ld != idx->dim) ISL_CUDA_TRY(cudaMemsetAsync(sc->q_stage.p, 0, nq * idx->ld * 4, st));
  ISL_CUDA_TRY(cudaMemcpy2DAsync(sc->q_stage.p, (size_t)idx->ld * 4, queries, (size_t)idx->dim * 4,
                                 (size_t)idx->dim * 4, nq, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, 4 * sizeof(unsigned int), st));

  ISL_CUDA_TRY(cudaEventRecord(sc->ev0, st));
  // K3: per-query tables LUT[j][c] = Σ (q - c)^2 (pq.rs:307-338), then the traversal.
  ISL_TRY(launch_pq_tables(pq->dev(), sc->q_stage.p, idx->ld, nq, sc->aux_f32.p, idx->sms, st));

  SearchArgs a{};
  a.vectors = idx->vectors.p;
  a.sqnorms = idx->sqnorms.p;
  a.ld = idx->ld;
  a.d = idx->dim;
  a.n = (uint32_t)idx->n;
  search_args_set_graph(idx, &a);
  a.queries = sc->q_stage.p;
  a.q_ld = idx->ld;
  a.nq = (uint32_t)nq;
  a.entry = (uint32_t)idx->entry;
  a.k = k;
  a.ef = ef;
  a.metric = idx->cfg.metric;
  a.prune_ratio = 0.0f;  // the PQ queue replaces the frontier pruning strategies (leann.rs:351-353)
  a.strategy = 0;
  a.visited = sc->visited.p;
  a.vis_words = vis_words;
  a.r_global = sc->r_global.p;
  a.ties_global = sc->ties_global.p;
  a.u_cap = u_cap;
  a.out_ids = sc->out_ids.p;
  a.out_dist = sc->out_dist.p;
  a.out_count = sc->out_count.p;
  a.stats = stats ? sc->out_stats.p : nullptr;
  a.work_counter = sc->counters.p;
  a.error_flag = sc->counters.p + 1;
  a.luts = sc->aux_f32.p;
  a.codes8 = idx->codes8.p;
  a.codes16 = idx->codes8.p ? nullptr : idx->codes16.p;
  a.pq_m = m;
  a.pq_ksub = pq->ksub;
  a.rerank_ratio = rerank_ratio;
  a.aq_cap = aq_cap;
  a.aq_global = sc->aux_u2.p;
  a.lut_smem_floats = plan.lut_smem_floats;
  a.aq_smem_entries = plan.aq_smem_entries;
  ISL_TRY(launch_search(plan, a, st));
  ISL_CUDA_TRY(cudaEventRecord(sc->ev1, st));

  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, sc->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, sc->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, sc->out_count.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (stats)
    ISL_CUDA_TRY(cudaMemcpyAsync(stats, sc->out_stats.p, nq * sizeof(isl_search_stats), cudaMemcpyDeviceToHost, st));
  return search_finish(idx, sc, 2);
}

}  // namespace isl

extern "C" {

static isl_status pq_search_common(int mode, const isl_index* idx, const float* queries, uint64_t nq, uint32_t query_dim,
                                   uint32_t k, uint32_t ef, float rerank_ratio, uint64_t* out_ids, float* out_dist,
                                   uint32_t* out_count, isl_search_stats* stats) {
  bool trivial;
  ISL_TRY(search_checks(idx, queries, nq, query_dim, k, &ef, &trivial));
  if (trivial) {
    fill_empty(nq, k, out_ids, out_dist, out_count, stats);
    return ISL_OK;
  }
  DeviceGuard g(idx->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  return pq_search_on_scratch(mode, idx, sc.get(), queries, nq, k, ef, rerank_ratio, out_ids, out_dist, out_count, stats, nullptr);
}

isl_status isl_index_set_rerank_limit(isl_index* idx, uint32_t limit) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  std::unique_lock<std::shared_mutex> lock(idx->mu);
  idx->rerank_limit = limit;
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_index_search_two_level(const isl_index* idx, const float* queries, uint64_t nq,
                                      uint32_t query_dim, uint32_t k, uint32_t ef, float rerank_ratio,
                                      uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                      isl_search_stats* stats) try {
  return pq_search_common(1, idx, queries, nq, query_dim, k, ef, rerank_ratio, out_ids, out_dist, out_count, stats);
} ISL_ABI_GUARD

isl_status isl_index_search_adc_rerank(const isl_index* idx, const float* queries, uint64_t nq,
                                       uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                       float* out_dist, uint32_t* out_count, isl_search_stats* stats) try {
  return pq_search_common(2, idx, queries, nq, query_dim, k, ef, 0.0f, out_ids, out_dist, out_count, stats);
} ISL_ABI_GUARD

}  // extern "C"
