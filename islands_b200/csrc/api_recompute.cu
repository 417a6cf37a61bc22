// api_recompute.cu — search with on-demand embedding recomputation (LEANN's storage-free mode):
// the index keeps graph + PQ codes + token rows, NOT the f32 embeddings.  A query batch runs
//   1. the ADC traversal (search_core.cuh MODE 2, phase 1): ef survivors per query, PQ bytes only;
//   2. the recompute step: the distinct survivors of the whole batch are compacted (flag, scan,
//      gather of their token rows) and pushed through the bf16 tcgen05 encoder once (encoder.cu);
//   3. the exact rerank (MODE 2, phase 2): reference-order distances (distance.rs) of every
//      survivor against the recomputed embeddings, sorted by (distance, id).
// The reference's seam for this is EmbeddingProvider::compute_embeddings_batch (leann.rs:82-99)
// called on the search frontier (leann.rs:947-950); batching the whole frontier of the batch into
// one encoder pass is docs/leann-specification.md:364-394 ("dynamic batching").
#include <cub/cub.cuh>

#include <algorithm>

#include "api_common.h"

namespace isl {
namespace {

__global__ void mark_survivors_kernel(const uint32_t* __restrict__ surv, const uint32_t* __restrict__ cnt, uint32_t nq,
                                      uint32_t ef, uint32_t* __restrict__ flags) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)nq * ef; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t q = (uint32_t)(i / ef), j = (uint32_t)(i % ef);
    if (j < cnt[q]) flags[surv[i]] = 1u;
  }
}

// One warp per flagged node: its token row goes to row rows[id] of the compact batch.
__global__ void gather_tokens_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ rows, uint32_t n,
                                     const int32_t* __restrict__ tokens, const int32_t* __restrict__ lengths, uint32_t S,
                                     int32_t* __restrict__ out_tok, int32_t* __restrict__ out_len) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
  for (uint32_t id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; id < n; id += warps) {
    if (!flags[id]) continue;
    const uint32_t r = rows[id];
    for (uint32_t i = lane; i < S; i += 32) out_tok[(size_t)r * S + i] = tokens[(size_t)id * S + i];
    if (lane == 0) out_len[r] = lengths[id];
  }
}

// Hub cache: survivors whose embedding is resident are not recomputed.
__global__ void clear_cached_flags_kernel(const uint32_t* __restrict__ hub_row, uint32_t n, uint32_t* __restrict__ flags,
                                          unsigned int* __restrict__ hits) {
  uint32_t local = 0;
  for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
    if (flags[id] && hub_row[id] != 0xffffffffu) {
      flags[id] = 0u;
      local++;
    }
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(hits, local);
}

// row of the rerank matrix per node id: cached hubs first ([0, hubs)), recomputed rows after them
__global__ void final_rows_kernel(const uint32_t* __restrict__ hub_row, uint32_t hubs, uint32_t n, uint32_t* __restrict__ rows) {
  for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
    const uint32_t h = hub_row[id];
    rows[id] = h != 0xffffffffu ? h : hubs + rows[id];
  }
}

__global__ void gather_rows_by_id_kernel(const uint64_t* __restrict__ ids, uint32_t count, const int32_t* __restrict__ tokens,
                                         const int32_t* __restrict__ lengths, uint32_t S, int32_t* __restrict__ out_tok,
                                         int32_t* __restrict__ out_len) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < count; r += warps) {
    const uint64_t id = ids[r];
    for (uint32_t i = lane; i < S; i += 32) out_tok[(size_t)r * S + i] = tokens[id * S + i];
    if (lane == 0) out_len[r] = lengths[id];
  }
}

}  // namespace
}  // namespace isl

using namespace isl;

extern "C" {

isl_status isl_index_set_recompute(isl_index* idx, isl_encoder* enc, const int32_t* token_ids, const int32_t* lengths,
                                   uint32_t seq_len) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  DeviceGuard g(idx->device);
  std::unique_lock<std::shared_mutex> lock(idx->mu);
  if (!enc) {  // detach
    if (idx->n && !idx->vectors.p)
      return fail(ISL_INVALID_ARGUMENT, "the stored vectors were dropped: detaching the encoder would leave the index "
                                        "without any source of embeddings");
    idx->encoder = nullptr;
    idx->node_tokens.release();
    idx->node_lengths.release();
    idx->tok_len = 0;
    idx->hub_count = 0;
    idx->hub_emb.release();
    idx->hub_sq.release();
    idx->hub_row.release();
    return ISL_OK;
  }
  if (idx->n && (!token_ids || !lengths)) return fail(ISL_INVALID_ARGUMENT, "token_ids / lengths is null");
  if (seq_len == 0) return fail(ISL_INVALID_ARGUMENT, "seq_len must be > 0");
  if (idx->n && isl_encoder_dimension(enc) != idx->dim)  // EmbeddingProvider::dimension (leann.rs:97-98)
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                      std::to_string(isl_encoder_dimension(enc)));
  if (idx->n && idx->ld != idx->dim) return fail(ISL_INVALID_ARGUMENT, "recompute needs dim % 4 == 0");
  ISL_CUDA_TRY(idx->node_tokens.alloc(std::max<uint64_t>(idx->n * seq_len, 1)));
  ISL_CUDA_TRY(idx->node_lengths.alloc(std::max<uint64_t>(idx->n, 1)));
  if (idx->n) {
    ISL_CUDA_TRY(cudaMemcpy(idx->node_tokens.p, token_ids, idx->n * seq_len * 4, cudaMemcpyHostToDevice));
    ISL_CUDA_TRY(cudaMemcpy(idx->node_lengths.p, lengths, idx->n * 4, cudaMemcpyHostToDevice));
  }
  idx->encoder = enc;
  idx->tok_len = seq_len;
  idx->hub_count = 0;  // a cache built with another provider is void
  idx->hub_emb.release();
  idx->hub_sq.release();
  idx->hub_row.release();
  return ISL_OK;
} ISL_ABI_GUARD

// Hub-embedding cache (docs/leann-specification.md:661-690): the `count` nodes with the highest in-degree
// (ties: smaller id) keep their embedding resident — they are the rows a traversal reaches most often —
// and the recompute search skips them.  Computed with the attached encoder, so cached and recomputed
// rows are the same bits.  count == 0 drops the cache.
isl_status isl_index_set_hub_cache(isl_index* idx, uint64_t count) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  DeviceGuard g(idx->device);
  std::unique_lock<std::shared_mutex> lock(idx->mu);
  if (!idx->encoder) return fail(ISL_INVALID_ARGUMENT, "no recompute encoder attached (isl_index_set_recompute)");
  idx->hub_count = 0;
  idx->hub_emb.release();
  idx->hub_sq.release();
  idx->hub_row.release();
  const uint64_t n = idx->n;
  count = std::min<uint64_t>(count, n);
  if (count == 0) return ISL_OK;
  std::vector<uint32_t> indeg(n, 0);
  for (uint64_t v : idx->h_nbrs)
    if (v < n) indeg[v]++;
  std::vector<uint64_t> order(n);
  for (uint64_t i = 0; i < n; ++i) order[i] = i;
  auto hotter = [&](uint64_t a, uint64_t b) { return indeg[a] != indeg[b] ? indeg[a] > indeg[b] : a < b; };
  std::nth_element(order.begin(), order.begin() + (count - 1), order.end(), hotter);
  order.resize(count);
  std::sort(order.begin(), order.end());  // cache rows in ascending id order
  std::vector<uint32_t> h_row(n, 0xffffffffu);
  for (uint64_t r = 0; r < count; ++r) h_row[order[r]] = (uint32_t)r;
  const uint32_t S = idx->tok_len;
  DevBuf<uint64_t> d_ids;
  DevBuf<int32_t> tok, len;
  ISL_CUDA_TRY(d_ids.alloc(count));
  ISL_CUDA_TRY(tok.alloc(count * S));
  ISL_CUDA_TRY(len.alloc(count));
  ISL_CUDA_TRY(idx->hub_row.alloc(n));
  ISL_CUDA_TRY(idx->hub_emb.alloc(count * idx->ld));
  ISL_CUDA_TRY(idx->hub_sq.alloc(count));
  cudaStream_t st = idx->stream;
  ISL_CUDA_TRY(cudaMemcpyAsync(d_ids.p, order.data(), count * 8, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(idx->hub_row.p, h_row.data(), n * 4, cudaMemcpyHostToDevice, st));
  gather_rows_by_id_kernel<<<1184, 256, 0, st>>>(d_ids.p, (uint32_t)count, idx->node_tokens.p, idx->node_lengths.p, S, tok.p, len.p);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  ISL_TRY(isl_encoder_embed_dev(idx->encoder, tok.p, len.p, count, S, idx->hub_emb.p));
  ISL_TRY(launch_row_sqnorms(idx->hub_emb.p, count, idx->dim, idx->ld, idx->hub_sq.p, idx->sms, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  idx->hub_count = count;
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_index_hub_cache_info(const isl_index* idx, uint64_t* cached_nodes, uint64_t* last_hits) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (cached_nodes) *cached_nodes = idx->hub_count;
  if (last_hits) *last_hits = idx->last_hub_hits;
  return ISL_OK;
} ISL_ABI_GUARD

// Frees the resident f32 embeddings: afterwards only the recompute search works on this handle.
isl_status isl_index_drop_vectors(isl_index* idx) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  DeviceGuard g(idx->device);
  std::unique_lock<std::shared_mutex> lock(idx->mu);
  if (!idx->encoder) return fail(ISL_INVALID_ARGUMENT, "attach a recompute encoder before dropping the stored vectors");
  idx->vectors.release();
  idx->sqnorms.release();
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"

namespace isl {
isl_status adc_recompute_on_scratch(const isl_index* idx, SearchScratch* sc, const float* queries, uint64_t nq, uint32_t k,
                                    uint32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                    isl_search_stats* stats, const ShardOut* shard) {
  if (!idx->pq) return fail(ISL_PQ_ERROR, "no product quantizer attached (isl_index_attach_pq)");
  if (!idx->encoder) return fail(ISL_INVALID_ARGUMENT, "no recompute encoder attached (isl_index_set_recompute)");
  if (!shard && (!out_ids || !out_dist)) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  const isl_pq* pq = idx->pq;
  const uint32_t m = (uint32_t)pq->cfg.num_subquantizers;
  const uint32_t lut_floats = m * pq->ksub;
  const uint32_t n = (uint32_t)idx->n;
  const uint32_t maxdeg = std::max<uint32_t>(idx->max_degree, 1);
  const uint32_t u_cap_t = std::max<uint32_t>(32, round_up(maxdeg + 1, 32));
  const uint32_t u_cap_r = std::max<uint32_t>(32, round_up(ef, 32));
  SearchPlan plan, plan_r;  // lean traversal (MODE 3) and rerank (MODE 2 / phase 2)
  ISL_TRY(plan_search_adc_traverse(ef, u_cap_t, m, pq->ksub, idx->codes8.p != nullptr, idx->sms, &plan));
  ISL_TRY(plan_search_rerank(idx->cfg.metric, idx->ld, ef, u_cap_r, idx->sms, &plan_r));
  const uint32_t vis_words = round_up((uint32_t)((idx->n + 31) / 32), 4);
  const uint32_t slots = (uint32_t)std::min<uint64_t>(std::max(plan.grid, plan_r.grid), nq);
  ISL_TRY(ensure(sc->visited, (size_t)slots * vis_words));
  if (!plan.r_in_smem || !plan_r.r_in_smem) ISL_TRY(ensure(sc->r_global, (size_t)slots * ef));
  ISL_TRY(ensure(sc->ties_global, (size_t)slots * ef));
  ISL_TRY(ensure(sc->aux_f32, (size_t)nq * lut_floats + 2));
  ISL_TRY(ensure(sc->q_stage, nq * idx->ld));
  ISL_TRY(ensure(sc->out_ids, nq * k));
  ISL_TRY(ensure(sc->out_dist, nq * k));
  ISL_TRY(ensure(sc->out_count, nq));
  ISL_TRY(ensure(sc->out_stats, nq));
  ISL_TRY(ensure(sc->rc_surv, nq * (size_t)ef));
  ISL_TRY(ensure(sc->rc_surv_cnt, nq));
  ISL_TRY(ensure(sc->rc_flags, (size_t)n + 1));
  ISL_TRY(ensure(sc->rc_rows, (size_t)n + 1));
  cudaStream_t st = sc->stream;
  if (idx->ld != idx->dim) ISL_CUDA_TRY(cudaMemsetAsync(sc->q_stage.p, 0, nq * idx->ld * 4, st));
  ISL_CUDA_TRY(cudaMemcpy2DAsync(sc->q_stage.p, (size_t)idx->ld * 4, queries, (size_t)idx->dim * 4, (size_t)idx->dim * 4,
                                 nq, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, 4 * sizeof(unsigned int), st));

  // ---- 1. ADC traversal ------------------------------------------------------------------------
  ISL_CUDA_TRY(cudaEventRecord(sc->ev0, st));
  const bool fused_lut = plan.lut_smem_floats != 0;  // tables built per query inside the traversal kernel
  if (!fused_lut) ISL_TRY(launch_pq_tables(pq->dev(), sc->q_stage.p, idx->ld, nq, sc->aux_f32.p, idx->sms, st));
  SearchArgs a{};
  a.vectors = nullptr;  // phase 1 never reads an embedding
  a.sqnorms = nullptr;
  a.ld = idx->ld;
  a.d = idx->dim;
  a.n = n;
  search_args_set_graph(idx, &a);
  a.queries = sc->q_stage.p;
  a.q_ld = idx->ld;
  a.nq = (uint32_t)nq;
  a.entry = (uint32_t)idx->entry;
  a.k = k;
  a.ef = ef;
  a.metric = idx->cfg.metric;
  a.visited = sc->visited.p;
  a.vis_words = vis_words;
  a.r_global = sc->r_global.p;
  a.ties_global = sc->ties_global.p;
  a.u_cap = u_cap_t;
  a.out_ids = sc->out_ids.p;
  a.out_dist = sc->out_dist.p;
  a.out_count = sc->out_count.p;
  a.stats = sc->out_stats.p;
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
  a.lut_smem_floats = plan.lut_smem_floats;
  a.phase = 1;
  a.rerank_limit = idx->rerank_limit ? std::max(idx->rerank_limit, k) : 0u;
  a.surv_ids = sc->rc_surv.p;
  a.surv_cnt = sc->rc_surv_cnt.p;
  ISL_TRY(launch_search(plan, a, st));
  ISL_CUDA_TRY(cudaEventRecord(sc->ev1, st));

  // ---- 2. recompute the distinct survivors ----------------------------------------------------------
  ISL_CUDA_TRY(cudaMemsetAsync(sc->rc_flags.p, 0, ((size_t)n + 1) * 4, st));
  mark_survivors_kernel<<<1184, 256, 0, st>>>(sc->rc_surv.p, sc->rc_surv_cnt.p, (uint32_t)nq, ef, sc->rc_flags.p);
  count_launch();
  const uint32_t hubs = (uint32_t)idx->hub_count;
  if (hubs) {  // survivors with a resident embedding are not recomputed
    ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p + 2, 0, sizeof(unsigned int), st));
    clear_cached_flags_kernel<<<1184, 256, 0, st>>>(idx->hub_row.p, n, sc->rc_flags.p, sc->counters.p + 2);
    count_launch();
  }
  size_t scan_bytes = 0;
  ISL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, sc->rc_flags.p, sc->rc_rows.p, (int)(n + 1), st));
  ISL_TRY(ensure(sc->rc_tmp, scan_bytes + 16));
  ISL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(sc->rc_tmp.p, scan_bytes, sc->rc_flags.p, sc->rc_rows.p, (int)(n + 1), st));
  count_launch(2);
  uint32_t unique = 0, hub_hits = 0;
  ISL_CUDA_TRY(cudaMemcpyAsync(&unique, sc->rc_rows.p + n, 4, cudaMemcpyDeviceToHost, st));
  if (hubs) ISL_CUDA_TRY(cudaMemcpyAsync(&hub_hits, sc->counters.p + 2, 4, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  float traverse_ms = 0.0f, encoder_ms = 0.0f;
  cudaEventElapsedTime(&traverse_ms, sc->ev0, sc->ev1);
  const uint32_t S = idx->tok_len;
  ISL_TRY(ensure(sc->rc_tok, (size_t)unique * S + 1));
  ISL_TRY(ensure(sc->rc_len, (size_t)unique + 1));
  // rerank matrix: [cached hub rows | rows recomputed for this batch]
  ISL_TRY(ensure(sc->rc_emb, ((size_t)hubs + unique) * idx->ld + 4));
  ISL_TRY(ensure(sc->rc_sq, (size_t)hubs + unique + 1));
  if (hubs) {
    ISL_CUDA_TRY(cudaMemcpyAsync(sc->rc_emb.p, idx->hub_emb.p, (size_t)hubs * idx->ld * 4, cudaMemcpyDeviceToDevice, st));
    ISL_CUDA_TRY(cudaMemcpyAsync(sc->rc_sq.p, idx->hub_sq.p, (size_t)hubs * 4, cudaMemcpyDeviceToDevice, st));
  }
  gather_tokens_kernel<<<1184, 256, 0, st>>>(sc->rc_flags.p, sc->rc_rows.p, n, idx->node_tokens.p, idx->node_lengths.p, S,
                                            sc->rc_tok.p, sc->rc_len.p);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  if (unique) {
    ISL_TRY(isl_encoder_embed_dev(idx->encoder, sc->rc_tok.p, sc->rc_len.p, unique, S, sc->rc_emb.p + (size_t)hubs * idx->ld));
    isl_encoder_last_timing(idx->encoder, &encoder_ms, nullptr);
    ISL_TRY(launch_row_sqnorms(sc->rc_emb.p + (size_t)hubs * idx->ld, unique, idx->dim, idx->ld, sc->rc_sq.p + hubs, idx->sms, st));
  }
  if (hubs) {
    final_rows_kernel<<<1184, 256, 0, st>>>(idx->hub_row.p, hubs, n, sc->rc_rows.p);
    count_launch();
  }

  // ---- 3. exact rerank against the recomputed rows ---------------------------------------------------
  ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, sizeof(unsigned int), st));
  a.vectors = sc->rc_emb.p;
  a.sqnorms = sc->rc_sq.p;
  a.row_of_id = sc->rc_rows.p;
  a.phase = 2;
  a.u_cap = u_cap_r;
  a.lut_smem_floats = 0;
  if (shard) a.shard = *shard;
  ISL_CUDA_TRY(cudaEventRecord(sc->ev0, st));
  ISL_TRY(launch_search(plan_r, a, st));
  ISL_CUDA_TRY(cudaEventRecord(sc->ev1, st));
  {
    std::lock_guard<std::mutex> tl(idx->pool_mu);
    idx->last_traverse_ms = traverse_ms;
    idx->last_encoder_ms = encoder_ms;
    idx->last_recomputed = unique;
    idx->last_hub_hits = hub_hits;
  }
  if (shard) return ISL_OK;  // records written; exchange, merge and the final synchronisation follow in api_shard.cu

  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, sc->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, sc->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, sc->out_count.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (stats)
    ISL_CUDA_TRY(cudaMemcpyAsync(stats, sc->out_stats.p, nq * sizeof(isl_search_stats), cudaMemcpyDeviceToHost, st));
  ISL_TRY(search_finish(idx, sc, 3));
  {
    std::lock_guard<std::mutex> tl(idx->pool_mu);
    idx->last_rerank_ms = sc->kernel_ms;
    idx->last_traverse_ms = traverse_ms;
    idx->last_encoder_ms = encoder_ms;
    idx->last_recomputed = unique;
    idx->last_hub_hits = hub_hits;
  }
  return ISL_OK;
}

}  // namespace isl

extern "C" {

isl_status isl_index_search_adc_recompute(const isl_index* idx, const float* queries, uint64_t nq, uint32_t query_dim,
                                          uint32_t k, uint32_t ef, uint64_t* out_ids, float* out_dist,
                                          uint32_t* out_count, isl_search_stats* stats) try {
  bool trivial;
  ISL_TRY(search_checks(idx, queries, nq, query_dim, k, &ef, &trivial, /*need_vectors=*/false));
  if (trivial) {
    fill_empty(nq, k, out_ids, out_dist, out_count, stats);
    return ISL_OK;
  }
  DeviceGuard g(idx->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  return adc_recompute_on_scratch(idx, sc.get(), queries, nq, k, ef, out_ids, out_dist, out_count, stats, nullptr);
} ISL_ABI_GUARD

isl_status isl_index_last_recompute(const isl_index* idx, uint64_t* unique_nodes, float* traverse_ms, float* encoder_ms,
                                    float* rerank_ms) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (unique_nodes) *unique_nodes = idx->last_recomputed;
  if (traverse_ms) *traverse_ms = idx->last_traverse_ms;
  if (encoder_ms) *encoder_ms = idx->last_encoder_ms;
  if (rerank_ms) *rerank_ms = idx->last_rerank_ms;
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
