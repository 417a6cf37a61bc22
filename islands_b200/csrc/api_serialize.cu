// api_serialize.cu — LeannIndex / ProductQuantizer to_bytes / from_bytes in the reference's byte
// layout (bincode.h).  LeannIndex serialises "only graph structure, not embeddings"
// (leann.rs:1058): from_bytes therefore takes the embeddings the provider would return.
#include <memory>

#include "api_common.h"
#include "bincode.h"

using namespace isl;

namespace {

isl_status emit(const ByteWriter& w, uint8_t* out, uint64_t cap, uint64_t* out_len) {
  if (out_len) *out_len = w.buf.size();
  if (!out) return ISL_OK;  // length query
  if (cap < w.buf.size()) return fail(ISL_INVALID_ARGUMENT, "output buffer too small: need " + std::to_string(w.buf.size()) + " bytes");
  std::memcpy(out, w.buf.data(), w.buf.size());
  return ISL_OK;
}

}  // namespace

extern "C" {

// LeannIndex { config: LeannConfig, graph: CsrGraph, dimension: Option<usize> } (leann.rs:493-500)
isl_status isl_index_to_bytes(const isl_index* idx, uint8_t* out, uint64_t cap, uint64_t* out_len) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  ByteWriter w;
  const isl_leann_config& c = idx->cfg;  // LeannConfig (leann.rs:322-371), declaration order
  w.u64(c.m);
  w.u64(c.m0);
  w.u64(c.ef_construction);
  w.f64(c.ml);
  w.u64(c.max_layers);
  w.u32((uint32_t)c.metric);
  w.u64(c.ef_search);
  w.u64(c.beam_width);
  w.f32(c.prune_ratio);
  w.u32((uint32_t)c.pruning_strategy);
  w.boolean(c.high_degree_pruning != 0);
  w.f32(c.hub_percentile);
  w.boolean(c.is_compact != 0);
  w.boolean(c.is_recompute != 0);
  // CsrGraph (leann.rs:193-208)
  w.vec_u64(idx->h_offsets.data(), idx->h_offsets.size());
  w.vec_u64(idx->h_nbrs.data(), idx->h_nbrs.size());
  w.vec_u64(idx->h_levels.data(), idx->h_levels.size());
  w.opt_u64(idx->entry >= 0, (uint64_t)idx->entry);
  w.u64(idx->max_level);
  w.u64(idx->n);
  std::vector<uint64_t> deg(idx->n);
  for (uint64_t i = 0; i < idx->n; ++i) deg[i] = idx->h_offsets[i + 1] - idx->h_offsets[i];
  w.vec_u64(deg.data(), deg.size());
  w.opt_u64(idx->n > 0, idx->dim);  // dimension is set by build() (leann.rs:569)
  return emit(w, out, cap, out_len);
} ISL_ABI_GUARD

isl_status isl_index_from_bytes(const uint8_t* bytes, uint64_t len, const float* vectors, uint32_t dim,
                                isl_index** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (!bytes) return fail(ISL_INVALID_ARGUMENT, "bytes is null");
  ByteReader r(bytes, len);
  isl_leann_config c{};
  c.m = r.u64();
  c.m0 = r.u64();
  c.ef_construction = r.u64();
  c.ml = r.f64();
  c.max_layers = r.u64();
  c.metric = (int32_t)r.u32();
  c.ef_search = r.u64();
  c.beam_width = r.u64();
  c.prune_ratio = r.f32();
  c.pruning_strategy = (int32_t)r.u32();
  c.high_degree_pruning = r.boolean();
  c.hub_percentile = r.f32();
  c.is_compact = r.boolean();
  c.is_recompute = r.boolean();
  std::vector<uint64_t> offsets, nbrs, levels, degs;
  r.vec_u64(&offsets);
  r.vec_u64(&nbrs);
  r.vec_u64(&levels);
  uint64_t entry = 0;
  const bool has_entry = r.opt_u64(&entry);
  const uint64_t max_level = r.u64();
  const uint64_t n = r.u64();
  r.vec_u64(&degs);
  uint64_t sdim = 0;
  const bool has_dim = r.opt_u64(&sdim);
  if (!r.done()) return fail(ISL_SERIALIZATION, "deserialization failed: truncated or trailing bytes");
  if (c.metric < 0 || c.metric > 3 || c.pruning_strategy < 0 || c.pruning_strategy > 2)
    return fail(ISL_SERIALIZATION, "deserialization failed: invalid enum variant");
  if (offsets.size() != n + 1 || levels.size() != n || degs.size() != n || offsets[0] != 0 || offsets[n] != nbrs.size())
    return fail(ISL_SERIALIZATION, "deserialization failed: inconsistent CSR arrays");
  if (n > 0) {
    if (!vectors) return fail(ISL_INVALID_ARGUMENT, "vectors is null: the bytes hold the graph only (leann.rs:1058)");
    if (has_dim && sdim != dim)
      return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(sdim) + ", got " + std::to_string(dim));
  }
  isl_status st = isl_index_from_csr(&c, dim, n, offsets.data(), nbrs.data(), levels.data(),
                                     has_entry ? (int64_t)entry : ISL_NO_ENTRY, vectors, out);
  if (st != ISL_OK) return st;
  (*out)->max_level = max_level;
  return ISL_OK;
} ISL_ABI_GUARD

// ProductQuantizer { config, codebooks: Vec<PQCodebook{centroids: Vec<Vec<f32>>, subvector_dim}>, dimension,
// subvector_dim, metric, trained } (pq.rs:116-129)
isl_status isl_pq_to_bytes(const isl_pq* pq, uint8_t* out, uint64_t cap, uint64_t* out_len) try {
  if (!pq) return fail(ISL_INVALID_ARGUMENT, "quantizer is null");
  ByteWriter w;
  w.u64(pq->cfg.num_subquantizers);  // PQConfig (pq.rs:13-22)
  w.u64(pq->cfg.num_centroids);
  w.u64(pq->cfg.training_iterations);
  w.opt_u64(pq->cfg.has_seed != 0, pq->cfg.seed);
  const uint64_t m = pq->trained ? pq->cfg.num_subquantizers : 0;  // codebooks is empty until trained (pq.rs:142)
  w.u64(m);
  for (uint64_t j = 0; j < m; ++j) {
    w.u64(pq->ksub);
    for (uint64_t c = 0; c < pq->ksub; ++c)
      w.vec_f32(pq->h_codebooks.data() + (j * pq->ksub + c) * pq->dsub, pq->dsub);
    w.u64(pq->dsub);
  }
  w.u64(pq->dim);
  w.u64(pq->dsub);
  w.u32((uint32_t)pq->metric);
  w.boolean(pq->trained);
  return emit(w, out, cap, out_len);
} ISL_ABI_GUARD

isl_status isl_pq_from_bytes(const uint8_t* bytes, uint64_t len, isl_pq** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (!bytes) return fail(ISL_INVALID_ARGUMENT, "bytes is null");
  ByteReader r(bytes, len);
  isl_pq_config c{};
  c.num_subquantizers = r.u64();
  c.num_centroids = r.u64();
  c.training_iterations = r.u64();
  uint64_t seed = 0;
  c.has_seed = r.opt_u64(&seed) ? 1 : 0;
  c.seed = c.has_seed ? seed : 0;
  const uint64_t m = r.seq_len(16);
  std::vector<float> cb;
  uint64_t ksub = 0, dsub_cb = 0;
  std::vector<float> row;
  for (uint64_t j = 0; j < m && r.ok; ++j) {
    const uint64_t k = r.seq_len(8);
    if (j == 0) ksub = k;
    if (k != ksub) r.ok = false;
    for (uint64_t cidx = 0; cidx < k && r.ok; ++cidx) {
      r.vec_f32(&row);
      if (j == 0 && cidx == 0) dsub_cb = row.size();
      if (row.size() != dsub_cb) r.ok = false;
      cb.insert(cb.end(), row.begin(), row.end());
    }
    if (r.u64() != dsub_cb) r.ok = false;  // PQCodebook::subvector_dim
  }
  const uint64_t dim = r.u64();
  const uint64_t dsub = r.u64();
  const uint32_t metric = r.u32();
  const bool trained = r.boolean();
  if (!r.done() || metric > 3) return fail(ISL_SERIALIZATION, "deserialization failed: malformed ProductQuantizer bytes");
  if (dim > 0xffffffffull) return fail(ISL_SERIALIZATION, "deserialization failed: dimension out of range");
  isl_pq* pq = nullptr;
  ISL_TRY(isl_pq_new((uint32_t)dim, &c, &pq));
  std::unique_ptr<isl_pq, void (*)(isl_pq*)> guard(pq, isl_pq_free);
  ISL_TRY(isl_pq_set_metric(pq, (int32_t)metric));
  if (trained) {
    if (m != c.num_subquantizers || dsub != dsub_cb || dsub * m != dim || ksub == 0)
      return fail(ISL_SERIALIZATION, "deserialization failed: codebook shape does not match the configuration");
    ISL_TRY(isl_pq_set_codebooks(pq, cb.data(), ksub));
  }
  *out = guard.release();
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_index_get_config(const isl_index* idx, isl_leann_config* out) try {
  if (!idx || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  *out = idx->cfg;
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_pq_get_config(const isl_pq* pq, isl_pq_config* out) try {
  if (!pq || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  *out = pq->cfg;
  return ISL_OK;
} ISL_ABI_GUARD
uint32_t isl_pq_dimension(const isl_pq* pq) { return pq ? pq->dim : 0; }

// InMemoryEmbeddingProvider::compute_embedding (leann.rs:141-150) for the resident embeddings.
isl_status isl_index_get_vector(const isl_index* idx, uint64_t node_id, float* out) try {
  if (!idx || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (node_id >= idx->n) return fail(ISL_NODE_NOT_FOUND, "node " + std::to_string(node_id) + " not found");
  if (!idx->vectors.p) return fail(ISL_INVALID_ARGUMENT, "the stored vectors were dropped (recompute-only index)");
  DeviceGuard g(idx->device);
  ISL_CUDA_TRY(cudaMemcpy(out, idx->vectors.p + node_id * idx->ld, (size_t)idx->dim * 4, cudaMemcpyDeviceToHost));
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
