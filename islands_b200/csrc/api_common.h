// api_common.h — handle definitions shared by the C-ABI translation units.
#pragma once

#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "search.h"

namespace isl {

// Switches to `device` for the lifetime of the guard (handles are bound to their device).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

isl_status current_device(int* device, int* sms);
inline uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

}  // namespace isl

struct isl_pq {
  isl_pq_config cfg{};
  uint32_t dim = 0, dsub = 0, ld_sub = 0;
  uint32_t ksub = 0;  // centroids actually held per subspace (min(num_centroids, n_train))
  int32_t metric = ISL_METRIC_EUCLIDEAN;  // pq.rs:146
  bool trained = false;
  int device = 0, sms = 0;
  cudaStream_t stream = nullptr;
  std::vector<float> h_codebooks;      // [m][ksub][dsub]
  isl::DevBuf<float> d_codebooks;      // [m][ksub][ld_sub]
  mutable std::mutex mu;
  isl::PqDev dev() const {
    isl::PqDev p;
    p.codebooks = d_codebooks.p;
    p.m = (uint32_t)cfg.num_subquantizers;
    p.ksub = ksub;
    p.dsub = dsub;
    p.ld_sub = ld_sub;
    p.metric = metric;
    return p;
  }
};

struct isl_index {
  isl_leann_config cfg{};
  uint32_t dim = 0, ld = 0;
  uint64_t n = 0;
  int device = 0, sms = 0;
  // CsrGraph (leann.rs:193-208), reference layout on the host for export
  std::vector<uint64_t> h_offsets, h_nbrs, h_levels;
  int64_t entry = ISL_NO_ENTRY;
  uint64_t max_level = 0;
  uint32_t max_degree = 0;
  // resident in HBM
  isl::DevBuf<float> vectors;    // [n][ld]
  isl::DevBuf<float> sqnorms;    // [n]
  isl::DevBuf<uint64_t> offsets; // [n+1]
  isl::DevBuf<uint32_t> nbrs;    // [E]
  // Search-time adjacency: fixed-stride rows padded with 0xffffffff (stride = max degree rounded
  // up to 32).  Used instead of the CSR arrays when it costs at most ~2x their memory.
  isl::DevBuf<uint32_t> adj_pad; // [n][adj_stride]
  uint32_t adj_stride = 0;       // 0 => search walks the CSR arrays
  bool lists_unique = false;     // no neighbour list names an id twice (scanned once per graph)
  // two-level search attachment
  const isl_pq* pq = nullptr;
  isl::DevBuf<uint8_t> codes8;    // [n][m] when ksub <= 256
  isl::DevBuf<uint16_t> codes16;  // [n][m] otherwise
  // on-demand recompute attachment (EmbeddingProvider seam, leann.rs:82-99): token rows per node
  isl_encoder* encoder = nullptr;
  isl::DevBuf<int32_t> node_tokens;   // [n][tok_len]
  isl::DevBuf<int32_t> node_lengths;  // [n]
  // hub-embedding cache (docs/leann-specification.md:661-690): resident rows of the highest in-degree nodes
  uint64_t hub_count = 0;
  isl::DevBuf<float> hub_emb;         // [hub_count][ld]
  isl::DevBuf<float> hub_sq;          // [hub_count]
  isl::DevBuf<uint32_t> hub_row;      // [n] row in hub_emb or 0xffffffff
  mutable uint64_t last_hub_hits = 0;
  uint32_t rerank_limit = 0;  // ADC traversal + rerank / recompute: survivors that get an exact distance (0 = all ef)
  uint32_t tok_len = 0;
  mutable isl::DevBuf<uint32_t> rc_flags, rc_rows, rc_surv, rc_surv_cnt;
  mutable isl::DevBuf<int32_t> rc_tok, rc_len;
  mutable isl::DevBuf<float> rc_emb, rc_sq;
  mutable isl::DevBuf<uint8_t> rc_tmp;
  mutable uint64_t last_recomputed = 0;
  mutable float last_encoder_ms = 0.0f, last_traverse_ms = 0.0f, last_rerank_ms = 0.0f;
  // per-handle stream, timing and scratch (guarded by mu: searches on one handle serialise)
  mutable std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mutable float last_kernel_ms = 0.0f;
  mutable uint64_t last_launches = 0;
  mutable isl::DevBuf<uint32_t> visited;
  mutable isl::DevBuf<uint2> r_global;
  mutable isl::DevBuf<unsigned int> counters;  // [0] work counter, [1] error flag
  mutable isl::DevBuf<float> q_stage;
  mutable isl::DevBuf<uint64_t> out_ids;
  mutable isl::DevBuf<float> out_dist;
  mutable isl::DevBuf<uint32_t> out_count;
  mutable isl::DevBuf<isl_search_stats> out_stats;
  mutable isl::DevBuf<float> aux_f32;   // two-level: LUTs / approximate queues
  mutable isl::DevBuf<uint2> aux_u2;
  ~isl_index();
};

namespace isl {
// Shared by api_index.cu and build.cu.
isl_status index_alloc_common(isl_index* idx);
isl_status index_finish_graph(isl_index* idx);  // uploads CSR, computes max degree
isl_status index_make_padded_adjacency(isl_index* idx);  // device CSR -> adj_pad (or adj_stride = 0)
void search_args_set_graph(const isl_index* idx, SearchArgs* a);
void fill_empty(uint64_t nq, uint32_t k, uint64_t* ids, float* dist, uint32_t* count, isl_search_stats* stats);
isl_status search_checks(const isl_index* idx, const void* queries, uint64_t nq, uint32_t query_dim, uint32_t k,
                         uint32_t* ef, bool* trivial);
isl_status search_finish(const isl_index* idx);
isl_status pq_upload_codebooks(isl_pq* pq);
template <class T>
isl_status ensure(DevBuf<T>& b, size_t count) {
  if (b.n >= count) return ISL_OK;
  cudaError_t e = b.alloc(count);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
  return ISL_OK;
}
}  // namespace isl
