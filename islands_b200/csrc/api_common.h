// api_common.h — handle definitions shared by the C-ABI translation units.
#pragma once

#include <memory>
#include <mutex>
#include <shared_mutex>
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

// Everything one search call needs besides the resident index: its own stream, events and scratch buffers.
// A handle keeps a pool of these (ScratchLease), so searches on one handle from several threads — legal in the
// reference, where search takes `&self` (leann.rs:858-896) — run side by side instead of queueing on a mutex.
struct SearchScratch {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_in = nullptr;
  DevBuf<uint32_t> visited;
  DevBuf<uint2> r_global;
  DevBuf<uint2> ties_global;
  DevBuf<unsigned int> counters;  // [0] work counter, [1] invariant guard, [2] hub-cache hits
  DevBuf<float> q_stage;
  DevBuf<uint64_t> out_ids;
  DevBuf<float> out_dist;
  DevBuf<uint32_t> out_count;
  DevBuf<isl_search_stats> out_stats;
  DevBuf<float> aux_f32;  // two-level: LUTs / approximate queues
  DevBuf<uint2> aux_u2;
  // ADC traversal -> rerank / recompute hand-over
  DevBuf<uint32_t> rc_flags, rc_rows, rc_surv, rc_surv_cnt;
  DevBuf<int32_t> rc_tok, rc_len;
  DevBuf<float> rc_emb, rc_sq;
  DevBuf<uint8_t> rc_tmp;
  // shard exchange (api_shard.cu)
  DevBuf<uint4> packed, gathered;
  float kernel_ms = 0.0f;
  isl_status init();
  ~SearchScratch();
};

// The stream whose already-enqueued work a `_dev` entry point must wait for (isl_set_caller_stream; legacy
// default stream unless set).  Thread-local.
cudaStream_t caller_stream();
// Makes `st` wait for everything enqueued so far on the caller's stream (device pointers handed to a `_dev`
// entry point may still be being written there).
isl_status order_after_caller(cudaStream_t st, cudaEvent_t ev);

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
  isl::DevBuf<uint32_t> deg_counts;  // [n] graph.degree_counts (PruningStrategy::Proportional)
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
  isl_build_stats build_stats{};  // counters of the construction that produced this graph (zero for adopted graphs)
  uint32_t rerank_limit = 0;  // ADC traversal + rerank / recompute: survivors that get an exact distance (0 = all ef)
  uint32_t tok_len = 0;
  mutable uint64_t last_recomputed = 0;
  mutable float last_encoder_ms = 0.0f, last_traverse_ms = 0.0f, last_rerank_ms = 0.0f;
  // Searches hold `mu` shared (they only read the handle), mutators (attach / drop / set_*) hold it exclusively.
  mutable std::shared_mutex mu;
  // set-up stream and scratch (construction, uploads); searches lease a SearchScratch from the pool instead
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mutable isl::DevBuf<uint32_t> visited;      // construction only (released when the graph is finished)
  mutable isl::DevBuf<uint2> r_global;        // construction only
  mutable isl::DevBuf<uint2> ties_global;     // construction only
  mutable isl::DevBuf<unsigned int> counters; // construction only
  mutable std::mutex pool_mu;                 // guards `pool` and the last_* timing fields
  mutable std::vector<std::unique_ptr<isl::SearchScratch>> pool;
  mutable float last_kernel_ms = 0.0f;
  mutable uint64_t last_launches = 0;
  ~isl_index();
};

namespace isl {
// Takes a SearchScratch out of the handle's pool (creating one when all are in use) and puts it back on scope exit.
struct ScratchLease {
  const isl_index* idx;
  std::unique_ptr<SearchScratch> sc;
  isl_status status = ISL_OK;
  explicit ScratchLease(const isl_index* i);
  ~ScratchLease();
  SearchScratch* operator->() const { return sc.get(); }
  SearchScratch* get() const { return sc.get(); }
};
// Shared by api_index.cu and build.cu.
isl_status index_alloc_common(isl_index* idx);
isl_status index_finish_graph(isl_index* idx);  // uploads CSR, computes max degree
isl_status index_make_padded_adjacency(isl_index* idx);  // device CSR -> adj_pad (or adj_stride = 0)
void search_args_set_graph(const isl_index* idx, SearchArgs* a);
void fill_empty(uint64_t nq, uint32_t k, uint64_t* ids, float* dist, uint32_t* count, isl_search_stats* stats);
isl_status search_checks(const isl_index* idx, const void* queries, uint64_t nq, uint32_t query_dim, uint32_t k,
                         uint32_t* ef, bool* trivial, bool need_vectors = true);
isl_status search_device(const isl_index* idx, SearchScratch* sc, const float* d_queries, uint32_t q_ld, uint64_t nq,
                         uint32_t k, uint32_t ef, uint64_t* d_ids, float* d_dist, uint32_t* d_count,
                         isl_search_stats* d_stats, const ShardOut* shard);
isl_status fill_empty_dev(uint64_t nq, uint32_t k, uint64_t* d_ids, float* d_dist, uint32_t* d_count,
                          isl_search_stats* d_stats);
isl_status stage_device_queries(const isl_index* idx, SearchScratch* sc, const float* d_queries, uint64_t nq,
                                uint32_t query_dim, const float** q, uint32_t* q_ld);
isl_status search_finish(const isl_index* idx, SearchScratch* sc, uint64_t launches);
isl_status pq_upload_codebooks(isl_pq* pq);
// PQ searches (1 two-level, 2 ADC traversal + exact rerank) and the ADC recompute search on a caller-held scratch
// (device guard + shared lock + search_checks done by the caller); shard != null: also write the shard-exchange
// records and leave host copies / final synchronisation to the caller (api_shard.cu).
isl_status pq_search_on_scratch(int mode, const isl_index* idx, SearchScratch* sc, const float* queries, uint64_t nq, uint32_t k,
                                uint32_t ef, float rerank_ratio, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                isl_search_stats* stats, const ShardOut* shard);
isl_status adc_recompute_on_scratch(const isl_index* idx, SearchScratch* sc, const float* queries, uint64_t nq, uint32_t k,
                                    uint32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                    isl_search_stats* stats, const ShardOut* shard);
template <class T>
isl_status ensure(DevBuf<T>& b, size_t count) {
  if (b.n >= count) return ISL_OK;
  cudaError_t e = b.alloc(count);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
  return ISL_OK;
}
}  // namespace isl
