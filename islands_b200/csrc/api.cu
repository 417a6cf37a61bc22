// api.cu — C ABI: library, configs, distance entry points, LEANN index (from_csr / search /
// export), shard merge.  See include/islands_b200.h for the reference lines each one replaces.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <stdexcept>

#include "api_common.h"

namespace isl {

static thread_local std::string t_last_error;
static thread_local uint64_t t_err_a = 0, t_err_b = 0;  // payload of the last error (isl_last_error_detail)
std::atomic<uint64_t> g_launch_count{0};

void set_last_error(const std::string& msg) { t_last_error = msg; }

// The numbers a CoreError variant carries (error.rs:9-62) travel beside the message: DimensionMismatch{expected,
// actual} and NodeNotFound(id).  Every such message of this library is written "expected X, got Y" / names the id
// first, so the payload is taken from the text in one place instead of at fifty call sites.
static void capture_payload(isl_status st, const std::string& msg) {
  t_err_a = t_err_b = 0;
  if (st != ISL_DIM_MISMATCH && st != ISL_NODE_NOT_FOUND) return;
  uint64_t v[2] = {0, 0};
  int found = 0;
  for (size_t i = 0; i < msg.size() && found < 2;) {
    if (msg[i] >= '0' && msg[i] <= '9') {
      uint64_t x = 0;
      while (i < msg.size() && msg[i] >= '0' && msg[i] <= '9') x = x * 10 + (uint64_t)(msg[i++] - '0');
      v[found++] = x;
    } else {
      ++i;
    }
  }
  t_err_a = v[0];
  t_err_b = st == ISL_DIM_MISMATCH ? v[1] : 0;
}
isl_status fail(isl_status st, const std::string& msg) {
  t_last_error = msg;
  capture_payload(st, msg);
  return st;
}
// Called from a catch (...) handler at the ABI (ISL_ABI_GUARD): turns the exception in flight into a status.
isl_status exception_status() noexcept {
  try {
    try {
      throw;
    } catch (const std::bad_alloc&) {
      return fail(ISL_INVALID_ARGUMENT, "host memory allocation failed (std::bad_alloc)");
    } catch (const std::exception& e) {
      return fail(ISL_INVALID_ARGUMENT, std::string("host exception: ") + e.what());
    } catch (...) {
      return fail(ISL_INVALID_ARGUMENT, "host exception of unknown type");
    }
  } catch (...) {  // not even the message could be stored
    return ISL_INVALID_ARGUMENT;
  }
}
isl_status cuda_fail(cudaError_t e, const char* what) {
  t_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  cudaGetLastError();  // clear the sticky-free error state
  return ISL_CUDA_ERROR;
}

isl_status current_device(int* device, int* sms) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(ISL_CUDA_ERROR,
                "no usable CUDA device: islands_b200 has no CPU fallback (cudaGetDeviceCount: " +
                    std::string(e == cudaSuccess ? "0 devices" : cudaGetErrorString(e)) + ")");
  }
  ISL_CUDA_TRY(cudaGetDevice(device));
  ISL_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, *device));
  return ISL_OK;
}

// leann.rs:432-460
static isl_status validate_leann(const isl_leann_config* c) {
  if (!c) return fail(ISL_INVALID_ARGUMENT, "config is null");
  if (c->m == 0) return fail(ISL_INVALID_CONFIG, "M must be > 0");
  if (c->m0 < c->m) return fail(ISL_INVALID_CONFIG, "M0 must be >= M");
  if (c->ef_construction < c->m) return fail(ISL_INVALID_CONFIG, "ef_construction must be >= M");
  if (!(c->prune_ratio >= 0.0f && c->prune_ratio <= 1.0f))
    return fail(ISL_INVALID_CONFIG, "prune_ratio must be in [0.0, 1.0]");
  if (c->beam_width == 0) return fail(ISL_INVALID_CONFIG, "beam_width must be > 0");
  if (!(c->hub_percentile >= 0.0f && c->hub_percentile <= 1.0f))
    return fail(ISL_INVALID_CONFIG, "hub_percentile must be in [0.0, 1.0]");
  if (c->metric < 0 || c->metric > 3) return fail(ISL_INVALID_CONFIG, "unknown metric");
  if (c->pruning_strategy < 0 || c->pruning_strategy > 2)
    return fail(ISL_INVALID_CONFIG, "unknown pruning strategy");
  return ISL_OK;
}

isl_status index_alloc_common(isl_index* idx) {
  ISL_TRY(current_device(&idx->device, &idx->sms));
  ISL_CUDA_TRY(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
  ISL_CUDA_TRY(cudaEventCreate(&idx->ev0));
  ISL_CUDA_TRY(cudaEventCreate(&idx->ev1));
  ISL_TRY(ensure(idx->counters, 4));
  return ISL_OK;
}

isl_status SearchScratch::init() {
  ISL_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  ISL_CUDA_TRY(cudaEventCreate(&ev0));
  ISL_CUDA_TRY(cudaEventCreate(&ev1));
  ISL_CUDA_TRY(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
  ISL_TRY(ensure(counters, 4));
  return ISL_OK;
}
SearchScratch::~SearchScratch() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (ev_in) cudaEventDestroy(ev_in);
  if (stream) cudaStreamDestroy(stream);
}

ScratchLease::ScratchLease(const isl_index* i) : idx(i) {
  {
    std::lock_guard<std::mutex> lock(idx->pool_mu);
    if (!idx->pool.empty()) {
      sc = std::move(idx->pool.back());
      idx->pool.pop_back();
    }
  }
  if (!sc) {
    sc.reset(new SearchScratch());
    status = sc->init();
  }
}
ScratchLease::~ScratchLease() {
  if (!sc || status != ISL_OK) return;  // a scratch that failed to initialise is dropped
  std::lock_guard<std::mutex> lock(idx->pool_mu);
  idx->pool.push_back(std::move(sc));
}

static thread_local cudaStream_t t_caller_stream = nullptr;  // nullptr = the legacy default stream
cudaStream_t caller_stream() { return t_caller_stream ? t_caller_stream : cudaStreamLegacy; }
isl_status order_after_caller(cudaStream_t st, cudaEvent_t ev) {
  ISL_CUDA_TRY(cudaEventRecord(ev, caller_stream()));
  ISL_CUDA_TRY(cudaStreamWaitEvent(st, ev, 0));
  return ISL_OK;
}

// Uploads h_offsets / h_nbrs to the device (u64 offsets, u32 ids) and derives max_degree.
isl_status index_finish_graph(isl_index* idx) {
  const uint64_t n = idx->n;
  uint32_t maxdeg = 0;
  for (uint64_t i = 0; i < n; ++i)
    maxdeg = std::max<uint32_t>(maxdeg, (uint32_t)(idx->h_offsets[i + 1] - idx->h_offsets[i]));
  idx->max_degree = maxdeg;
  const uint64_t e = idx->h_nbrs.size();
  ISL_CUDA_TRY(idx->offsets.alloc(n + 1));
  ISL_CUDA_TRY(cudaMemcpyAsync(idx->offsets.p, idx->h_offsets.data(), (n + 1) * 8,
                               cudaMemcpyHostToDevice, idx->stream));
  std::vector<uint32_t> n32(e);
  for (uint64_t i = 0; i < e; ++i) n32[i] = (uint32_t)idx->h_nbrs[i];
  ISL_CUDA_TRY(idx->nbrs.alloc(std::max<uint64_t>(e, 1)));
  if (e)
    ISL_CUDA_TRY(cudaMemcpyAsync(idx->nbrs.p, n32.data(), e * 4, cudaMemcpyHostToDevice, idx->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(idx->stream));
  return index_make_padded_adjacency(idx);
}

isl_status index_make_padded_adjacency(isl_index* idx) {
  idx->adj_pad.release();
  idx->adj_stride = 0;
  idx->lists_unique = false;
  const uint64_t n = idx->n, e = idx->h_nbrs.size();
  if (n == 0) return ISL_OK;
  ISL_CUDA_TRY(idx->deg_counts.alloc(n));  // graph.degree_counts on the device (PruningStrategy::Proportional)
  ISL_TRY(launch_degree_counts(idx->offsets.p, n, idx->deg_counts.p, idx->stream));
  {
    DevBuf<unsigned int> flag;
    ISL_CUDA_TRY(flag.alloc(1));
    ISL_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(unsigned int), idx->stream));
    ISL_TRY(launch_list_duplicates(idx->offsets.p, idx->nbrs.p, n, flag.p, idx->stream));
    unsigned int h = 1;
    ISL_CUDA_TRY(cudaMemcpyAsync(&h, flag.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, idx->stream));
    ISL_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    idx->lists_unique = (h == 0);
  }
  const uint32_t stride = std::max<uint32_t>(32, round_up(idx->max_degree, 32));
  if ((uint64_t)n * stride > 2 * e + 64 * n) return ISL_OK;  // very skewed degrees: keep CSR
  ISL_CUDA_TRY(idx->adj_pad.alloc(n * stride));
  ISL_TRY(launch_pad_adjacency(idx->offsets.p, idx->nbrs.p, n, stride, idx->adj_pad.p, idx->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(idx->stream));
  idx->adj_stride = stride;
  return ISL_OK;
}

void search_args_set_graph(const isl_index* idx, SearchArgs* a) {
  a->degrees = nullptr;
  a->lists_unique = idx->lists_unique ? 1u : 0u;
  if (idx->adj_stride) {
    a->offsets = nullptr;
    a->nbrs = idx->adj_pad.p;
    a->adj_stride = idx->adj_stride;
  } else {
    a->offsets = idx->offsets.p;
    a->nbrs = idx->nbrs.p;
    a->adj_stride = 0;
  }
}

void fill_empty(uint64_t nq, uint32_t k, uint64_t* ids, float* dist, uint32_t* count,
                       isl_search_stats* stats) {
  for (uint64_t i = 0; i < nq * (uint64_t)k; ++i) {
    if (ids) ids[i] = ISL_INVALID_ID;
    if (dist) dist[i] = std::numeric_limits<float>::infinity();
  }
  for (uint64_t q = 0; q < nq; ++q) {
    if (count) count[q] = 0;
    if (stats) stats[q] = isl_search_stats{0, 0, 0, 0, 0};
  }
}

// Core of isl_index_search*: queries already on the device as [nq][q_ld] (q_ld % 4 == 0, zero
// padded), outputs on the device.  shard: where the shard-exchange records go (api_shard.cu), or null.
isl_status search_device(const isl_index* idx, SearchScratch* sc, const float* d_queries, uint32_t q_ld,
                         uint64_t nq, uint32_t k, uint32_t ef, uint64_t* d_ids, float* d_dist,
                         uint32_t* d_count, isl_search_stats* d_stats, const ShardOut* shard) {
  SearchPlan plan;
  const uint32_t u_cap = std::max<uint32_t>(32, round_up(idx->max_degree, 32));
  ISL_TRY(plan_search(idx->cfg.metric, idx->ld, ef, u_cap, idx->sms, &plan));
  const uint32_t vis_words = round_up((uint32_t)((idx->n + 31) / 32), 4);
  const uint32_t slots = (uint32_t)std::min<uint64_t>(plan.grid, nq);
  ISL_TRY(ensure(sc->visited, (size_t)slots * vis_words));
  if (!plan.r_in_smem) ISL_TRY(ensure(sc->r_global, (size_t)slots * ef));
  ISL_TRY(ensure(sc->ties_global, (size_t)slots * ef));
  ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, 4 * sizeof(unsigned int), sc->stream));

  SearchArgs a{};
  a.vectors = idx->vectors.p;
  a.sqnorms = idx->sqnorms.p;
  a.ld = idx->ld;
  a.d = idx->dim;
  a.n = (uint32_t)idx->n;
  search_args_set_graph(idx, &a);
  a.queries = d_queries;
  a.q_ld = q_ld;
  a.nq = (uint32_t)nq;
  a.entry = (uint32_t)idx->entry;
  a.k = k;
  a.ef = ef;
  a.metric = idx->cfg.metric;
  a.prune_ratio = idx->cfg.prune_ratio;
  a.strategy = idx->cfg.pruning_strategy;
  a.prune_seed = idx->cfg.prune_seed;
  a.deg_counts = idx->deg_counts.p;
  a.visited = sc->visited.p;
  a.vis_words = vis_words;
  a.r_global = sc->r_global.p;
  a.ties_global = sc->ties_global.p;
  a.u_cap = u_cap;
  a.out_ids = d_ids;
  a.out_ids32 = nullptr;
  a.out_dist = d_dist;
  a.out_count = d_count;
  if (shard) a.shard = *shard;
  a.stats = d_stats;
  a.work_counter = sc->counters.p;
  a.error_flag = sc->counters.p + 1;

  ISL_CUDA_TRY(cudaEventRecord(sc->ev0, sc->stream));
  ISL_TRY(launch_search(plan, a, sc->stream));
  ISL_CUDA_TRY(cudaEventRecord(sc->ev1, sc->stream));
  return ISL_OK;
}

// Waits for the call's stream, reads the guard flag and publishes the kernel time on the handle.
isl_status search_finish(const isl_index* idx, SearchScratch* sc, uint64_t launches) {
  unsigned int h[4] = {0, 0, 0, 0};
  ISL_CUDA_TRY(cudaMemcpyAsync(h, sc->counters.p, sizeof(h), cudaMemcpyDeviceToHost, sc->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(sc->stream));
  float ms = 0.0f;
  if (cudaEventElapsedTime(&ms, sc->ev0, sc->ev1) == cudaSuccess) sc->kernel_ms = ms;
  {
    std::lock_guard<std::mutex> lock(idx->pool_mu);
    idx->last_kernel_ms = sc->kernel_ms;
    idx->last_launches = launches;
  }
  if (h[1]) return fail(ISL_CUDA_ERROR, "search: internal invariant violated (tie list overflow)");
  return ISL_OK;
}

isl_status search_checks(const isl_index* idx, const void* queries, uint64_t nq,
                                uint32_t query_dim, uint32_t k, uint32_t* ef, bool* trivial, bool need_vectors) {
  *trivial = false;
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (nq > 0 && !queries) return fail(ISL_INVALID_ARGUMENT, "queries is null");
  if (nq > 0xffffffffull) return fail(ISL_INVALID_ARGUMENT, "too many queries in one batch");
  if (idx->n == 0 || nq == 0 || k == 0) {  // leann.rs:875-877 / take(0)
    *trivial = true;
    return ISL_OK;
  }
  if (query_dim != idx->dim)  // leann.rs:880-887
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(idx->dim) +
                                      ", got " + std::to_string(query_dim));
  if (idx->entry < 0) return fail(ISL_INDEX_NOT_BUILT, "index not built");  // leann.rs:889
  if (need_vectors && !idx->vectors.p)
    return fail(ISL_INVALID_ARGUMENT, "the stored vectors were dropped (isl_index_drop_vectors): only "
                                      "isl_index_search_adc_recompute works on this handle");
  *ef = std::max(*ef, k);  // leann.rs:890
  if (*ef > (1u << 24)) return fail(ISL_INVALID_ARGUMENT, "ef too large");
  return ISL_OK;
}

// Device outputs of a search over an empty index / zero k: ids all ones, distances +inf, counts 0.
isl_status fill_empty_dev(uint64_t nq, uint32_t k, uint64_t* d_ids, float* d_dist, uint32_t* d_count,
                          isl_search_stats* d_stats) {
  if (nq && k) {
    if (d_ids) ISL_CUDA_TRY(cudaMemset(d_ids, 0xff, nq * k * 8));
    if (d_dist) {
      std::vector<float> inf(nq * k, std::numeric_limits<float>::infinity());
      ISL_CUDA_TRY(cudaMemcpy(d_dist, inf.data(), nq * k * 4, cudaMemcpyHostToDevice));
    }
  }
  if (d_count && nq) ISL_CUDA_TRY(cudaMemset(d_count, 0, nq * 4));
  if (d_stats && nq) ISL_CUDA_TRY(cudaMemset(d_stats, 0, nq * sizeof(isl_search_stats)));
  return ISL_OK;
}

// Queries on the device: used as they are when aligned, else padded into the call's staging buffer.
isl_status stage_device_queries(const isl_index* idx, SearchScratch* sc, const float* d_queries, uint64_t nq,
                                uint32_t query_dim, const float** q, uint32_t* q_ld) {
  *q = d_queries;
  *q_ld = query_dim;
  if (query_dim % 4 != 0 || (reinterpret_cast<uintptr_t>(d_queries) & 15)) {
    ISL_TRY(ensure(sc->q_stage, nq * idx->ld));
    ISL_TRY(launch_pad_rows(d_queries, query_dim, sc->q_stage.p, idx->ld, nq, sc->stream));
    *q = sc->q_stage.p;
    *q_ld = idx->ld;
  }
  return ISL_OK;
}

}  // namespace isl

using namespace isl;

isl_index::~isl_index() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (stream) cudaStreamDestroy(stream);
}

extern "C" {

int isl_abi_version(void) { return ISL_ABI_VERSION; }
const char* isl_last_error(void) { return t_last_error.c_str(); }
void isl_last_error_detail(uint64_t* a, uint64_t* b) {
  if (a) *a = t_err_a;
  if (b) *b = t_err_b;
}
int isl_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return c;
}
isl_status isl_set_caller_stream(void* cuda_stream) try {
  t_caller_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  return ISL_OK;
} ISL_ABI_GUARD
uint64_t isl_kernel_launch_count(void) { return g_launch_count.load(); }
void isl_kernel_launch_count_reset(void) { g_launch_count.store(0); }

// ---- configs ------------------------------------------------------------------------------
isl_status isl_leann_config_default(isl_leann_config* c) try {
  if (!c) return fail(ISL_INVALID_ARGUMENT, "out is null");
  c->m = 30;  // leann.rs:386-403
  c->m0 = 60;
  c->ef_construction = 128;
  c->ml = 1.0 / std::log(30.0);
  c->max_layers = 16;
  c->metric = ISL_METRIC_COSINE;
  c->ef_search = 64;
  c->beam_width = 1;
  c->prune_ratio = 0.0f;
  c->pruning_strategy = ISL_PRUNE_GLOBAL;
  c->high_degree_pruning = 1;
  c->hub_percentile = 0.02f;
  c->is_compact = 1;
  c->is_recompute = 1;
  c->prune_seed = 0;
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_leann_config_fast(isl_leann_config* c) try {
  ISL_TRY(isl_leann_config_default(c));  // leann.rs:406-416
  c->m = 16;
  c->m0 = 32;
  c->ef_construction = 100;
  c->ef_search = 32;
  c->beam_width = 1;
  c->prune_ratio = 0.3f;
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_leann_config_accurate(isl_leann_config* c) try {
  ISL_TRY(isl_leann_config_default(c));  // leann.rs:419-429
  c->m = 48;
  c->m0 = 96;
  c->ef_construction = 400;
  c->ef_search = 128;
  c->beam_width = 1;
  c->prune_ratio = 0.0f;
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_leann_config_validate(const isl_leann_config* c) { return validate_leann(c); }

isl_status isl_hnsw_config_default(isl_hnsw_config* c) try {
  if (!c) return fail(ISL_INVALID_ARGUMENT, "out is null");
  c->m = 16;  // hnsw.rs:37-48
  c->m0 = 32;
  c->ef_construction = 200;
  c->ml = 1.0 / std::log(16.0);
  c->metric = ISL_METRIC_COSINE;
  c->max_layers = 16;
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_hnsw_config_validate(const isl_hnsw_config* c) try {
  if (!c) return fail(ISL_INVALID_ARGUMENT, "config is null");
  if (c->m == 0) return fail(ISL_INVALID_CONFIG, "M must be > 0");  // hnsw.rs:72-85
  if (c->m0 < c->m) return fail(ISL_INVALID_CONFIG, "M0 must be >= M");
  if (c->ef_construction < c->m) return fail(ISL_INVALID_CONFIG, "ef_construction must be >= M");
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_pq_config_default(isl_pq_config* c) try {
  if (!c) return fail(ISL_INVALID_ARGUMENT, "out is null");
  c->num_subquantizers = 8;  // pq.rs:24-33
  c->num_centroids = 256;
  c->training_iterations = 25;
  c->seed = 0;  // None
  c->has_seed = 0;
  return ISL_OK;
} ISL_ABI_GUARD
isl_status isl_pq_config_validate(const isl_pq_config* c, uint64_t dimension) try {
  if (!c) return fail(ISL_INVALID_ARGUMENT, "config is null");
  if (c->num_subquantizers == 0)  // pq.rs:37-55
    return fail(ISL_INVALID_CONFIG, "num_subquantizers must be > 0");
  if (dimension % c->num_subquantizers != 0)
    return fail(ISL_INVALID_CONFIG, "dimension " + std::to_string(dimension) +
                                        " must be divisible by num_subquantizers " +
                                        std::to_string(c->num_subquantizers));
  if (c->num_centroids == 0 || c->num_centroids > 65536)
    return fail(ISL_INVALID_CONFIG, "num_centroids must be in range [1, 65536]");
  return ISL_OK;
} ISL_ABI_GUARD
uint64_t isl_pq_config_bytes_per_vector(const isl_pq_config* c) {
  if (!c) return 0;
  return c->num_centroids <= 256 ? c->num_subquantizers : c->num_subquantizers * 2;  // pq.rs:58-64
}

// ---- distances ------------------------------------------------------------------------------
static isl_status distance_host(int32_t metric, const float* q, const float* rows, uint64_t n_rows,
                                uint32_t dim, float* out, bool squared) {
  if (metric < 0 || metric > 3) return fail(ISL_INVALID_CONFIG, "unknown metric");
  if (n_rows == 0) return ISL_OK;
  if (!q || !rows || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  const uint32_t ld = std::max<uint32_t>(4, round_up(dim, 4));
  DevBuf<float> dq, dr, dout;
  ISL_CUDA_TRY(dq.alloc(ld));
  ISL_CUDA_TRY(dr.alloc(n_rows * ld));
  ISL_CUDA_TRY(dout.alloc(n_rows));
  ISL_CUDA_TRY(cudaMemset(dq.p, 0, dq.bytes()));
  ISL_CUDA_TRY(cudaMemset(dr.p, 0, dr.bytes()));
  if (dim) {
    ISL_CUDA_TRY(cudaMemcpy(dq.p, q, (size_t)dim * 4, cudaMemcpyHostToDevice));
    ISL_CUDA_TRY(cudaMemcpy2D(dr.p, (size_t)ld * 4, rows, (size_t)dim * 4, (size_t)dim * 4, n_rows,
                              cudaMemcpyHostToDevice));
  }
  ISL_TRY(launch_distance_batch(metric, squared, dq.p, dr.p, n_rows, dim, ld, dout.p, sms, 0));
  ISL_CUDA_TRY(cudaMemcpy(out, dout.p, n_rows * 4, cudaMemcpyDeviceToHost));
  return ISL_OK;
}

isl_status isl_distance_calculate(int32_t metric, const float* a, uint64_t len_a, const float* b,
                                  uint64_t len_b, float* out) try {
  if (len_a != len_b)  // distance.rs:39-44
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(len_a) + ", got " +
                                      std::to_string(len_b));
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  return distance_host(metric, a, b, 1, (uint32_t)len_a, out, false);
} ISL_ABI_GUARD

isl_status isl_distance_calculate_squared(int32_t metric, const float* a, uint64_t len_a,
                                          const float* b, uint64_t len_b, float* out) try {
  if (len_a != len_b)  // distance.rs:55-60
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(len_a) + ", got " +
                                      std::to_string(len_b));
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  return distance_host(metric, a, b, 1, (uint32_t)len_a, out, true);
} ISL_ABI_GUARD

isl_status isl_distance_batch(int32_t metric, const float* query, const float* rows, uint64_t n_rows,
                              uint32_t dim, float* out) try {
  return distance_host(metric, query, rows, n_rows, dim, out, false);
} ISL_ABI_GUARD

isl_status isl_distance_batch_dev(int32_t metric, const float* d_query, const float* d_rows,
                                  uint64_t n_rows, uint32_t dim, float* d_out) try {
  if (metric < 0 || metric > 3) return fail(ISL_INVALID_CONFIG, "unknown metric");
  if (n_rows == 0) return ISL_OK;
  if (!d_query || !d_rows || !d_out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (dim % 4 != 0 || (reinterpret_cast<uintptr_t>(d_rows) & 15) || (reinterpret_cast<uintptr_t>(d_query) & 15))
    return fail(ISL_INVALID_ARGUMENT, "isl_distance_batch_dev needs dim % 4 == 0 and 16-byte aligned pointers");
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  ISL_TRY(launch_distance_batch(metric, false, d_query, d_rows, n_rows, dim, dim, d_out, sms, 0));
  ISL_CUDA_TRY(cudaStreamSynchronize(0));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_normalize_rows(float* rows, uint64_t n_rows, uint32_t dim) try {
  if (n_rows == 0 || dim == 0) return ISL_OK;
  if (!rows) return fail(ISL_INVALID_ARGUMENT, "rows is null");
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  const uint32_t ld = round_up(dim, 4);
  DevBuf<float> dr;
  ISL_CUDA_TRY(dr.alloc(n_rows * ld));
  ISL_CUDA_TRY(cudaMemset(dr.p, 0, dr.bytes()));
  ISL_CUDA_TRY(cudaMemcpy2D(dr.p, (size_t)ld * 4, rows, (size_t)dim * 4, (size_t)dim * 4, n_rows,
                            cudaMemcpyHostToDevice));
  ISL_TRY(launch_normalize_rows(dr.p, n_rows, dim, ld, sms, 0));
  ISL_CUDA_TRY(cudaMemcpy2D(rows, (size_t)dim * 4, dr.p, (size_t)ld * 4, (size_t)dim * 4, n_rows,
                            cudaMemcpyDeviceToHost));
  return ISL_OK;
} ISL_ABI_GUARD

// ---- LEANN index -----------------------------------------------------------------------------
isl_status isl_index_from_csr(const isl_leann_config* cfg, uint32_t dim, uint64_t n,
                              const uint64_t* node_offsets, const uint64_t* neighbors,
                              const uint64_t* levels, int64_t entry_point, const float* vectors,
                              isl_index** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  ISL_TRY(validate_leann(cfg));  // LeannIndex::new (leann.rs:504-511)
  if (n >= (1ull << 31)) return fail(ISL_INVALID_ARGUMENT, "n must be < 2^31 per index (shard larger sets)");
  if (n > 0) {
    if (!node_offsets || !vectors) return fail(ISL_INVALID_ARGUMENT, "null pointer");
    if (dim == 0) return fail(ISL_INVALID_ARGUMENT, "dim must be > 0");
    if (node_offsets[0] != 0) return fail(ISL_SERIALIZATION, "node_offsets[0] must be 0");
    for (uint64_t i = 0; i < n; ++i)
      if (node_offsets[i + 1] < node_offsets[i]) return fail(ISL_SERIALIZATION, "node_offsets must be non-decreasing");
    if (node_offsets[n] > 0 && !neighbors) return fail(ISL_INVALID_ARGUMENT, "neighbors is null");
    for (uint64_t i = 0; i < node_offsets[n]; ++i)
      if (neighbors[i] >= n)  // the provider would fail with NodeNotFound (leann.rs:146-149)
        return fail(ISL_NODE_NOT_FOUND, "neighbor id " + std::to_string(neighbors[i]) + " out of range");
    if (entry_point >= (int64_t)n) return fail(ISL_NODE_NOT_FOUND, "entry point out of range");
  }
  std::unique_ptr<isl_index> idx(new isl_index());
  idx->cfg = *cfg;
  idx->n = n;
  idx->dim = n ? dim : 0;
  idx->ld = n ? std::max<uint32_t>(4, round_up(dim, 4)) : 0;
  idx->entry = n ? entry_point : ISL_NO_ENTRY;
  ISL_TRY(index_alloc_common(idx.get()));
  if (n) {
    idx->h_offsets.assign(node_offsets, node_offsets + n + 1);
    idx->h_nbrs.assign(neighbors, neighbors + node_offsets[n]);
    if (levels)
      idx->h_levels.assign(levels, levels + n);
    else
      idx->h_levels.assign(n, 0);
    idx->max_level = 0;
    if (entry_point >= 0) idx->max_level = idx->h_levels[entry_point];
    ISL_CUDA_TRY(idx->vectors.alloc(n * idx->ld));
    if (idx->ld != dim) ISL_CUDA_TRY(cudaMemsetAsync(idx->vectors.p, 0, idx->vectors.bytes(), idx->stream));
    ISL_CUDA_TRY(cudaMemcpy2DAsync(idx->vectors.p, (size_t)idx->ld * 4, vectors, (size_t)dim * 4,
                                   (size_t)dim * 4, n, cudaMemcpyHostToDevice, idx->stream));
    ISL_CUDA_TRY(idx->sqnorms.alloc(n));
    ISL_TRY(launch_row_sqnorms(idx->vectors.p, n, dim, idx->ld, idx->sqnorms.p, idx->sms, idx->stream));
    ISL_TRY(index_finish_graph(idx.get()));
  } else {
    idx->h_offsets.assign(1, 0);
  }
  *out = idx.release();
  return ISL_OK;
} ISL_ABI_GUARD

void isl_index_free(isl_index* idx) {
  if (!idx) return;
  DeviceGuard g(idx->device);
  delete idx;
}
uint64_t isl_index_len(const isl_index* idx) { return idx ? idx->n : 0; }
uint32_t isl_index_dimension(const isl_index* idx) { return idx ? idx->dim : 0; }
uint64_t isl_index_num_edges(const isl_index* idx) { return idx ? idx->h_nbrs.size() : 0; }
int64_t isl_index_entry_point(const isl_index* idx) { return idx ? idx->entry : ISL_NO_ENTRY; }
uint64_t isl_index_max_level(const isl_index* idx) { return idx ? idx->max_level : 0; }
uint64_t isl_index_storage_bytes(const isl_index* idx) {
  if (!idx) return 0;  // leann.rs:296-301: offsets + neighbors + levels + degree_counts, 8 B each
  return (idx->h_offsets.size() + idx->h_nbrs.size() + idx->h_levels.size() + idx->n) * 8;
}

isl_status isl_index_export_csr(const isl_index* idx, uint64_t* node_offsets, uint64_t* neighbors,
                                uint64_t* levels, uint64_t* degree_counts) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (node_offsets) std::memcpy(node_offsets, idx->h_offsets.data(), idx->h_offsets.size() * 8);
  if (neighbors && !idx->h_nbrs.empty()) std::memcpy(neighbors, idx->h_nbrs.data(), idx->h_nbrs.size() * 8);
  if (levels && idx->n) std::memcpy(levels, idx->h_levels.data(), idx->n * 8);
  if (degree_counts)
    for (uint64_t i = 0; i < idx->n; ++i) degree_counts[i] = idx->h_offsets[i + 1] - idx->h_offsets[i];
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_index_get_neighbors(const isl_index* idx, uint64_t node_id, uint64_t* out, uint64_t cap,
                                   uint64_t* out_count) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (node_id >= idx->n) return fail(ISL_NODE_NOT_FOUND, "node " + std::to_string(node_id) + " not found");
  const uint64_t s = idx->h_offsets[node_id], e = idx->h_offsets[node_id + 1];
  if (out_count) *out_count = e - s;
  for (uint64_t i = 0; i < e - s && i < cap; ++i) out[i] = idx->h_nbrs[s + i];
  return ISL_OK;
} ISL_ABI_GUARD

// CsrGraph::set_neighbors (leann.rs:256-293).  Same-length lists are overwritten in place, any other length
// rebuilds offsets and neighbours (the reference's "expensive but rare" branch); the device copy of the graph
// (CSR + padded adjacency) is refreshed either way.
isl_status isl_index_set_neighbors(isl_index* idx, uint64_t node_id, const uint64_t* neighbors, uint64_t count) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (node_id >= idx->n) return ISL_OK;  // leann.rs:258-260: silently ignored
  if (count && !neighbors) return fail(ISL_INVALID_ARGUMENT, "neighbors is null");
  for (uint64_t i = 0; i < count; ++i)
    if (neighbors[i] >= idx->n)  // the provider would fail with NodeNotFound at search time (leann.rs:146-149)
      return fail(ISL_NODE_NOT_FOUND, "neighbor id " + std::to_string(neighbors[i]) + " out of range");
  DeviceGuard g(idx->device);
  std::unique_lock<std::shared_mutex> lock(idx->mu);
  const uint64_t s = idx->h_offsets[node_id], e = idx->h_offsets[node_id + 1];
  if (count == e - s) {
    std::copy(neighbors, neighbors + count, idx->h_nbrs.begin() + s);
  } else {
    std::vector<uint64_t> nb;
    nb.reserve(idx->h_nbrs.size() + count - (e - s));
    nb.insert(nb.end(), idx->h_nbrs.begin(), idx->h_nbrs.begin() + s);
    nb.insert(nb.end(), neighbors, neighbors + count);
    nb.insert(nb.end(), idx->h_nbrs.begin() + e, idx->h_nbrs.end());
    idx->h_nbrs.swap(nb);
    const int64_t delta = (int64_t)count - (int64_t)(e - s);
    for (uint64_t i = node_id + 1; i <= idx->n; ++i) idx->h_offsets[i] = (uint64_t)((int64_t)idx->h_offsets[i] + delta);
  }
  return index_finish_graph(idx);
} ISL_ABI_GUARD

isl_status isl_index_search(const isl_index* idx, const float* queries, uint64_t nq,
                            uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                            float* out_dist, uint32_t* out_count, isl_search_stats* stats) try {
  bool trivial;
  ISL_TRY(search_checks(idx, queries, nq, query_dim, k, &ef, &trivial));
  if (trivial) {
    fill_empty(nq, k, out_ids, out_dist, out_count, stats);
    return ISL_OK;
  }
  if (!out_ids || !out_dist) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  DeviceGuard g(idx->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  cudaStream_t st = sc->stream;
  ISL_TRY(ensure(sc->q_stage, nq * idx->ld));
  ISL_TRY(ensure(sc->out_ids, nq * k));
  ISL_TRY(ensure(sc->out_dist, nq * k));
  ISL_TRY(ensure(sc->out_count, nq));
  if (stats) ISL_TRY(ensure(sc->out_stats, nq));
  if (idx->ld != idx->dim) ISL_CUDA_TRY(cudaMemsetAsync(sc->q_stage.p, 0, nq * idx->ld * 4, st));
  ISL_CUDA_TRY(cudaMemcpy2DAsync(sc->q_stage.p, (size_t)idx->ld * 4, queries, (size_t)idx->dim * 4,
                                 (size_t)idx->dim * 4, nq, cudaMemcpyHostToDevice, st));
  ISL_TRY(search_device(idx, sc.get(), sc->q_stage.p, idx->ld, nq, k, ef, sc->out_ids.p, sc->out_dist.p,
                        sc->out_count.p, stats ? sc->out_stats.p : nullptr, nullptr));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, sc->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, sc->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, sc->out_count.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (stats)
    ISL_CUDA_TRY(cudaMemcpyAsync(stats, sc->out_stats.p, nq * sizeof(isl_search_stats), cudaMemcpyDeviceToHost, st));
  return search_finish(idx, sc.get(), 1);
} ISL_ABI_GUARD

isl_status isl_index_search_dev(const isl_index* idx, const float* d_queries, uint64_t nq,
                                uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                                float* d_out_dist, uint32_t* d_out_count,
                                isl_search_stats* d_stats) try {
  bool trivial;
  ISL_TRY(search_checks(idx, d_queries, nq, query_dim, k, &ef, &trivial));
  if (!d_out_ids || !d_out_dist) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  if (trivial) return fill_empty_dev(nq, k, d_out_ids, d_out_dist, d_out_count, d_stats);
  DeviceGuard g(idx->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  // the caller's buffers may still be being written on the caller's stream (isl_set_caller_stream)
  ISL_TRY(order_after_caller(sc->stream, sc->ev_in));
  const float* q;
  uint32_t q_ld;
  ISL_TRY(stage_device_queries(idx, sc.get(), d_queries, nq, query_dim, &q, &q_ld));
  ISL_TRY(search_device(idx, sc.get(), q, q_ld, nq, k, ef, d_out_ids, d_out_dist, d_out_count, d_stats, nullptr));
  return search_finish(idx, sc.get(), 1);
} ISL_ABI_GUARD

isl_status isl_index_search_default(const isl_index* idx, const float* queries, uint64_t nq,
                                    uint32_t query_dim, uint32_t k, uint64_t* out_ids,
                                    float* out_dist, uint32_t* out_count) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  return isl_index_search(idx, queries, nq, query_dim, k, (uint32_t)idx->cfg.ef_search, out_ids,
                          out_dist, out_count, nullptr);
} ISL_ABI_GUARD

isl_status isl_index_last_search_timing(const isl_index* idx, float* kernel_ms, uint64_t* kernel_launches) try {
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  std::lock_guard<std::mutex> lock(idx->pool_mu);
  if (kernel_ms) *kernel_ms = idx->last_kernel_ms;
  if (kernel_launches) *kernel_launches = idx->last_launches;
  return ISL_OK;
} ISL_ABI_GUARD

// ---- merge ----------------------------------------------------------------------------------
isl_status isl_merge_topk_dev(const uint64_t* d_ids, const float* d_dist, uint32_t parts, uint64_t nq,
                              uint32_t k, uint64_t* d_out_ids, float* d_out_dist,
                              uint32_t* d_out_count) try {
  if (nq == 0 || k == 0) return ISL_OK;
  if (!d_ids || !d_dist || !d_out_ids || !d_out_dist) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (parts == 0) return fail(ISL_INVALID_ARGUMENT, "parts must be > 0");
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  ISL_TRY(launch_merge_topk(d_ids, d_dist, parts, nq, k, d_out_ids, d_out_dist, d_out_count, 0));
  ISL_CUDA_TRY(cudaStreamSynchronize(0));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_merge_topk(const uint64_t* ids, const float* dist, uint32_t parts, uint64_t nq,
                          uint32_t k, uint64_t* out_ids, float* out_dist, uint32_t* out_count) try {
  if (nq == 0 || k == 0) return ISL_OK;
  if (!ids || !dist || !out_ids || !out_dist) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (parts == 0) return fail(ISL_INVALID_ARGUMENT, "parts must be > 0");
  const size_t total = (size_t)parts * nq * k;
  DevBuf<uint64_t> di, doi;
  DevBuf<float> dd, dod;
  DevBuf<uint32_t> dc;
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  ISL_CUDA_TRY(di.alloc(total));
  ISL_CUDA_TRY(dd.alloc(total));
  ISL_CUDA_TRY(doi.alloc(nq * k));
  ISL_CUDA_TRY(dod.alloc(nq * k));
  ISL_CUDA_TRY(dc.alloc(nq));
  ISL_CUDA_TRY(cudaMemcpy(di.p, ids, total * 8, cudaMemcpyHostToDevice));
  ISL_CUDA_TRY(cudaMemcpy(dd.p, dist, total * 4, cudaMemcpyHostToDevice));
  ISL_TRY(launch_merge_topk(di.p, dd.p, parts, nq, k, doi.p, dod.p, dc.p, 0));
  ISL_CUDA_TRY(cudaMemcpy(out_ids, doi.p, nq * k * 8, cudaMemcpyDeviceToHost));
  ISL_CUDA_TRY(cudaMemcpy(out_dist, dod.p, nq * k * 4, cudaMemcpyDeviceToHost));
  if (out_count) ISL_CUDA_TRY(cudaMemcpy(out_count, dc.p, nq * 4, cudaMemcpyDeviceToHost));
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
