// islands_b200.hpp — C++17 host-side mirror of the reference's `src/core` API over the C ABI.
//
// The reference is compiled Rust; its toolchain is not available in this image, so the host
// side above the C ABI is C++ (header only) with the reference's names, argument meaning and
// error behaviour:
//   islands::DistanceMetric / calculate / batch_calculate   src/core/distance.rs:7-67
//   islands::CoreError (+ kind)                             src/core/error.rs:9-62
//   islands::LeannConfig / LeannIndex / CsrGraph            src/core/leann.rs:193-302, 322-461, 493-1067
//   islands::HnswConfig / HnswGraph                       src/core/hnsw.rs:15-86, 151-531
//   islands::EncoderConfig / Encoder                     src/core/embedding/candle_provider.rs:353-507
//   islands::PQConfig / ProductQuantizer                    src/core/pq.rs:13-65, 116-359
//   islands::SearchConfig / SearchResult / Searcher / MultiIndexSearcher   src/core/search.rs:9-249
//   islands::ShardComm (+ LeannIndex::search_sharded)       src/indexer/service.rs:777-801
// Link with -lislands_b200 (islands_b200/lib).  All compute runs on the GPU; nothing here has a
// CPU fallback.
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "islands_b200.h"

namespace islands {

enum class ErrorKind {
  DimensionMismatch = ISL_DIM_MISMATCH,
  EmptyCollection = ISL_EMPTY_COLLECTION,
  InvalidConfig = ISL_INVALID_CONFIG,
  IndexNotBuilt = ISL_INDEX_NOT_BUILT,
  NodeNotFound = ISL_NODE_NOT_FOUND,
  PQError = ISL_PQ_ERROR,
  Serialization = ISL_SERIALIZATION,
  Cuda = ISL_CUDA_ERROR,
  InvalidArgument = ISL_INVALID_ARGUMENT,
};

// CoreError (error.rs:9-62): the variant is `kind`, the String payload is what().
class CoreError : public std::runtime_error {
 public:
  CoreError(ErrorKind k, const std::string& msg) : std::runtime_error(msg), kind(k) {
    isl_last_error_detail(&a, &b);  // DimensionMismatch{expected = a, actual = b}, NodeNotFound(a)
  }
  uint64_t a = 0, b = 0;
  ErrorKind kind;
};

inline void check(isl_status st) {
  if (st != ISL_OK) throw CoreError(static_cast<ErrorKind>(st), isl_last_error());
}

// ---- distance.rs ------------------------------------------------------------------------------
enum class DistanceMetric : int32_t { Cosine = 0, Euclidean = 1, DotProduct = 2, Manhattan = 3 };

inline float calculate(DistanceMetric m, const std::vector<float>& a, const std::vector<float>& b) {
  float out = 0.0f;
  check(isl_distance_calculate(static_cast<int32_t>(m), a.data(), a.size(), b.data(), b.size(), &out));
  return out;
}
inline float calculate_squared(DistanceMetric m, const std::vector<float>& a, const std::vector<float>& b) {
  float out = 0.0f;
  check(isl_distance_calculate_squared(static_cast<int32_t>(m), a.data(), a.size(), b.data(), b.size(), &out));
  return out;
}
// rows: [n][query.size()] row-major
inline std::vector<float> batch_calculate(DistanceMetric m, const std::vector<float>& query,
                                          const std::vector<float>& rows) {
  const uint64_t n = query.empty() ? 0 : rows.size() / query.size();
  std::vector<float> out(n);
  check(isl_distance_batch(static_cast<int32_t>(m), query.data(), rows.data(), n, (uint32_t)query.size(), out.data()));
  return out;
}

// ---- leann.rs -----------------------------------------------------------------------------------
struct LeannConfig : isl_leann_config {
  LeannConfig() { check(isl_leann_config_default(this)); }  // paper_default (leann.rs:386-403)
  static LeannConfig paper_default() { return LeannConfig(); }
  static LeannConfig fast() {
    LeannConfig c;
    check(isl_leann_config_fast(&c));
    return c;
  }
  static LeannConfig accurate() {
    LeannConfig c;
    check(isl_leann_config_accurate(&c));
    return c;
  }
  void validate() const { check(isl_leann_config_validate(this)); }
};

struct CsrGraph {  // leann.rs:193-208
  std::vector<uint64_t> node_offsets{0}, neighbors, levels, degree_counts;
  int64_t entry_point = ISL_NO_ENTRY;
  uint64_t max_level = 0, num_nodes = 0;
  // CsrGraph::set_neighbors (leann.rs:256-293): same length overwrites, any other length rebuilds the arrays
  void set_neighbors(uint64_t node_id, const std::vector<uint64_t>& nb) {
    if (node_id >= num_nodes) return;
    const uint64_t s = node_offsets[node_id], e = node_offsets[node_id + 1];
    if (nb.size() == e - s) {
      std::copy(nb.begin(), nb.end(), neighbors.begin() + s);
    } else {
      const int64_t delta = (int64_t)nb.size() - (int64_t)(e - s);
      neighbors.erase(neighbors.begin() + s, neighbors.begin() + e);
      neighbors.insert(neighbors.begin() + s, nb.begin(), nb.end());
      for (uint64_t i = node_id + 1; i <= num_nodes; ++i) node_offsets[i] = (uint64_t)((int64_t)node_offsets[i] + delta);
    }
    if (node_id < degree_counts.size()) degree_counts[node_id] = nb.size();
  }
};

using SearchResults = std::vector<std::pair<uint64_t, float>>;  // Vec<(u64, f32)>

// Batched results: ids / dist are [nq][k] row-major (padded with ISL_INVALID_ID / +inf), count[q] <= k entries are valid.
struct BatchResults {
  uint32_t k = 0;
  std::vector<uint64_t> ids;
  std::vector<float> dist;
  std::vector<uint32_t> count;
  BatchResults(uint64_t nq, uint32_t k_) : k(k_), ids(nq * k_, ISL_INVALID_ID), dist(nq * k_, 0.0f), count(nq, 0) {}
  SearchResults row(uint64_t q) const {
    SearchResults out;
    for (uint32_t i = 0; i < count[q]; ++i) out.emplace_back(ids[q * k + i], dist[q * k + i]);
    return out;
  }
};

namespace detail {
// The two-call protocol of the *_to_bytes entry points: size query, then fill.
template <class Handle, class Fn>
std::vector<uint8_t> bytes_of(const Handle* h, Fn to_bytes) {
  uint64_t len = 0;
  check(to_bytes(h, nullptr, 0, &len));
  std::vector<uint8_t> out(len);
  check(to_bytes(h, out.data(), len, &len));
  return out;
}
}  // namespace detail

// One rank's membership in a sharded index: NCCL communicator inside the library (csrc/api_shard.cu).
class ShardComm {
 public:
  using UniqueId = std::array<uint8_t, 128>;
  // ncclGetUniqueId: call on one rank, hand the 128 bytes to the others by any host channel.
  static UniqueId unique_id() {
    UniqueId id{};
    check(isl_shard_unique_id(id.data(), id.size()));
    return id;
  }
  ShardComm(int rank, int world, const UniqueId& id) { check(isl_shard_init(rank, world, id.data(), &h_)); }
  ~ShardComm() { isl_shard_free(h_); }
  ShardComm(const ShardComm&) = delete;
  ShardComm& operator=(const ShardComm&) = delete;
  int rank() const { return isl_shard_rank(h_); }
  int world() const { return isl_shard_world(h_); }
  // Exchange by peer stores over NVLink instead of ncclAllGather (collective; ranks of one node).
  void enable_peer_exchange(uint64_t max_records) { check(isl_shard_enable_peer_exchange(h_, max_records)); }
  isl_shard* handle() const { return h_; }

 private:
  isl_shard* h_ = nullptr;
};

class LeannIndex {
 public:
  explicit LeannIndex(const LeannConfig& cfg = LeannConfig()) : cfg_(cfg) { cfg_.validate(); }
  ~LeannIndex() { isl_index_free(h_); }
  LeannIndex(const LeannIndex&) = delete;
  LeannIndex& operator=(const LeannIndex&) = delete;

  // LeannIndex::build (leann.rs:560-631).  `embeddings` plays the InMemoryEmbeddingProvider:
  // row i is compute_embedding(i).  `levels` (optional) replaces the thread_rng draw.
  void build(const std::vector<float>& embeddings, uint32_t dim, uint64_t num_vectors,
             const std::vector<uint64_t>* levels = nullptr, uint64_t seed = 0, uint32_t batch = 1) {
    reset();
    check(isl_index_build(&cfg_, dim, num_vectors, embeddings.data(), levels ? levels->data() : nullptr, seed, batch,
                          &h_));
  }
  void from_csr(const CsrGraph& g, const std::vector<float>& embeddings, uint32_t dim) {
    reset();
    check(isl_index_from_csr(&cfg_, dim, g.num_nodes, g.node_offsets.data(), g.neighbors.data(),
                             g.levels.empty() ? nullptr : g.levels.data(), g.entry_point, embeddings.data(), &h_));
  }
  uint64_t len() const { return isl_index_len(h_); }
  bool is_empty() const { return len() == 0; }
  uint32_t dimension() const { return isl_index_dimension(h_); }
  uint64_t storage_bytes() const { return isl_index_storage_bytes(h_); }
  // graph.set_neighbors (leann.rs:256-293) on the resident graph
  void set_neighbors(uint64_t node_id, const std::vector<uint64_t>& new_neighbors) {
    check(isl_index_set_neighbors(h_, node_id, new_neighbors.data(), new_neighbors.size()));
  }
  isl_build_stats last_build_stats() const {
    isl_build_stats st{};
    check(isl_index_last_build_stats(h_, &st));
    return st;
  }
  // Sharded search (service.rs:777-801): collective over the ranks of `shard`; queries [nq][dim] row-major,
  // the same on every rank; ids / dist come back as [nq][k] (padded with ISL_INVALID_ID / +inf).
  void search_sharded(isl_shard* shard, uint64_t id_base, const std::vector<float>& queries, uint64_t nq, uint32_t k,
                      uint32_t ef, std::vector<uint64_t>* ids, std::vector<float>* dist, std::vector<uint32_t>* count) const {
    ids->assign(nq * k, ISL_INVALID_ID);
    dist->assign(nq * k, 0.0f);
    count->assign(nq, 0);
    check(isl_index_search_sharded(h_, shard, id_base, queries.data(), nq, nq ? (uint32_t)(queries.size() / nq) : 0, k, ef,
                                   ids->data(), dist->data(), count->data()));
  }

  // LeannIndex::search / search_with_params (leann.rs:858-896)
  SearchResults search(const std::vector<float>& query, uint32_t k) const {
    return search_with_params(query, k, (uint32_t)cfg_.ef_search);
  }
  SearchResults search_with_params(const std::vector<float>& query, uint32_t k, uint32_t ef) const {
    std::vector<uint64_t> ids(k);
    std::vector<float> dist(k);
    uint32_t count = 0;
    check(isl_index_search(h_, query.data(), 1, (uint32_t)query.size(), k, ef, ids.data(), dist.data(), &count, nullptr));
    SearchResults out;
    for (uint32_t i = 0; i < count; ++i) out.emplace_back(ids[i], dist[i]);
    return out;
  }
  // Batched form (what a GPU wants): queries [nq][dim] row-major, one kernel launch over all of them.
  BatchResults search_batch(const std::vector<float>& queries, uint64_t nq, uint32_t k, uint32_t ef) const {
    BatchResults r(nq, k);
    check(isl_index_search(h_, queries.data(), nq, nq ? (uint32_t)(queries.size() / nq) : 0, k, ef, r.ids.data(), r.dist.data(),
                           r.count.data(), nullptr));
    return r;
  }
  BatchResults search_sharded(const ShardComm& shard, uint64_t id_base, const std::vector<float>& queries, uint64_t nq,
                              uint32_t k, uint32_t ef) const {
    BatchResults r(nq, k);
    search_sharded(shard.handle(), id_base, queries, nq, k, ef, &r.ids, &r.dist, &r.count);
    return r;
  }
  // to_bytes / from_bytes (leann.rs:1059-1066): the bincode image holds the graph only, so from_bytes also takes the
  // embeddings [n][dim] the provider returns.
  std::vector<uint8_t> to_bytes() const { return detail::bytes_of(h_, isl_index_to_bytes); }
  void from_bytes(const std::vector<uint8_t>& bytes, const std::vector<float>& embeddings, uint32_t dim) {
    reset();
    check(isl_index_from_bytes(bytes.data(), bytes.size(), embeddings.data(), dim, &h_));
    check(isl_index_get_config(h_, &cfg_));
  }
  // InMemoryEmbeddingProvider::compute_embedding (leann.rs:141-150) of the resident embeddings
  std::vector<float> get_vector(uint64_t node_id) const {
    std::vector<float> out(dimension());
    check(isl_index_get_vector(h_, node_id, out.data()));
    return out;
  }
  // On-demand recompute (leann.rs:82-99, 947-950): `enc` over one token row per node [n][seq_len] plays the
  // EmbeddingProvider; drop_vectors frees the resident embeddings (graph + codes + token rows remain).
  void set_recompute(isl_encoder* enc, const std::vector<int32_t>& token_ids, const std::vector<int32_t>& lengths,
                     uint32_t seq_len) {
    check(isl_index_set_recompute(h_, enc, token_ids.data(), lengths.data(), seq_len));
  }
  void drop_vectors() { check(isl_index_drop_vectors(h_)); }
  void set_hub_cache(uint64_t count) { check(isl_index_set_hub_cache(h_, count)); }
  // the reference's loop: every hop's unvisited neighbours are embedded by the provider (one encoder pass per frontier)
  BatchResults search_recompute(const std::vector<float>& queries, uint64_t nq, uint32_t k, uint32_t ef) const {
    BatchResults r(nq, k);
    check(isl_index_search_recompute(h_, queries.data(), nq, nq ? (uint32_t)(queries.size() / nq) : 0, k, ef, r.ids.data(),
                                     r.dist.data(), r.count.data(), nullptr));
    return r;
  }
  // the cheap variant: ADC traversal -> encoder over the distinct survivors -> exact rerank
  BatchResults search_adc_recompute(const std::vector<float>& queries, uint64_t nq, uint32_t k, uint32_t ef) const {
    BatchResults r(nq, k);
    check(isl_index_search_adc_recompute(h_, queries.data(), nq, nq ? (uint32_t)(queries.size() / nq) : 0, k, ef,
                                         r.ids.data(), r.dist.data(), r.count.data(), nullptr));
    return r;
  }
  CsrGraph graph() const {
    CsrGraph g;
    g.num_nodes = len();
    g.node_offsets.assign(g.num_nodes + 1, 0);
    g.neighbors.assign(isl_index_num_edges(h_), 0);
    g.levels.assign(g.num_nodes, 0);
    g.degree_counts.assign(g.num_nodes, 0);
    check(isl_index_export_csr(h_, g.node_offsets.data(), g.neighbors.data(), g.levels.data(), g.degree_counts.data()));
    g.entry_point = isl_index_entry_point(h_);
    g.max_level = isl_index_max_level(h_);
    return g;
  }
  // Two-level search family (docs/leann-specification.md:223-269): PQ codes [n][m] attached to the index, then
  // either the AQ-promotion search (rerank ratio `a`) or the traversal on table distances + exact rerank.
  void attach_pq(const isl_pq* pq, const std::vector<uint16_t>& codes) { check(isl_index_attach_pq(h_, pq, codes.data())); }
  void set_rerank_limit(uint32_t limit) { check(isl_index_set_rerank_limit(h_, limit)); }
  SearchResults search_two_level(const std::vector<float>& query, uint32_t k, uint32_t ef, float rerank_ratio) const {
    std::vector<uint64_t> ids(k);
    std::vector<float> dist(k);
    uint32_t count = 0;
    check(isl_index_search_two_level(h_, query.data(), 1, (uint32_t)query.size(), k, ef, rerank_ratio, ids.data(), dist.data(),
                                     &count, nullptr));
    SearchResults out;
    for (uint32_t i = 0; i < count; ++i) out.emplace_back(ids[i], dist[i]);
    return out;
  }
  SearchResults search_adc_rerank(const std::vector<float>& query, uint32_t k, uint32_t ef) const {
    std::vector<uint64_t> ids(k);
    std::vector<float> dist(k);
    uint32_t count = 0;
    check(isl_index_search_adc_rerank(h_, query.data(), 1, (uint32_t)query.size(), k, ef, ids.data(), dist.data(), &count,
                                      nullptr));
    SearchResults out;
    for (uint32_t i = 0; i < count; ++i) out.emplace_back(ids[i], dist[i]);
    return out;
  }
  isl_index* handle() const { return h_; }

 private:
  void reset() {
    isl_index_free(h_);
    h_ = nullptr;
  }
  LeannConfig cfg_;
  isl_index* h_ = nullptr;
};

// ---- hnsw.rs ------------------------------------------------------------------------------------
struct HnswConfig : isl_hnsw_config {
  HnswConfig() { check(isl_hnsw_config_default(this)); }  // hnsw.rs:37-48
  void validate() const { check(isl_hnsw_config_validate(this)); }
};

struct HnswNode {  // hnsw.rs:88-125, materialised from the device on request
  uint64_t id = 0;
  std::vector<float> vector;
  std::vector<std::vector<uint64_t>> connections;  // [layer] -> neighbour ids, layers 0..level
  uint64_t level = 0;
  const std::vector<uint64_t>* neighbors_at(uint64_t layer) const {
    return layer < connections.size() ? &connections[layer] : nullptr;
  }
};

class HnswGraph {  // hnsw.rs:151-531
 public:
  explicit HnswGraph(const HnswConfig& cfg = HnswConfig()) { check(isl_hnsw_new(&cfg, &h_)); }
  // HnswGraph::from_bytes (hnsw.rs:512-514)
  explicit HnswGraph(const std::vector<uint8_t>& bytes) { check(isl_hnsw_from_bytes(bytes.data(), bytes.size(), &h_)); }
  std::vector<uint8_t> to_bytes() const { return detail::bytes_of(h_, isl_hnsw_to_bytes); }  // hnsw.rs:507-509
  HnswConfig config() const {
    HnswConfig c;
    check(isl_hnsw_get_config(h_, &c));
    return c;
  }
  // get_node (hnsw.rs:201-203): nullopt for an unknown id
  std::optional<HnswNode> get_node(uint64_t id) const {
    HnswNode node;
    node.id = id;
    if (id >= len() || isl_hnsw_node_level(h_, id, &node.level) != ISL_OK) return std::nullopt;
    node.vector.resize(dimension());
    check(isl_hnsw_get_vector(h_, id, node.vector.data()));
    for (uint64_t layer = 0; layer <= node.level; ++layer) node.connections.push_back(neighbors_at(id, layer));
    return node;
  }
  BatchResults search_batch(const std::vector<float>& queries, uint64_t nq, uint32_t k, uint32_t ef) const {
    BatchResults r(nq, k);
    if (nq == 0 || is_empty()) return r;  // hnsw.rs:459-461
    check(isl_hnsw_search(h_, queries.data(), nq, (uint32_t)(queries.size() / nq), k, ef, r.ids.data(), r.dist.data(),
                          r.count.data()));
    return r;
  }
  ~HnswGraph() { isl_hnsw_free(h_); }
  HnswGraph(const HnswGraph&) = delete;
  HnswGraph& operator=(const HnswGraph&) = delete;
  uint64_t len() const { return isl_hnsw_len(h_); }
  bool is_empty() const { return len() == 0; }
  uint32_t dimension() const { return isl_hnsw_dimension(h_); }
  int64_t entry_point() const { return isl_hnsw_entry_point(h_); }
  uint64_t max_level() const { return isl_hnsw_max_level(h_); }
  // HnswGraph::insert (hnsw.rs:214-250) -> id; `level` (optional) replaces the thread_rng draw.
  uint64_t insert(const std::vector<float>& v, const uint64_t* level = nullptr, uint64_t seed = 0) {
    uint64_t id = 0;
    check(isl_hnsw_insert_batch(h_, v.data(), 1, (uint32_t)v.size(), level, seed, 1, &id));
    return id;
  }
  // `count` inserts, up to `batch` nodes per graph snapshot (GPU-parallel construction).
  uint64_t insert_batch(const std::vector<float>& vectors, uint32_t dim, const std::vector<uint64_t>* levels = nullptr,
                        uint64_t seed = 0, uint32_t batch = 1) {
    uint64_t first = 0;
    check(isl_hnsw_insert_batch(h_, vectors.data(), dim ? vectors.size() / dim : 0, dim,
                                levels ? levels->data() : nullptr, seed, batch, &first));
    return first;
  }
  // neighbors_at(layer) of get_node(id) (hnsw.rs:108-110, 201-203)
  std::vector<uint64_t> neighbors_at(uint64_t id, uint64_t layer) const {
    uint64_t cnt = 0;
    check(isl_hnsw_get_neighbors(h_, id, layer, nullptr, 0, &cnt));
    std::vector<uint64_t> out(cnt);
    check(isl_hnsw_get_neighbors(h_, id, layer, out.data(), cnt, &cnt));
    return out;
  }
  // HnswGraph::search (hnsw.rs:458-504)
  SearchResults search(const std::vector<float>& query, uint32_t k, uint32_t ef) const {
    std::vector<uint64_t> ids(k);
    std::vector<float> dist(k);
    uint32_t count = 0;
    check(isl_hnsw_search(h_, query.data(), 1, (uint32_t)query.size(), k, ef, ids.data(), dist.data(), &count));
    SearchResults out;
    for (uint32_t i = 0; i < count; ++i) out.emplace_back(ids[i], dist[i]);
    return out;
  }
  isl_hnsw* handle() const { return h_; }

 private:
  isl_hnsw* h_ = nullptr;
};

// ---- pq.rs ----------------------------------------------------------------------------------------
struct PQConfig : isl_pq_config {
  PQConfig() { check(isl_pq_config_default(this)); }
  void validate(uint64_t dimension) const { check(isl_pq_config_validate(this, dimension)); }
  uint64_t bytes_per_vector() const { return isl_pq_config_bytes_per_vector(this); }
};

class ProductQuantizer {
 public:
  ProductQuantizer(uint32_t dimension, const PQConfig& cfg = PQConfig()) : dim_(dimension) {
    check(isl_pq_new(dimension, &cfg, &h_));
  }
  ~ProductQuantizer() { isl_pq_free(h_); }
  ProductQuantizer(const ProductQuantizer&) = delete;
  ProductQuantizer& operator=(const ProductQuantizer&) = delete;
  ProductQuantizer& with_metric(DistanceMetric m) {
    check(isl_pq_set_metric(h_, static_cast<int32_t>(m)));
    return *this;
  }
  bool is_trained() const { return isl_pq_is_trained(h_) != 0; }
  uint64_t num_subquantizers() const { return isl_pq_num_subquantizers(h_); }
  float compression_ratio() const { return isl_pq_compression_ratio(h_); }
  void train(const std::vector<float>& vectors) { check(isl_pq_train(h_, vectors.data(), vectors.size() / dim_, dim_)); }
  std::vector<uint16_t> encode(const std::vector<float>& v) const {
    std::vector<uint16_t> codes(num_subquantizers());
    check(isl_pq_encode(h_, v.data(), 1, (uint32_t)v.size(), codes.data()));
    return codes;
  }
  std::vector<uint16_t> encode_batch(const std::vector<float>& vectors) const {  // [n][dim] -> [n][m]
    const uint64_t n = vectors.size() / dim_;
    std::vector<uint16_t> codes(n * num_subquantizers());
    check(isl_pq_encode(h_, vectors.data(), n, dim_, codes.data()));
    return codes;
  }
  std::vector<float> decode(const std::vector<uint16_t>& codes) const {
    std::vector<float> out(dim_);
    check(isl_pq_decode(h_, codes.data(), 1, codes.size(), out.data()));
    return out;
  }
  float asymmetric_distance(const std::vector<float>& q, const std::vector<uint16_t>& codes) const {
    float out = 0.0f;
    check(isl_pq_asymmetric_distance(h_, q.data(), (uint32_t)q.size(), codes.data(), 1, &out));
    return out;
  }
  // build_distance_tables (pq.rs:307-338): [m][ksub] row-major; table_distance (pq.rs:341-348) for one code row
  std::vector<float> build_distance_tables(const std::vector<float>& q) const {
    uint64_t ksub = 0;
    check(isl_pq_get_codebooks(h_, nullptr, &ksub));
    std::vector<float> tables(num_subquantizers() * ksub);
    check(isl_pq_build_tables(h_, q.data(), (uint32_t)q.size(), tables.data()));
    return tables;
  }
  float table_distance(const std::vector<float>& tables, const std::vector<uint16_t>& codes) const {
    float out = 0.0f;
    check(isl_pq_table_distance(h_, tables.data(), codes.data(), 1, &out));
    return out;
  }
  // to_bytes / from_bytes (pq.rs:351-358)
  std::vector<uint8_t> to_bytes() const { return detail::bytes_of(h_, isl_pq_to_bytes); }
  explicit ProductQuantizer(const std::vector<uint8_t>& bytes) : dim_(0) {
    check(isl_pq_from_bytes(bytes.data(), bytes.size(), &h_));
    dim_ = isl_pq_dimension(h_);
  }
  isl_pq* handle() const { return h_; }

 private:
  uint32_t dim_;
  isl_pq* h_ = nullptr;
};

// ---- embedding/candle_provider.rs ---------------------------------------------------------------
struct EncoderConfig : isl_encoder_config {
  EncoderConfig() { check(isl_encoder_config_default(this)); }  // BERT-base shape
};

// The model behind CandleEmbedder::embed_texts_raw (candle_provider.rs:353-507) for token ids.
class Encoder {
 public:
  explicit Encoder(const EncoderConfig& cfg = EncoderConfig()) { check(isl_encoder_new(&cfg, &h_)); }
  ~Encoder() { isl_encoder_free(h_); }
  Encoder(const Encoder&) = delete;
  Encoder& operator=(const Encoder&) = delete;
  uint32_t dimension() const { return isl_encoder_dimension(h_); }  // EmbeddingProvider::dimension
  void init_random(uint64_t seed = 46, float stddev = 0.02f) { check(isl_encoder_init_random(h_, seed, stddev)); }
  void set_parameter(const std::string& name, const std::vector<float>& data) {
    check(isl_encoder_set_parameter(h_, name.c_str(), data.data(), data.size()));
  }
  // token_ids [batch][seq_len] (0-padded), lengths [batch] -> [batch][dimension()] pooled, normalised
  std::vector<float> embed(const std::vector<int32_t>& token_ids, const std::vector<int32_t>& lengths, uint32_t seq_len) {
    std::vector<float> out(lengths.size() * (size_t)dimension());
    check(isl_encoder_embed(h_, token_ids.data(), lengths.data(), lengths.size(), seq_len, out.data()));
    return out;
  }
  isl_encoder* handle() const { return h_; }

 private:
  isl_encoder* h_ = nullptr;
};

// ---- search.rs ----------------------------------------------------------------------------------
inline float to_similarity(float score) { return 1.0f / (1.0f + score); }  // search.rs:99-102

struct SearchConfig {  // search.rs:9-52
  uint32_t top_k = 10, ef = 100;
  bool include_vectors = false, include_metadata = true;
  std::optional<float> min_similarity;
  static SearchConfig fast(uint32_t k) {
    SearchConfig c;
    c.top_k = k, c.ef = k * 2;
    return c;
  }
  static SearchConfig accurate(uint32_t k) {
    SearchConfig c;
    c.top_k = k, c.ef = k * 10;
    return c;
  }
};

struct SearchResult {  // search.rs:54-103 (metadata / text are the caller's: the index stores neither)
  uint64_t id = 0;
  float score = 0.0f;
  std::optional<std::vector<float>> vector;
  std::optional<std::string> text;
  float to_similarity() const { return islands::to_similarity(score); }
};

namespace detail {
// (id, distance) rows of one graph -> SearchResults under `cfg`: vectors on request, then the min_similarity filter
// (search.rs:155-176).
inline std::vector<SearchResult> decorate(const HnswGraph& g, const SearchConfig& cfg, const SearchResults& rows,
                                          bool filter) {
  std::vector<SearchResult> out;
  for (const auto& [id, d] : rows) {
    SearchResult r;
    r.id = id, r.score = d;
    if (filter && cfg.min_similarity && r.to_similarity() < *cfg.min_similarity) continue;
    if (cfg.include_vectors)
      if (auto node = g.get_node(id)) r.vector = std::move(node->vector);
    out.push_back(std::move(r));
  }
  return out;
}
}  // namespace detail

class Searcher {  // search.rs:106-182
 public:
  explicit Searcher(const HnswGraph& graph, SearchConfig cfg = SearchConfig()) : g_(graph), cfg_(cfg) {}
  Searcher& top_k(uint32_t k) { return cfg_.top_k = k, *this; }
  Searcher& ef(uint32_t ef) { return cfg_.ef = ef, *this; }
  Searcher& include_vectors() { return cfg_.include_vectors = true, *this; }
  Searcher& min_similarity(float t) { return cfg_.min_similarity = t, *this; }
  std::vector<SearchResult> search(const std::vector<float>& query) const {
    if (g_.is_empty()) return {};
    return detail::decorate(g_, cfg_, g_.search(query, cfg_.top_k, cfg_.ef), true);
  }
  // search.rs:179-181 maps `search` over the queries; here the batch is one library call over [nq][dim] queries.
  std::vector<std::vector<SearchResult>> search_batch(const std::vector<float>& queries, uint64_t nq) const {
    std::vector<std::vector<SearchResult>> out;
    const BatchResults r = g_.search_batch(queries, nq, cfg_.top_k, cfg_.ef);
    for (uint64_t q = 0; q < nq; ++q) out.push_back(detail::decorate(g_, cfg_, r.row(q), true));
    return out;
  }

 private:
  const HnswGraph& g_;
  SearchConfig cfg_;
};

// search.rs:185-254: every island is searched, the lists are merged by score with the island order breaking ties
// (the reference's stable sort) and cut to top_k; min_similarity is not applied, as in the reference.
class MultiIndexSearcher {
 public:
  void add_index(std::string name, const HnswGraph* graph) { graphs_.emplace_back(std::move(name), graph); }
  MultiIndexSearcher& with_config(SearchConfig cfg) { return cfg_ = cfg, *this; }
  std::vector<std::pair<std::string, SearchResult>> search(const std::vector<float>& query) const {
    std::vector<std::pair<std::string, SearchResult>> all;
    for (const auto& [name, g] : graphs_) {
      if (g->is_empty()) continue;
      for (auto& r : detail::decorate(*g, cfg_, g->search(query, cfg_.top_k, cfg_.ef), false)) all.emplace_back(name, std::move(r));
    }
    std::stable_sort(all.begin(), all.end(), [](const auto& a, const auto& b) { return a.second.score < b.second.score; });
    if (all.size() > cfg_.top_k) all.resize(cfg_.top_k);
    return all;
  }
  size_t num_indexes() const { return graphs_.size(); }
  uint64_t total_vectors() const {
    uint64_t n = 0;
    for (const auto& g : graphs_) n += g.second->len();
    return n;
  }

 private:
  std::vector<std::pair<std::string, const HnswGraph*>> graphs_;  // not owned: handles are not copyable
  SearchConfig cfg_;
};

}  // namespace islands
