// Compiled and run by tests/test_cpp_host.py: the C++ host mirror (include/islands_b200.hpp)
// links against the C ABI, keeps the reference's defaults / error behaviour, and — on a box
// with a GPU ("gpu" argument) — builds and searches a small index.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "islands_b200.hpp"

#define EXPECT(c)                                             \
  do {                                                        \
    if (!(c)) {                                               \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
      return 1;                                               \
    }                                                         \
  } while (0)

int main(int argc, char** argv) {
  using namespace islands;
  LeannConfig c;
  EXPECT(c.m == 30 && c.m0 == 60 && c.ef_construction == 128 && c.ef_search == 64);  // leann.rs:1091-1103
  EXPECT(LeannConfig::fast().prune_ratio > 0.0f && LeannConfig::accurate().m0 == 96);
  LeannConfig bad;
  bad.m = 0;
  try {
    bad.validate();
    EXPECT(false);
  } catch (const CoreError& e) {
    EXPECT(e.kind == ErrorKind::InvalidConfig);
  }
  try {  // distance.rs:206-212
    calculate(DistanceMetric::Cosine, {1.f, 2.f}, {1.f, 2.f, 3.f});
    EXPECT(false);
  } catch (const CoreError& e) {
    EXPECT(e.kind == ErrorKind::DimensionMismatch);
  }
  PQConfig pc;
  EXPECT(pc.bytes_per_vector() == 8);
  EncoderConfig ecd;
  EXPECT(ecd.hidden_size == 768 && ecd.num_layers == 12 && ecd.intermediate_size == 3072);
  HnswConfig hc;
  EXPECT(hc.m == 16 && hc.m0 == 32 && hc.ef_construction == 200 && hc.max_layers == 16);  // hnsw.rs:533-545
  EXPECT(std::fabs(to_similarity(1.0f) - 0.5f) < 1e-7f);
  {  // search.rs:9-103 host-side types (reference tests search.rs:262-324)
    SearchConfig sc;
    EXPECT(sc.top_k == 10 && sc.ef == 100 && !sc.include_vectors && sc.include_metadata && !sc.min_similarity);
    EXPECT(SearchConfig::fast(5).ef == 10 && SearchConfig::accurate(5).ef == 50 && SearchConfig::fast(5).top_k == 5);
    SearchResult sr;
    sr.score = 1.0f;
    EXPECT(std::fabs(sr.to_similarity() - 0.5f) < 1e-7f && !sr.vector && !sr.text);
    HnswNode node;
    node.connections = {{1, 2}, {3}};
    EXPECT(node.neighbors_at(1) && node.neighbors_at(1)->size() == 1 && node.neighbors_at(2) == nullptr);  // hnsw.rs:117-119
    BatchResults br(2, 3);
    br.ids[3] = 7, br.dist[3] = 0.25f, br.count[1] = 1;
    EXPECT(br.row(0).empty() && br.row(1).size() == 1 && br.row(1)[0].first == 7 && br.ids[0] == ISL_INVALID_ID);
    MultiIndexSearcher multi;
    EXPECT(multi.num_indexes() == 0 && multi.total_vectors() == 0 && multi.search({1.f, 2.f}).empty());  // search.rs:416-424
  }
  if (argc > 1 && std::strcmp(argv[1], "gpu") == 0) {
    EXPECT(std::fabs(calculate(DistanceMetric::Euclidean, {0.f, 0.f}, {3.f, 4.f}) - 5.0f) < 1e-6f);
    const uint32_t n = 300, d = 16;
    std::vector<float> v(n * d);
    uint32_t s = 12345;
    for (auto& x : v) {
      s = s * 1664525u + 1013904223u;
      x = (float)(s >> 8) / 8388608.0f - 1.0f;
    }
    LeannIndex idx;
    idx.build(v, d, n, nullptr, 1, 8);
    EXPECT(idx.len() == n && idx.dimension() == d);
    std::vector<float> q(v.begin(), v.begin() + d);
    auto r = idx.search(q, 5);
    EXPECT(r.size() == 5 && r[0].first == 0 && r[0].second < 0.01f);
    for (size_t i = 1; i < r.size(); ++i) EXPECT(r[i - 1].second <= r[i].second);
    CsrGraph g = idx.graph();
    EXPECT(g.num_nodes == n && g.node_offsets.back() == g.neighbors.size());
    {  // two-level family: PQ codes attached, traversal on table distances + exact rerank, AQ-promotion search
      PQConfig pc;
      pc.num_subquantizers = 4;
      pc.num_centroids = 16;
      pc.training_iterations = 3;
      pc.seed = 1;
      pc.has_seed = 1;
      ProductQuantizer pq(d, pc);
      pq.train(v);
      idx.attach_pq(pq.handle(), pq.encode_batch(v));
      auto a1 = idx.search_adc_rerank(q, 5, 64);
      EXPECT(a1.size() == 5 && a1[0].first == 0);
      for (size_t i = 1; i < a1.size(); ++i) EXPECT(a1[i - 1].second <= a1[i].second);
      idx.set_rerank_limit(8);
      EXPECT(idx.search_adc_rerank(q, 5, 64).size() == 5);
      idx.set_rerank_limit(0);
      auto a2 = idx.search_two_level(q, 5, 64, 0.5f);
      EXPECT(a2.size() == 5 && a2[0].first == 0);
    }
    HnswGraph hg;  // hnsw.rs:571-640
    EXPECT(hg.is_empty() && hg.entry_point() == ISL_NO_ENTRY);
    EXPECT(hg.insert(q) == 0 && hg.len() == 1 && hg.dimension() == d);
    std::vector<float> rest(v.begin() + d, v.end());
    EXPECT(hg.insert_batch(rest, d, nullptr, 7, 16) == 1 && hg.len() == n);
    auto hr = hg.search(q, 5, 50);
    EXPECT(hr.size() == 5 && hr[0].first == 0 && hr[0].second < 0.01f);
    EXPECT(!hg.neighbors_at(0, 0).empty());
    EncoderConfig ec;
    ec.vocab_size = 200; ec.hidden_size = 64; ec.num_layers = 1; ec.num_heads = 1; ec.intermediate_size = 128; ec.max_position = 16;
    Encoder enc(ec);
    enc.init_random(3, 0.05f);
    std::vector<int32_t> toks = {5, 6, 7, 0, 9, 10, 11, 12};
    auto emb = enc.embed(toks, {3, 4}, 4);
    EXPECT(emb.size() == 2 * 64);
    float nrm = 0.f;
    for (int i = 0; i < 64; ++i) nrm += emb[i] * emb[i];
    EXPECT(std::fabs(nrm - 1.0f) < 1e-3f);  // L2-normalised (candle_provider.rs:477-494)
    try {
      hg.insert(std::vector<float>(d + 1, 0.f));
      EXPECT(false);
    } catch (const CoreError& e) {
      EXPECT(e.kind == ErrorKind::DimensionMismatch);
    }
  } else {
    try {  // no device: loud failure, never a CPU fallback
      calculate(DistanceMetric::Euclidean, {0.f, 0.f}, {3.f, 4.f});
      if (isl_device_count() == 0) EXPECT(false);
    } catch (const CoreError& e) {
      EXPECT(e.kind == ErrorKind::Cuda);
    }
  }
  std::printf("OK\n");
  return 0;
}
