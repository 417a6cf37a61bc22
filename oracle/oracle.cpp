// oracle.cpp — CPU restatement of the reference's LEANN / HNSW / PQ hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Every function cites the reference Rust it
// restates (paths relative to the reference root).  Arithmetic rules taken from the Rust:
// f32 everywhere, `x * y` and `+` rounded separately (rustc never contracts to FMA),
// `.sum()` is a left fold from 0.0, `.powi(2)` is `x * x`, sorts are stable.
// Build with -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// distance.rs:71-122
// ---------------------------------------------------------------------------------------
inline float cosine_distance(const float* a, const float* b, uint64_t d) {
  float dot = 0.0f, na = 0.0f, nb = 0.0f;  // distance.rs:72-74
  for (uint64_t i = 0; i < d; ++i) {       // distance.rs:76-80, one interleaved loop
    dot = dot + a[i] * b[i];
    na = na + a[i] * a[i];
    nb = nb + b[i] * b[i];
  }
  float norm = std::sqrt(na * nb);         // distance.rs:82
  if (norm == 0.0f) return 1.0f;           // distance.rs:83-85
  return 1.0f - (dot / norm);              // distance.rs:87
}

inline float l2_squared(const float* a, const float* b, uint64_t d) {
  float s = 0.0f;                          // distance.rs:98-107
  for (uint64_t i = 0; i < d; ++i) {
    float diff = a[i] - b[i];
    s = s + diff * diff;
  }
  return s;
}

inline float dot_distance(const float* a, const float* b, uint64_t d) {
  float s = 0.0f;                          // distance.rs:112-115
  for (uint64_t i = 0; i < d; ++i) s = s + a[i] * b[i];
  return -s;
}

inline float manhattan_distance(const float* a, const float* b, uint64_t d) {
  float s = 0.0f;                          // distance.rs:119-122
  for (uint64_t i = 0; i < d; ++i) s = s + std::fabs(a[i] - b[i]);
  return s;
}

inline float calc(int32_t metric, const float* a, const float* b, uint64_t d) {
  switch (metric) {                        // distance.rs:46-51
    case ISL_METRIC_COSINE: return cosine_distance(a, b, d);
    case ISL_METRIC_EUCLIDEAN: return std::sqrt(l2_squared(a, b, d));  // distance.rs:92-94
    case ISL_METRIC_DOT: return dot_distance(a, b, d);
    default: return manhattan_distance(a, b, d);
  }
}

// ---------------------------------------------------------------------------------------
// Ordering of (OrderedFloat<f32>, u64) tuples (leann.rs:701-702, 907-908; ordered-float
// 5.1.0: NaN is greater than every number and equal to itself; -0.0 == +0.0).
// ---------------------------------------------------------------------------------------
struct Key {
  float d;
  uint64_t id;
};
inline bool of_lt(float a, float b) {  // OrderedFloat a < b
  bool an = std::isnan(a), bn = std::isnan(b);
  if (an || bn) return !an && bn;
  return a < b;
}
inline bool of_gt(float a, float b) { return of_lt(b, a); }
inline bool key_lt(const Key& x, const Key& y) {
  if (of_lt(x.d, y.d)) return true;
  if (of_lt(y.d, x.d)) return false;
  return x.id < y.id;
}
struct MaxFirst {  // std::priority_queue top() = greatest key   (results: BinaryHeap<(dist,id)>)
  bool operator()(const Key& x, const Key& y) const { return key_lt(x, y); }
};
struct MinFirst {  // top() = smallest key                        (candidates: BinaryHeap<Reverse<..>>)
  bool operator()(const Key& x, const Key& y) const { return key_lt(y, x); }
};
using MaxHeap = std::priority_queue<Key, std::vector<Key>, MaxFirst>;
using MinHeap = std::priority_queue<Key, std::vector<Key>, MinFirst>;

// Exact visited set over dense ids (HashSet<u64> in the reference; ids are < n here).
struct Visited {
  std::vector<uint64_t> bits;
  std::vector<uint64_t> touched;
  void reset(uint64_t n) {
    if (bits.size() < (n + 63) / 64) {
      bits.assign((n + 63) / 64, 0);
      touched.clear();
      return;
    }
    for (uint64_t w : touched) bits[w] = 0;
    touched.clear();
  }
  bool insert(uint64_t id) {  // true when newly inserted (HashSet::insert)
    uint64_t w = id >> 6, m = 1ull << (id & 63);
    if (bits[w] & m) return false;
    if (bits[w] == 0) touched.push_back(w);
    bits[w] |= m;
    return true;
  }
};

// The seeded stand-in for thread_rng in PruningStrategy::Proportional (include/islands_b200.h, isl_pruning_strategy):
// draw c of query q = 24 random bits of a splitmix64-mixed counter, as a uniform f32 in [0, 1).
inline uint64_t splitmix_mix(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
inline float prune_draw(uint64_t seed, uint64_t query, uint64_t draw) {
  uint64_t h = splitmix_mix(splitmix_mix(seed + 0x9E3779B97F4A7C15ull * (query + 1)) + 0x9E3779B97F4A7C15ull * (draw + 1));
  return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

// leann.rs:1017-1053: Proportional.  Rewrites `cands` to the selected ids, in order.
template <class G>
void prune_proportional(const isl_leann_config& cfg, const G& g, std::vector<uint64_t>& cands, uint64_t query,
                        uint64_t* draw_ctr) {
  float fn = (float)cands.size();
  uint64_t num_to_keep = (uint64_t)std::ceil(fn * (1.0f - cfg.prune_ratio));  // :1001-1002
  if (num_to_keep < 1) num_to_keep = 1;                                       // :1003
  auto degree = [&](uint64_t id) -> uint64_t {  // graph.degree_counts.get(id)
    const uint64_t* nb;
    uint64_t cnt;
    g.get(id, nb, cnt);
    return cnt;
  };
  uint64_t total = 0;
  for (uint64_t id : cands) total += degree(id);                              // :1019-1028
  if (total == 0) {                                                           // :1029-1031
    cands.resize(std::min<uint64_t>(num_to_keep, cands.size()));
    return;
  }
  std::vector<uint64_t> selected;
  for (uint64_t id : cands) {                                                 // :1034-1048
    float prob = (float)degree(id) / (float)total;
    float u = prune_draw(cfg.prune_seed, query, (*draw_ctr)++);               // rand::thread_rng().gen::<f32>() in the reference
    if (u < prob * (float)num_to_keep) {
      selected.push_back(id);
      if (selected.size() >= num_to_keep) break;
    }
  }
  if (selected.empty()) selected.push_back(cands[0]);                         // :1049-1051
  cands.swap(selected);
}

// leann.rs:991-1016 (Global / Local).  `cands` is the unvisited list; returns how many of its prefix to keep.
inline uint64_t prune_keep(const isl_leann_config& cfg, uint64_t n_cands, uint64_t results_len,
                           uint64_t ef) {
  if (cfg.prune_ratio == 0.0f || n_cands == 0) return n_cands;  // leann.rs:997-999
  float fn = (float)n_cands;
  uint64_t num_to_keep = (uint64_t)std::ceil(fn * (1.0f - cfg.prune_ratio));  // :1001-1002
  if (num_to_keep < 1) num_to_keep = 1;                                       // :1003
  if (cfg.pruning_strategy == ISL_PRUNE_GLOBAL) {                             // :1006-1012
    float ratio = (float)results_len / (float)ef;
    uint64_t adjusted = (uint64_t)std::ceil(fn * (1.0f - ratio * cfg.prune_ratio));
    if (adjusted < 1) adjusted = 1;
    return std::min<uint64_t>(adjusted, n_cands);
  }
  return std::min<uint64_t>(num_to_keep, n_cands);  // Local, :1013-1016
}

template <class F>
void parallel_for(uint64_t n, int threads, F fn) {
  if (threads <= 1 || n <= 1) {
    fn(0, n, 0);
    return;
  }
  uint64_t t = std::min<uint64_t>((uint64_t)threads, n);
  std::vector<std::thread> pool;
  // Dynamic chunks of 1 via an atomic counter keep long queries from serialising a static split.
  auto counter = std::make_shared<std::atomic<uint64_t>>(0);
  for (uint64_t w = 0; w < t; ++w) {
    pool.emplace_back([=]() {
      for (;;) {
        uint64_t i = counter->fetch_add(1);
        if (i >= n) break;
        fn(i, i + 1, (int)w);
      }
    });
  }
  for (auto& th : pool) th.join();
}

// Adjacency accessor shared by the CSR search and the build-time search.
struct CsrView {
  const uint64_t* offsets;
  const uint64_t* nbrs;
  uint64_t n;
  inline void get(uint64_t id, const uint64_t*& p, uint64_t& cnt) const {
    if (id >= n) {  // CsrGraph::get_neighbors -> None (leann.rs:227-229)
      p = nullptr;
      cnt = 0;
      return;
    }
    p = nbrs + offsets[id];
    cnt = offsets[id + 1] - offsets[id];
  }
};
struct AdjView {
  const std::vector<std::vector<uint64_t>>* adj;
  inline void get(uint64_t id, const uint64_t*& p, uint64_t& cnt) const {
    const auto& v = (*adj)[id];
    p = v.data();
    cnt = v.size();
  }
};

// Best-first search shared by leann.rs:692-749 (build, no pruning) and leann.rs:899-988
// (search, pruning strategy applied).  Returns the whole result list sorted by (dist,id).
template <class G>
void best_first(const isl_leann_config& cfg, bool apply_pruning, const G& g, const float* vectors,
                uint32_t d, const float* q, uint64_t entry, uint64_t ef, Visited& visited,
                std::vector<Key>& out, isl_search_stats* st, uint64_t query_index = 0) {
  const int32_t metric = cfg.metric;
  uint64_t draw_ctr = 0;
  MinHeap candidates;
  MaxHeap results;
  float entry_dist = calc(metric, q, vectors + entry * (uint64_t)d, d);  // leann.rs:911-912
  visited.insert(entry);                                                 // :914
  candidates.push({entry_dist, entry});                                  // :915
  results.push({entry_dist, entry});                                     // :916
  uint64_t n_hop = 0, n_edge = 0, n_dist = 1;
  std::vector<uint64_t> unvisited;
  while (!candidates.empty()) {                                          // :922
    Key c = candidates.top();
    candidates.pop();
    if (results.size() >= ef && of_gt(c.d, results.top().d)) break;      // :924-928
    const uint64_t* nb;
    uint64_t cnt;
    g.get(c.id, nb, cnt);                                                // :931
    n_hop++;
    n_edge += cnt;
    unvisited.clear();
    for (uint64_t i = 0; i < cnt; ++i)                                   // :933-937
      if (visited.insert(nb[i])) unvisited.push_back(nb[i]);
    if (unvisited.empty()) continue;                                     // :939-941
    uint64_t keep;
    if (apply_pruning && cfg.pruning_strategy == ISL_PRUNE_PROPORTIONAL && cfg.prune_ratio != 0.0f) {
      prune_proportional(cfg, g, unvisited, query_index, &draw_ctr);     // :1017-1053
      keep = unvisited.size();
    } else {
      keep = apply_pruning ? prune_keep(cfg, unvisited.size(), results.size(), ef) : unvisited.size();  // :944
    }
    n_dist += keep;                                                      // :950
    for (uint64_t i = 0; i < keep; ++i) {                                // :953-970
      uint64_t nbid = unvisited[i];
      float nd = calc(metric, q, vectors + nbid * (uint64_t)d, d);
      bool should_add = results.size() < ef || nd < results.top().d;     // raw f32 `<` (:956-960)
      if (should_add) {
        candidates.push({nd, nbid});
        results.push({nd, nbid});
        if (results.size() > ef) results.pop();
      }
    }
  }
  out.clear();
  out.reserve(results.size());
  while (!results.empty()) {
    out.push_back(results.top());
    results.pop();
  }
  // leann.rs:984-987 sorts by distance only (stable over heap order, i.e. unspecified for
  // ties); the fixed rule is (dist, id).
  std::sort(out.begin(), out.end(), key_lt);
  if (st) {
    st->n_hop = n_hop;
    st->n_edge = n_edge;
    st->n_dist = n_dist;
    st->n_adc = 0;
    st->n_rerank = 0;
  }
}

inline void write_topk(const std::vector<Key>& res, uint32_t k, uint64_t* ids, float* dist,
                       uint32_t* count) {
  uint32_t c = (uint32_t)std::min<uint64_t>(k, res.size());  // take(k), leann.rs:895
  for (uint32_t i = 0; i < k; ++i) {
    ids[i] = i < c ? res[i].id : ISL_INVALID_ID;
    dist[i] = i < c ? res[i].d : std::numeric_limits<float>::infinity();
  }
  if (count) *count = c;
}

// leann.rs:761-833
std::vector<Key> hub_preserving_select(const isl_leann_config& cfg, const std::vector<Key>& cands,
                                       const std::vector<uint64_t>& degree_of_cand,
                                       uint64_t max_conn) {
  if (cands.size() <= max_conn) return cands;  // :767-769
  std::vector<uint64_t> degrees = degree_of_cand;
  std::sort(degrees.begin(), degrees.end(), std::greater<uint64_t>());  // :778
  uint64_t hub_count = (uint64_t)std::ceil((float)degrees.size() * cfg.hub_percentile);  // :780
  const uint64_t NONE = std::numeric_limits<uint64_t>::max();
  uint64_t thr = (hub_count > 0 && hub_count < degrees.size()) ? degrees[hub_count - 1] : NONE;
  struct Hub {
    Key k;
    uint64_t deg;
  };
  std::vector<Hub> hubs;
  std::vector<Key> regular;
  for (size_t i = 0; i < cands.size(); ++i) {  // :790-797
    uint64_t deg = degree_of_cand[i];
    if (deg >= thr && thr < NONE)
      hubs.push_back({cands[i], deg});
    else
      regular.push_back(cands[i]);
  }
  std::stable_sort(hubs.begin(), hubs.end(),
                   [](const Hub& a, const Hub& b) { return a.deg > b.deg; });  // :800
  std::stable_sort(regular.begin(), regular.end(), [](const Key& a, const Key& b) {
    return a.d < b.d;  // partial_cmp, Equal when unordered (:802)
  });
  std::vector<Key> sel;
  sel.reserve(max_conn);
  uint64_t hub_slots = std::max<uint64_t>(max_conn / 4, 1);  // :807
  for (size_t i = 0; i < hubs.size() && i < hub_slots; ++i) sel.push_back(hubs[i].k);  // :808-810
  auto has = [&](uint64_t id) {
    for (auto& s : sel)
      if (s.id == id) return true;
    return false;
  };
  for (auto& r : regular) {  // :813-820
    if (sel.size() >= max_conn) break;
    if (!has(r.id)) sel.push_back(r);
  }
  for (size_t i = hub_slots; i < hubs.size(); ++i) {  // :823-830
    if (sel.size() >= max_conn) break;
    if (!has(hubs[i].k.id)) sel.push_back(hubs[i].k);
  }
  return sel;
}

// leann.rs:634-658
std::vector<uint64_t> prune_neighbors(int32_t metric, const float* vectors, uint32_t d,
                                      uint64_t node, const std::vector<uint64_t>& nbrs,
                                      uint64_t max_conn) {
  struct S {
    uint64_t id;
    float dist;
  };
  std::vector<S> scored;
  scored.reserve(nbrs.size());
  const float* nv = vectors + node * (uint64_t)d;
  for (uint64_t id : nbrs) scored.push_back({id, calc(metric, nv, vectors + id * (uint64_t)d, d)});
  std::stable_sort(scored.begin(), scored.end(),
                   [](const S& a, const S& b) { return a.dist < b.dist; });  // :652
  std::vector<uint64_t> out;
  for (size_t i = 0; i < scored.size() && i < max_conn; ++i) out.push_back(scored[i].id);
  return out;
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// StdRng of rand 0.8.5 = ChaCha12Rng (rand_chacha 0.3.1) seeded by SeedableRng::seed_from_u64 (rand_core 0.6.4):
// third-party crates, restated from their published algorithms (pq.rs:190-193 seeds it, :380 / :404 / :454 draw
// from it).  Written as one continuous stream of 32-bit words — word i = ChaCha12 block i / 16, word i % 16 —
// which is what BlockRng's 64-word buffer delivers: next_u32 = the next word, next_u64 = the next two words, low
// first (also across a refill).  The block function is pinned by RFC 7539 and the ChaCha12 / ChaCha8
// known answers of draft-strombergson-chacha-test-vectors (tests/test_std_rng.py).
static void chacha_words(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
  const uint32_t c[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  uint32_t x[16], in[16];
  for (int i = 0; i < 4; ++i) in[i] = c[i];
  for (int i = 0; i < 8; ++i) in[4 + i] = key[i];
  in[12] = (uint32_t)counter;
  in[13] = (uint32_t)(counter >> 32);
  in[14] = (uint32_t)stream;
  in[15] = (uint32_t)(stream >> 32);
  std::memcpy(x, in, sizeof(x));
#define ORC_ROTL(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define ORC_QR(a, b, c_, d)                                   \
  x[a] += x[b]; x[d] ^= x[a]; x[d] = ORC_ROTL(x[d], 16);      \
  x[c_] += x[d]; x[b] ^= x[c_]; x[b] = ORC_ROTL(x[b], 12);    \
  x[a] += x[b]; x[d] ^= x[a]; x[d] = ORC_ROTL(x[d], 8);       \
  x[c_] += x[d]; x[b] ^= x[c_]; x[b] = ORC_ROTL(x[b], 7);
  for (int r = 0; r < rounds / 2; ++r) {
    ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
    ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
  }
#undef ORC_QR
#undef ORC_ROTL
  for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

struct Rng {
  uint32_t key[8];
  uint64_t pos = 0;          // index of the next 32-bit word of the stream
  uint64_t have = ~0ull;     // block held in `blk`
  uint32_t blk[16];
  explicit Rng(uint64_t seed) {  // seed_from_u64: eight PCG32 (XSH-RR) outputs, little-endian, form the 32-byte seed
    uint64_t st = seed;
    for (int i = 0; i < 8; ++i) {
      st = st * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27), rot = (uint32_t)(st >> 59);
      key[i] = rot ? ((xs >> rot) | (xs << (32 - rot))) : xs;
    }
  }
  uint32_t word(uint64_t i) {
    if (i / 16 != have) {
      have = i / 16;
      chacha_words(key, have, 0, 12, blk);
    }
    return blk[i % 16];
  }
  uint32_t next_u32() { return word(pos++); }
  uint64_t next_u64() {
    uint64_t lo = word(pos), hi = word(pos + 1);
    pos += 2;
    return (hi << 32) | lo;
  }
  float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }  // Standard: 24 bits, [0,1)
  uint64_t choose(uint64_t len) {  // SliceRandom::choose -> gen_index -> gen_range(0..len) (u32 when it fits)
    if (len <= 0xffffffffull) {
      uint32_t n = (uint32_t)len, lz = 0;
      while (!((n << lz) & 0x80000000u)) ++lz;
      uint32_t zone = (n << lz) - 1u;
      for (;;) {
        uint64_t m = (uint64_t)next_u32() * n;
        if ((uint32_t)m <= zone) return m >> 32;
      }
    }
    uint32_t lz = 0;
    while (!((len << lz) & 0x8000000000000000ull)) ++lz;
    uint64_t zone = (len << lz) - 1ull;
    for (;;) {
      unsigned __int128 m = (unsigned __int128)next_u64() * len;
      if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
    }
  }
};

// pq.rs:362-463
int32_t kmeans(const std::vector<const float*>& vecs, uint32_t dim, uint32_t k_in,
               uint32_t iterations, int32_t metric, Rng& rng, std::vector<float>& centroids,
               uint32_t& k_out) {
  if (vecs.empty()) return ISL_EMPTY_COLLECTION;  // :369-371
  uint64_t n = vecs.size();
  uint32_t k = (uint32_t)std::min<uint64_t>(k_in, n);  // :374
  centroids.clear();
  auto push_centroid = [&](const float* v) { centroids.insert(centroids.end(), v, v + dim); };
  push_centroid(vecs[rng.next_u64() % n]);  // :380-381
  std::vector<float> distances(n);
  while (centroids.size() / dim < k) {  // :384-415
    uint64_t nc = centroids.size() / dim;
    for (uint64_t i = 0; i < n; ++i) {
      float best = std::numeric_limits<float>::max();
      for (uint64_t c = 0; c < nc; ++c)
        best = std::fmin(best, calc(metric, vecs[i], centroids.data() + c * dim, dim));
      distances[i] = best;
    }
    float total = 0.0f;
    for (uint64_t i = 0; i < n; ++i) total = total + distances[i];  // :396
    if (total > 0.0f)
      for (uint64_t i = 0; i < n; ++i) distances[i] = distances[i] / total;
    float threshold = rng.next_f32();  // :404
    float cumsum = 0.0f;
    uint64_t selected = 0;
    for (uint64_t i = 0; i < n; ++i) {  // :407-413
      cumsum = cumsum + distances[i];
      if (cumsum >= threshold) {
        selected = i;
        break;
      }
    }
    push_centroid(vecs[selected]);
  }
  std::vector<uint64_t> assign(n, 0);
  for (uint32_t it = 0; it < iterations; ++it) {  // :420-460
    for (uint64_t i = 0; i < n; ++i) {
      float best = std::numeric_limits<float>::max();
      uint64_t bc = 0;
      for (uint32_t c = 0; c < k; ++c) {
        float dist = calc(metric, vecs[i], centroids.data() + (uint64_t)c * dim, dim);
        if (dist < best) {
          best = dist;
          bc = c;
        }
      }
      assign[i] = bc;
    }
    std::vector<float> nc((uint64_t)k * dim, 0.0f);
    std::vector<uint64_t> counts(k, 0);
    for (uint64_t i = 0; i < n; ++i) {  // :439-445
      uint64_t c = assign[i];
      counts[c]++;
      float* dst = nc.data() + c * dim;
      for (uint32_t j = 0; j < dim; ++j) dst[j] = dst[j] + vecs[i][j];
    }
    for (uint32_t c = 0; c < k; ++c) {  // :447-457
      float* dst = nc.data() + (uint64_t)c * dim;
      if (counts[c] > 0) {
        float fc = (float)counts[c];
        for (uint32_t j = 0; j < dim; ++j) dst[j] = dst[j] / fc;
      } else {
        const float* rv = vecs[rng.choose(n)];  // vectors.choose(rng) (pq.rs:454)
        std::memcpy(dst, rv, sizeof(float) * dim);
      }
    }
    centroids.swap(nc);
  }
  k_out = k;
  return ISL_OK;
}

}  // namespace

// =========================================================================================
// C API
// =========================================================================================
extern "C" {

float orc_distance(int32_t metric, const float* a, const float* b, uint64_t d) {
  return calc(metric, a, b, d);
}

float orc_distance_squared(int32_t metric, const float* a, const float* b, uint64_t d) {
  if (metric == ISL_METRIC_EUCLIDEAN) return l2_squared(a, b, d);  // distance.rs:63
  float x = calc(metric, a, b, d);                                 // distance.rs:64
  return x * x;
}

void orc_distance_batch(int32_t metric, const float* q, const float* rows, uint64_t n, uint32_t d,
                        float* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = calc(metric, q, rows + i * (uint64_t)d, d);
}

void orc_normalize(float* v, uint64_t d) {
  float s = 0.0f;  // distance.rs:126
  for (uint64_t i = 0; i < d; ++i) s = s + v[i] * v[i];
  float norm = std::sqrt(s);
  if (norm > 0.0f)
    for (uint64_t i = 0; i < d; ++i) v[i] = v[i] / norm;
}

uint64_t orc_level_from_uniform(double u, double ml, uint64_t max_layers) {
  double lv = std::floor(-std::log(u) * ml);  // leann.rs:552, hnsw.rs:209
  uint64_t level;
  if (!(lv >= 0.0))
    level = 0;  // `as usize` saturates: negative / NaN -> 0
  else if (lv >= 1.8446744073709552e19)
    level = std::numeric_limits<uint64_t>::max();
  else
    level = (uint64_t)lv;
  return std::min<uint64_t>(level, max_layers - 1);  // :553
}

void orc_draw_levels(uint64_t seed, uint64_t n, double ml, uint64_t max_layers, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t r = splitmix64(seed + i) >> 11;
    if (r == 0) r = 1;
    double u = (double)r * (1.0 / 9007199254740992.0);
    out[i] = orc_level_from_uniform(u, ml, max_layers);
  }
}

int32_t orc_leann_search(const isl_leann_config* cfg, const float* vectors, uint64_t n, uint32_t d,
                         const uint64_t* offsets, const uint64_t* nbrs, int64_t entry,
                         const float* queries, uint64_t nq, uint32_t k, uint32_t ef_in,
                         uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                         isl_search_stats* stats, int32_t threads) {
  if (n == 0) {  // leann.rs:875-877
    for (uint64_t qi = 0; qi < nq; ++qi) {
      std::vector<Key> none;
      write_topk(none, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
      if (stats) stats[qi] = isl_search_stats{0, 0, 0, 0, 0};
    }
    return ISL_OK;
  }
  if (entry < 0) return ISL_INDEX_NOT_BUILT;  // leann.rs:889
  uint64_t ef = std::max<uint64_t>(ef_in, k);  // leann.rs:890
  CsrView g{offsets, nbrs, n};
  int nt = std::max(1, threads);
  std::vector<Visited> vis(nt);
  parallel_for(nq, threads, [&](uint64_t b, uint64_t e, int w) {
    std::vector<Key> res;
    for (uint64_t qi = b; qi < e; ++qi) {
      vis[w].reset(n);
      best_first(*cfg, true, g, vectors, d, queries + qi * (uint64_t)d, (uint64_t)entry, ef, vis[w],
                 res, stats ? stats + qi : nullptr, qi);
      write_topk(res, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
    }
  });
  return ISL_OK;
}

int32_t orc_leann_build_batched(const isl_leann_config* cfg, const float* vectors, uint64_t n,
                                uint32_t d, const uint64_t* levels, uint32_t batch,
                                uint64_t* out_offsets, uint64_t* out_nbrs, uint64_t* out_num_edges,
                                int64_t* out_entry, uint64_t* out_max_level, int32_t threads) {
  // batch == 1 is exactly leann.rs:560-631.  batch > 1 is the snapshot-round model of
  // nodes [s, e) all search the graph as it was after node s-1, then their
  // edges are applied in id order (which is what the sequential loop does for one node).
  out_offsets[0] = 0;
  *out_num_edges = 0;
  *out_entry = ISL_NO_ENTRY;
  *out_max_level = 0;
  if (n == 0) return ISL_OK;  // leann.rs:565-567
  if (batch == 0) batch = 1;
  const uint64_t m0 = cfg->m0;
  std::vector<std::vector<uint64_t>> adj;
  adj.reserve(n);
  int64_t entry = ISL_NO_ENTRY;
  uint64_t max_level = 0;
  int nt = std::max(1, threads);
  std::vector<Visited> vis(nt);
  AdjView g{&adj};
  uint64_t s = 0;
  while (s < n) {
    uint64_t round = std::min<uint64_t>(batch, std::max<uint64_t>(1, s / 2));  // ramp 1,1,1,1,2,3,4,6,...
    uint64_t e = std::min<uint64_t>(n, s + round);
    std::vector<std::vector<uint64_t>> fwd(e - s);
    if (s > 0) {
      uint64_t entry_id = entry >= 0 ? (uint64_t)entry : 0;  // leann.rs:669
      parallel_for(e - s, threads, [&](uint64_t b, uint64_t en, int w) {
        std::vector<Key> cands;
        for (uint64_t j = b; j < en; ++j) {
          uint64_t id = s + j;
          vis[w].reset(s);
          best_first(*cfg, false, g, vectors, d, vectors + id * (uint64_t)d, entry_id,
                     cfg->ef_construction, vis[w], cands, nullptr);  // leann.rs:672-678
          if (cfg->high_degree_pruning) {                            // :681-683 (adjacency non-empty)
            std::vector<uint64_t> degs(cands.size());
            for (size_t i = 0; i < cands.size(); ++i) degs[i] = adj[cands[i].id].size();
            cands = hub_preserving_select(*cfg, cands, degs, m0);
          } else if (cands.size() > m0) {
            cands.resize(m0);                                        // :685
          }
          fwd[j].reserve(cands.size());
          for (auto& c : cands) fwd[j].push_back(c.id);              // :688
        }
      });
    }
    for (uint64_t id = s; id < e; ++id) {
      const auto& nb = fwd[id - s];
      adj.push_back(nb);                                             // leann.rs:592
      for (uint64_t u : nb) {                                        // :593-607
        auto& lu = adj[u];
        if (std::find(lu.begin(), lu.end(), id) == lu.end()) {
          lu.push_back(id);
          if (lu.size() > m0) lu = prune_neighbors(cfg->metric, vectors, d, u, lu, m0);
        }
      }
      uint64_t level = levels ? levels[id] : 0;
      if (entry < 0 || level > max_level) {                          // :610-613
        entry = (int64_t)id;
        max_level = level;
      }
    }
    s = e;
  }
  uint64_t pos = 0;  // leann.rs:618-627
  for (uint64_t i = 0; i < n; ++i) {
    for (uint64_t x : adj[i]) out_nbrs[pos++] = x;
    out_offsets[i + 1] = pos;
  }
  *out_num_edges = pos;
  *out_entry = entry;
  *out_max_level = max_level;
  return ISL_OK;
}

int32_t orc_leann_build(const isl_leann_config* cfg, const float* vectors, uint64_t n, uint32_t d,
                        const uint64_t* levels, uint64_t* out_offsets, uint64_t* out_nbrs,
                        uint64_t* out_num_edges, int64_t* out_entry, uint64_t* out_max_level) {
  return orc_leann_build_batched(cfg, vectors, n, d, levels, 1, out_offsets, out_nbrs,
                                 out_num_edges, out_entry, out_max_level, 1);
}

// ---- PQ ----------------------------------------------------------------------------------
void orc_pq_encode(int32_t metric, const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                   const float* vectors, uint64_t n, uint16_t* out_codes) {
  uint64_t d = (uint64_t)m * dsub;
  for (uint64_t i = 0; i < n; ++i) {
    for (uint32_t j = 0; j < m; ++j) {  // pq.rs:234-241
      const float* sub = vectors + i * d + (uint64_t)j * dsub;
      const float* cb = codebooks + (uint64_t)j * ksub * dsub;
      uint32_t best = 0;  // pq.rs:94-103
      float bd = std::numeric_limits<float>::max();
      for (uint32_t c = 0; c < ksub; ++c) {
        float dist = calc(metric, sub, cb + (uint64_t)c * dsub, dsub);
        if (dist < bd) {
          bd = dist;
          best = c;
        }
      }
      out_codes[i * m + j] = (uint16_t)best;  // `code as u16`
    }
  }
}

int32_t orc_pq_decode(const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                      const uint16_t* codes, uint64_t n, float* out) {
  uint64_t d = (uint64_t)m * dsub;
  for (uint64_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < m; ++j) {  // pq.rs:261-268
      uint32_t c = codes[i * m + j];
      if (c >= ksub) return ISL_PQ_ERROR;
      std::memcpy(out + i * d + (uint64_t)j * dsub,
                  codebooks + ((uint64_t)j * ksub + c) * dsub, sizeof(float) * dsub);
    }
  return ISL_OK;
}

void orc_pq_build_tables(const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                         const float* query, float* out_tables) {
  for (uint32_t j = 0; j < m; ++j)  // pq.rs:317-335
    for (uint32_t c = 0; c < ksub; ++c) {
      const float* qs = query + (uint64_t)j * dsub;
      const float* ce = codebooks + ((uint64_t)j * ksub + c) * dsub;
      float s = 0.0f;
      for (uint32_t t = 0; t < dsub; ++t) {
        float diff = qs[t] - ce[t];
        s = s + diff * diff;  // (a - b).powi(2) summed
      }
      out_tables[(uint64_t)j * ksub + c] = s;
    }
}

void orc_pq_table_distance(const float* tables, uint32_t m, uint32_t ksub, const uint16_t* codes,
                           uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {  // pq.rs:341-348
    float s = 0.0f;
    for (uint32_t j = 0; j < m; ++j) s = s + tables[(uint64_t)j * ksub + codes[i * m + j]];
    out[i] = std::sqrt(s);
  }
}

void orc_pq_asymmetric_distance(const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                                const float* query, const uint16_t* codes, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {  // pq.rs:283-303
    float total = 0.0f;
    for (uint32_t j = 0; j < m; ++j) {
      const float* qs = query + (uint64_t)j * dsub;
      const float* ce = codebooks + ((uint64_t)j * ksub + codes[i * m + j]) * dsub;
      float sub = 0.0f;
      for (uint32_t t = 0; t < dsub; ++t) {
        float diff = qs[t] - ce[t];
        sub = sub + diff * diff;
      }
      total = total + sub;
    }
    out[i] = std::sqrt(total);
  }
}

int32_t orc_pq_train(int32_t metric, const float* vectors, uint64_t n, uint32_t d, uint32_t m,
                     uint32_t ksub, uint32_t iterations, uint64_t seed, float* out_codebooks,
                     uint32_t* out_ksub) {
  if (n == 0) return ISL_EMPTY_COLLECTION;  // pq.rs:176-178
  uint32_t dsub = d / m;
  Rng rng(seed);  // one generator shared across subspaces in order (pq.rs:190-214)
  uint32_t k_eff = (uint32_t)std::min<uint64_t>(ksub, n);
  for (uint32_t j = 0; j < m; ++j) {
    std::vector<const float*> subs(n);
    for (uint64_t i = 0; i < n; ++i) subs[i] = vectors + i * (uint64_t)d + (uint64_t)j * dsub;
    std::vector<float> cent;
    uint32_t k_out = 0;
    int32_t st = kmeans(subs, dsub, ksub, iterations, metric, rng, cent, k_out);
    if (st != ISL_OK) return st;
    std::memcpy(out_codebooks + (uint64_t)j * k_eff * dsub, cent.data(),
                sizeof(float) * (uint64_t)k_out * dsub);
  }
  *out_ksub = k_eff;
  return ISL_OK;
}

// ---- two-level search (definition: DESIGN.md "two-level search"; pseudocode
// docs/leann-specification.md:223-269; PQ primitives pq.rs:307-348) ------------------------
int32_t orc_leann_search_two_level(const isl_leann_config* cfg, const float* vectors, uint64_t n,
                                   uint32_t d, const uint64_t* offsets, const uint64_t* nbrs,
                                   int64_t entry, const float* codebooks, uint32_t m, uint32_t ksub,
                                   const uint16_t* codes, const float* queries, uint64_t nq,
                                   uint32_t k, uint32_t ef_in, float rerank_ratio,
                                   uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                   isl_search_stats* stats, int32_t threads) {
  if (n == 0) {
    for (uint64_t qi = 0; qi < nq; ++qi) {
      std::vector<Key> none;
      write_topk(none, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
      if (stats) stats[qi] = isl_search_stats{0, 0, 0, 0, 0};
    }
    return ISL_OK;
  }
  if (entry < 0) return ISL_INDEX_NOT_BUILT;
  if (!(rerank_ratio > 0.0f) || rerank_ratio > 1.0f) return ISL_INVALID_ARGUMENT;
  const uint64_t ef = std::max<uint64_t>(ef_in, k);
  const uint32_t dsub = d / m;
  const int32_t metric = cfg->metric;
  CsrView g{offsets, nbrs, n};
  int nt = std::max(1, threads);
  std::vector<Visited> vis(nt);
  parallel_for(nq, threads, [&](uint64_t b, uint64_t e, int w) {
    std::vector<float> lut((uint64_t)m * ksub);
    std::vector<Key> res;
    for (uint64_t qi = b; qi < e; ++qi) {
      const float* q = queries + qi * (uint64_t)d;
      orc_pq_build_tables(codebooks, m, ksub, dsub, q, lut.data());
      Visited& visited = vis[w];
      visited.reset(n);
      MinHeap eq, aq;  // exact queue, approximate queue
      MaxHeap results;
      float ed = calc(metric, q, vectors + (uint64_t)entry * d, d);
      visited.insert((uint64_t)entry);
      eq.push({ed, (uint64_t)entry});
      results.push({ed, (uint64_t)entry});
      uint64_t n_hop = 0, n_edge = 0, n_dist = 1, n_adc = 0, n_rerank = 0;
      while (!eq.empty()) {
        Key c = eq.top();
        eq.pop();
        if (results.size() >= ef && of_gt(c.d, results.top().d)) break;
        const uint64_t* nb;
        uint64_t cnt;
        g.get(c.id, nb, cnt);
        n_hop++;
        n_edge += cnt;
        for (uint64_t i = 0; i < cnt; ++i) {
          if (!visited.insert(nb[i])) continue;
          float s = 0.0f;  // table_distance, pq.rs:341-348
          const uint16_t* cd = codes + nb[i] * (uint64_t)m;
          for (uint32_t j = 0; j < m; ++j) s = s + lut[(uint64_t)j * ksub + cd[j]];
          aq.push({std::sqrt(s), nb[i]});
          n_adc++;
        }
        if (aq.empty()) continue;
        // promote ceil(a * |AQ|) best by (adc, id), at least one
        uint64_t promote = (uint64_t)std::ceil((float)aq.size() * rerank_ratio);
        if (promote < 1) promote = 1;
        if (promote > aq.size()) promote = aq.size();
        for (uint64_t p = 0; p < promote; ++p) {
          Key a = aq.top();
          aq.pop();
          float nd = calc(metric, q, vectors + a.id * (uint64_t)d, d);
          n_dist++;
          n_rerank++;
          bool should_add = results.size() < ef || nd < results.top().d;
          if (should_add) {
            eq.push({nd, a.id});
            results.push({nd, a.id});
            if (results.size() > ef) results.pop();
          }
        }
      }
      res.clear();
      while (!results.empty()) {
        res.push_back(results.top());
        results.pop();
      }
      std::sort(res.begin(), res.end(), key_lt);
      write_topk(res, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
      if (stats) stats[qi] = isl_search_stats{n_hop, n_edge, n_dist, n_adc, n_rerank};
    }
  });
  return ISL_OK;
}

// bfloat16 rounding of one table entry of the ADC traversal: round to nearest even on the bit pattern, NaN -> quiet NaN
// (restated here independently of the product's common.cuh).
static float bf16_round(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) {
    u = 0x7fc00000u;
  } else {
    const uint32_t lsb = (u >> 16) & 1u;
    u = (u + 0x7fffu + lsb) & 0xffff0000u;
  }
  std::memcpy(&x, &u, 4);
  return x;
}

void orc_adc_table_round(const float* in, uint64_t count, float* out) {
  for (uint64_t i = 0; i < count; ++i) out[i] = bf16_round(in[i]);
}

// PQ ADC traversal + exact rerank: the loop of leann.rs:899-988 with table_distance (pq.rs:341-348)
// as the distance, then exact distances for the ef survivors and a final (dist,id) sort.
int32_t orc_leann_search_adc_rerank(const isl_leann_config* cfg, const float* vectors, uint64_t n, uint32_t d,
                                    const uint64_t* offsets, const uint64_t* nbrs, int64_t entry,
                                    const float* codebooks, uint32_t m, uint32_t ksub, const uint16_t* codes,
                                    const float* queries, uint64_t nq, uint32_t k, uint32_t ef_in, uint64_t* out_ids,
                                    float* out_dist, uint32_t* out_count, isl_search_stats* stats, int32_t threads,
                                    uint32_t rerank_limit) {
  if (n == 0) {
    for (uint64_t qi = 0; qi < nq; ++qi) {
      std::vector<Key> none;
      write_topk(none, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
      if (stats) stats[qi] = isl_search_stats{0, 0, 0, 0, 0};
    }
    return ISL_OK;
  }
  if (entry < 0) return ISL_INDEX_NOT_BUILT;
  const uint64_t ef = std::max<uint64_t>(ef_in, k);
  const uint32_t dsub = d / m;
  CsrView g{offsets, nbrs, n};
  int nt = std::max(1, threads);
  std::vector<Visited> vis(nt);
  parallel_for(nq, threads, [&](uint64_t b, uint64_t e, int w) {
    std::vector<float> lut((uint64_t)m * ksub);
    std::vector<Key> res, surv;
    auto adc = [&](uint64_t id) {
      float s = 0.0f;
      const uint16_t* cd = codes + id * (uint64_t)m;
      for (uint32_t j = 0; j < m; ++j) s = s + lut[(uint64_t)j * ksub + cd[j]];
      return std::sqrt(s);
    };
    for (uint64_t qi = b; qi < e; ++qi) {
      const float* q = queries + qi * (uint64_t)d;
      orc_pq_build_tables(codebooks, m, ksub, dsub, q, lut.data());
      // the traversal folds bfloat16-rounded table entries (definition: include/islands_b200.h, isl_index_search_adc_rerank)
      for (float& t : lut) t = bf16_round(t);
      Visited& visited = vis[w];
      visited.reset(n);
      MinHeap cand;
      MaxHeap results;
      float ed = adc((uint64_t)entry);
      visited.insert((uint64_t)entry);
      cand.push({ed, (uint64_t)entry});
      results.push({ed, (uint64_t)entry});
      uint64_t n_hop = 0, n_edge = 0, n_adc = 1;
      while (!cand.empty()) {
        Key c = cand.top();
        cand.pop();
        if (results.size() >= ef && of_gt(c.d, results.top().d)) break;
        const uint64_t* nb;
        uint64_t cnt;
        g.get(c.id, nb, cnt);
        n_hop++;
        n_edge += cnt;
        for (uint64_t i = 0; i < cnt; ++i) {
          if (!visited.insert(nb[i])) continue;
          float nd = adc(nb[i]);
          n_adc++;
          bool should_add = results.size() < ef || nd < results.top().d;
          if (should_add) {
            cand.push({nd, nb[i]});
            results.push({nd, nb[i]});
            if (results.size() > ef) results.pop();
          }
        }
      }
      // the survivors in ascending (adc, id) order; with a rerank limit only the first max(limit, k) of them
      // get an exact distance (isl_index_set_rerank_limit)
      surv.clear();
      while (!results.empty()) {
        surv.push_back(results.top());
        results.pop();
      }
      std::sort(surv.begin(), surv.end(), key_lt);
      uint64_t keep = surv.size();
      if (rerank_limit) keep = std::min<uint64_t>(keep, std::max<uint64_t>(rerank_limit, k));
      res.clear();
      for (uint64_t i = 0; i < keep; ++i)
        res.push_back({calc(cfg->metric, q, vectors + surv[i].id * (uint64_t)d, d), surv[i].id});  // exact rerank
      std::sort(res.begin(), res.end(), key_lt);
      write_topk(res, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
      if (stats) stats[qi] = isl_search_stats{n_hop, n_edge, res.size(), n_adc, res.size()};
    }
  });
  return ISL_OK;
}

void orc_merge_topk(const uint64_t* ids, const float* dist, uint32_t parts, uint64_t nq, uint32_t k,
                    uint64_t* out_ids, float* out_dist, uint32_t* out_count) {
  std::vector<Key> all;
  for (uint64_t qi = 0; qi < nq; ++qi) {
    all.clear();
    for (uint32_t p = 0; p < parts; ++p)
      for (uint32_t i = 0; i < k; ++i) {
        uint64_t off = ((uint64_t)p * nq + qi) * k + i;
        if (ids[off] != ISL_INVALID_ID) all.push_back({dist[off], ids[off]});
      }
    std::sort(all.begin(), all.end(), key_lt);  // search.rs:231 under the (dist,id) rule
    write_topk(all, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
  }
}

// Test hooks for the generator restatement: one ChaCha block, and a scripted sequence of draws
// (kind 0 = next_u32, 1 = next_u64, 2 = f32 bits, 3 = choose(bound)).
void orc_chacha_block(const uint32_t* key8, uint64_t counter, uint64_t stream, int32_t rounds, uint32_t* out16) {
  chacha_words(key8, counter, stream, rounds, out16);
}
void orc_std_rng_draw(uint64_t seed, const uint8_t* kinds, uint64_t count, uint64_t bound, uint64_t* out) {
  Rng rng(seed);
  for (uint64_t i = 0; i < count; ++i) {
    if (kinds[i] == 0) out[i] = rng.next_u32();
    else if (kinds[i] == 1) out[i] = rng.next_u64();
    else if (kinds[i] == 2) { float f = rng.next_f32(); uint32_t b; std::memcpy(&b, &f, 4); out[i] = b; }
    else out[i] = rng.choose(bound);
  }
}

float orc_to_similarity(float score) { return 1.0f / (1.0f + score); }  // search.rs:100-102

// ---- HNSW (hnsw.rs) -----------------------------------------------------------------------
struct orc_hnsw {
  isl_hnsw_config cfg;
  uint32_t d;
  struct Node {
    std::vector<float> v;
    std::vector<std::vector<uint64_t>> conn;  // per layer 0..=level
    uint64_t level;
  };
  std::unordered_map<uint64_t, Node> nodes;
  int64_t entry = ISL_NO_ENTRY;
  uint64_t max_level = 0;
  uint64_t next_id = 0;
};

namespace {
inline const std::vector<uint64_t>* hnsw_nbrs(const orc_hnsw* g, uint64_t id, uint64_t layer) {
  auto it = g->nodes.find(id);
  if (it == g->nodes.end()) return nullptr;
  if (layer >= it->second.conn.size()) return nullptr;
  return &it->second.conn[layer];
}

// hnsw.rs:332-402 with the (dist,id) rule in place of the distance-only Candidate order.
int32_t hnsw_search_layer(const orc_hnsw* g, const float* q, uint64_t entry, uint64_t ef,
                          uint64_t layer, std::vector<Key>& out) {
  auto eit = g->nodes.find(entry);
  if (eit == g->nodes.end()) return ISL_NODE_NOT_FOUND;  // distance(), hnsw.rs:449-455
  std::unordered_map<uint64_t, char> visited;
  MinHeap candidates;
  MaxHeap results;
  float ed = calc(g->cfg.metric, q, eit->second.v.data(), g->d);
  visited[entry] = 1;
  candidates.push({ed, entry});
  results.push({ed, entry});
  while (!candidates.empty()) {
    Key c = candidates.top();
    candidates.pop();
    if (of_gt(c.d, results.top().d) && results.size() >= ef) break;  // :356-360
    const std::vector<uint64_t>* nb = hnsw_nbrs(g, c.id, layer);
    if (!nb) continue;
    for (uint64_t x : *nb) {
      if (!visited.emplace(x, 1).second) continue;
      auto xit = g->nodes.find(x);
      if (xit == g->nodes.end()) return ISL_NODE_NOT_FOUND;
      float nd = calc(g->cfg.metric, q, xit->second.v.data(), g->d);
      bool should_add = results.size() < ef || nd < results.top().d;
      if (should_add) {
        candidates.push({nd, x});
        results.push({nd, x});
        if (results.size() > ef) results.pop();
      }
    }
  }
  out.clear();
  while (!results.empty()) {
    out.push_back(results.top());
    results.pop();
  }
  std::sort(out.begin(), out.end(), key_lt);
  return ISL_OK;
}

// Greedy descent shared by insert (hnsw.rs:263-282) and search (hnsw.rs:478-497).
int32_t hnsw_greedy(const orc_hnsw* g, const float* q, uint64_t& current, float& current_dist,
                    uint64_t from_layer, uint64_t to_layer_inclusive) {
  for (uint64_t layer = from_layer; layer + 1 > to_layer_inclusive; --layer) {
    for (;;) {
      bool changed = false;
      const std::vector<uint64_t>* nb = hnsw_nbrs(g, current, layer);
      if (nb) {
        std::vector<uint64_t> snapshot = *nb;  // the loop iterates the list bound before `current` moves
        for (uint64_t x : snapshot) {
          auto xit = g->nodes.find(x);
          if (xit == g->nodes.end()) return ISL_NODE_NOT_FOUND;
          float dist = calc(g->cfg.metric, q, xit->second.v.data(), g->d);
          if (dist < current_dist) {
            current = x;
            current_dist = dist;
            changed = true;
          }
        }
      }
      if (!changed) break;
    }
    if (layer == 0) break;
  }
  return ISL_OK;
}
}  // namespace

orc_hnsw* orc_hnsw_new(const isl_hnsw_config* cfg, uint32_t d) {
  auto* g = new orc_hnsw();
  g->cfg = *cfg;
  g->d = d;
  return g;
}
void orc_hnsw_free(orc_hnsw* g) { delete g; }
uint64_t orc_hnsw_len(const orc_hnsw* g) { return g->nodes.size(); }
int64_t orc_hnsw_entry_point(const orc_hnsw* g) { return g->entry; }
uint64_t orc_hnsw_max_level(const orc_hnsw* g) { return g->max_level; }

int64_t orc_hnsw_neighbors(const orc_hnsw* g, uint64_t id, uint64_t layer, uint64_t* out,
                           uint64_t cap) {
  const std::vector<uint64_t>* nb = hnsw_nbrs(g, id, layer);
  if (!nb) return -1;
  for (uint64_t i = 0; i < nb->size() && i < cap; ++i) out[i] = (*nb)[i];
  return (int64_t)nb->size();
}

int32_t orc_hnsw_insert(orc_hnsw* g, const float* v, uint64_t level, uint64_t* out_id) {
  uint64_t id = g->next_id++;  // hnsw.rs:227-228
  if (out_id) *out_id = id;
  orc_hnsw::Node node;
  node.v.assign(v, v + g->d);
  node.level = level;
  node.conn.resize(level + 1);  // hnsw.rs:104-106
  if (g->entry < 0) {           // hnsw.rs:240-245
    g->entry = (int64_t)id;
    g->max_level = level;
    g->nodes.emplace(id, std::move(node));
    return ISL_OK;
  }
  const float* q = node.v.data();
  uint64_t current = (uint64_t)g->entry;  // hnsw.rs:259-260
  float current_dist = calc(g->cfg.metric, q, g->nodes.at(current).v.data(), g->d);
  if (g->max_level >= level + 1) {        // hnsw.rs:263-282
    int32_t st = hnsw_greedy(g, q, current, current_dist, g->max_level, level + 1);
    if (st != ISL_OK) return st;
  }
  std::vector<Key> found;
  for (uint64_t layer = level + 1; layer-- > 0;) {  // hnsw.rs:285-319
    int32_t st = hnsw_search_layer(g, q, current, g->cfg.ef_construction, layer, found);
    if (st != ISL_OK) return st;
    uint64_t mm = layer == 0 ? g->cfg.m0 : g->cfg.m;  // :290-294
    std::vector<uint64_t> selected;
    for (size_t i = 0; i < found.size() && i < mm; ++i) selected.push_back(found[i].id);  // :295
    node.conn[layer] = selected;                                                          // :298-300
    for (uint64_t nbid : selected) {  // :303-313
      auto it = g->nodes.find(nbid);
      if (it == g->nodes.end()) continue;
      if (layer >= it->second.conn.size()) continue;
      auto& conns = it->second.conn[layer];
      conns.push_back(id);
      if (conns.size() > mm) {
        // prune_connections (hnsw.rs:405-446): ids not present in `nodes` are filtered out;
        // the node being inserted is not in `nodes` yet (hnsw.rs:327), so it is dropped here.
        struct S {
          uint64_t id;
          float dist;
        };
        std::vector<S> scored;
        const float* nv = it->second.v.data();
        for (uint64_t x : conns) {
          auto xit = g->nodes.find(x);
          if (xit == g->nodes.end()) continue;
          scored.push_back({x, calc(g->cfg.metric, nv, xit->second.v.data(), g->d)});
        }
        std::stable_sort(scored.begin(), scored.end(),
                         [](const S& a, const S& b) { return a.dist < b.dist; });
        std::vector<uint64_t> pruned;
        for (size_t i = 0; i < scored.size() && i < mm; ++i) pruned.push_back(scored[i].id);
        conns = pruned;
      }
    }
    if (!selected.empty()) current = selected[0];  // :316-318
  }
  if (level > g->max_level) {  // :322-325
    g->max_level = level;
    g->entry = (int64_t)id;
  }
  g->nodes.emplace(id, std::move(node));  // :327
  return ISL_OK;
}

// Round model of HnswGraph::insert for GPU-parallel construction (islands_b200/csrc/hnsw.cu).
// `count` vectors are inserted in rounds of min(batch, max(1, len/2)) nodes.  Every node of a
// round runs the read-only half of insert_node (greedy descent hnsw.rs:263-282, per-layer
// search_layer + take(M) :285-300, entry for the next layer :316-318) against the graph as it stood
// before the round; then the mutating half (own lists :298-300, reverse edges + prune_connections
// :303-313, entry update :322-325, nodes.insert :327) is applied node by node in id order.
// Within one insert the two halves commute (a layer's reverse edges only touch that layer's lists,
// which no later search of the same insert reads), so batch = 1 is exactly orc_hnsw_insert.
int32_t orc_hnsw_insert_batch(orc_hnsw* g, const float* vectors, uint64_t count, const uint64_t* levels,
                              uint32_t batch, int32_t threads) {
  if (batch == 0) batch = 1;
  uint64_t done = 0;
  while (done < count) {
    const uint64_t have = g->nodes.size();
    if (have == 0) {  // hnsw.rs:240-245
      int32_t st = orc_hnsw_insert(g, vectors, levels[0], nullptr);
      if (st != ISL_OK) return st;
      done = 1;
      continue;
    }
    const uint64_t round = std::min<uint64_t>(count - done, std::min<uint64_t>(batch, std::max<uint64_t>(1, have / 2)));
    struct Plan {
      std::vector<std::vector<Key>> sel;  // per layer 0..=level: the first M of search_layer's result
    };
    std::vector<Plan> plans(round);
    std::vector<int32_t> status(round, ISL_OK);
    parallel_for(round, threads, [&](uint64_t b, uint64_t e, int) {
      std::vector<Key> found;
      for (uint64_t i = b; i < e; ++i) {
        const float* q = vectors + (done + i) * (uint64_t)g->d;
        const uint64_t level = levels[done + i];
        uint64_t current = (uint64_t)g->entry;
        float current_dist = calc(g->cfg.metric, q, g->nodes.at(current).v.data(), g->d);
        if (g->max_level >= level + 1) {
          int32_t st = hnsw_greedy(g, q, current, current_dist, g->max_level, level + 1);
          if (st != ISL_OK) { status[i] = st; continue; }
        }
        plans[i].sel.resize(level + 1);
        for (uint64_t layer = level + 1; layer-- > 0;) {
          int32_t st = hnsw_search_layer(g, q, current, g->cfg.ef_construction, layer, found);
          if (st != ISL_OK) { status[i] = st; break; }
          const uint64_t mm = layer == 0 ? g->cfg.m0 : g->cfg.m;
          for (size_t t = 0; t < found.size() && t < mm; ++t) plans[i].sel[layer].push_back(found[t]);
          if (!plans[i].sel[layer].empty()) current = plans[i].sel[layer][0].id;
        }
      }
    });
    for (int32_t st : status)
      if (st != ISL_OK) return st;
    for (uint64_t i = 0; i < round; ++i) {
      const uint64_t id = g->next_id++;
      const uint64_t level = levels[done + i];
      orc_hnsw::Node node;
      node.v.assign(vectors + (done + i) * (uint64_t)g->d, vectors + (done + i + 1) * (uint64_t)g->d);
      node.level = level;
      node.conn.resize(level + 1);
      for (uint64_t layer = level + 1; layer-- > 0;) {
        const uint64_t mm = layer == 0 ? g->cfg.m0 : g->cfg.m;
        for (const Key& kx : plans[i].sel[layer]) node.conn[layer].push_back(kx.id);
        for (const Key& kx : plans[i].sel[layer]) {
          auto it = g->nodes.find(kx.id);
          if (it == g->nodes.end() || layer >= it->second.conn.size()) continue;
          auto& conns = it->second.conn[layer];
          conns.push_back(id);
          if (conns.size() > mm) {  // prune_connections: the id being inserted is not in `nodes` yet
            struct S { uint64_t id; float dist; };
            std::vector<S> scored;
            const float* nv = it->second.v.data();
            for (uint64_t x : conns) {
              auto xit = g->nodes.find(x);
              if (xit == g->nodes.end()) continue;
              scored.push_back({x, calc(g->cfg.metric, nv, xit->second.v.data(), g->d)});
            }
            std::stable_sort(scored.begin(), scored.end(), [](const S& a, const S& b) { return a.dist < b.dist; });
            std::vector<uint64_t> pruned;
            for (size_t t = 0; t < scored.size() && t < mm; ++t) pruned.push_back(scored[t].id);
            conns = pruned;
          }
        }
      }
      if (level > g->max_level) {
        g->max_level = level;
        g->entry = (int64_t)id;
      }
      g->nodes.emplace(id, std::move(node));
    }
    done += round;
  }
  return ISL_OK;
}

int64_t orc_hnsw_node_level(const orc_hnsw* g, uint64_t id) {
  auto it = g->nodes.find(id);
  return it == g->nodes.end() ? -1 : (int64_t)it->second.level;
}

int32_t orc_hnsw_search(const orc_hnsw* g, const float* queries, uint64_t nq, uint32_t k,
                        uint32_t ef_in, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                        int32_t threads) {
  if (g->nodes.empty()) {  // hnsw.rs:459-461
    for (uint64_t qi = 0; qi < nq; ++qi) {
      std::vector<Key> none;
      write_topk(none, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
    }
    return ISL_OK;
  }
  if (g->entry < 0) return ISL_INDEX_NOT_BUILT;
  std::vector<int32_t> status(nq, ISL_OK);
  parallel_for(nq, threads, [&](uint64_t b, uint64_t e, int) {
    std::vector<Key> res;
    for (uint64_t qi = b; qi < e; ++qi) {
      const float* q = queries + qi * (uint64_t)g->d;
      uint64_t current = (uint64_t)g->entry;  // hnsw.rs:473-475
      float current_dist = calc(g->cfg.metric, q, g->nodes.at(current).v.data(), g->d);
      if (g->max_level >= 1) {                // hnsw.rs:478-497
        int32_t st = hnsw_greedy(g, q, current, current_dist, g->max_level, 1);
        if (st != ISL_OK) {
          status[qi] = st;
          continue;
        }
      }
      uint64_t ef = std::max<uint64_t>(ef_in, k);  // hnsw.rs:500
      int32_t st = hnsw_search_layer(g, q, current, ef, 0, res);
      if (st != ISL_OK) {
        status[qi] = st;
        continue;
      }
      write_topk(res, k, out_ids + qi * k, out_dist + qi * k, out_count ? out_count + qi : nullptr);
    }
  });
  for (int32_t st : status)
    if (st != ISL_OK) return st;
  return ISL_OK;
}

}  // extern "C"
