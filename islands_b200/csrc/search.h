// search.h — host-visible types of the batched best-first search (search_core.cuh / search.cu).
#pragma once

#include "common.cuh"

namespace isl {

// Lean ADC traversal: direct-mapped cache of admitted ids in shared memory.  An entry holds id >> kIdcBits as 16
// bits (the slot supplies the low bits), which identifies the id exactly for n <= 2^(16 + kIdcBits) - the host
// only enables the bitset-free traversal below that size.
constexpr uint32_t kIdcBits = 9;
constexpr uint32_t kIdcEntries = 1u << kIdcBits;
constexpr uint32_t kIdcMaxNodes = (0xffffu << kIdcBits);  // tag 0xffff is the empty marker

// Where a sharded search writes its per-query result records: 16 bytes {dist bits, 0, global id lo, hi}, laid out
// [nq][k].  `packed` is a local buffer (NCCL exchange); `peer[0 .. n_peer)` are this rank's slots in the gather
// buffers of all ranks of the node (its own included), mapped with CUDA IPC: the search kernel stores each
// finished query's records straight into them over NVLink, so the exchange overlaps the search query by query.
constexpr int kMaxPeers = 16;
struct ShardOut {
  uint4* packed = nullptr;
  uint64_t id_base = 0;     // added to every local id (node-range shards: first global id of the shard)
  uint4* peer[kMaxPeers] = {};
  uint32_t n_peer = 0;
};

struct SearchArgs {
  // resident index
  const float* vectors;     // [n][ld] f32, rows 16B aligned, zero padded to ld
  const float* sqnorms;     // [n] Σ y² folded in reference order (cosine only)
  uint32_t ld;
  uint32_t d;
  uint32_t n;
  const uint64_t* offsets;  // CSR node_offsets [n+1]; nullptr => fixed-stride adjacency
  const uint32_t* nbrs;     // CSR neighbours, or adjacency [n][adj_stride]
  const uint32_t* degrees;  // fixed-stride mode: live degree per node
  uint32_t adj_stride;
  uint32_t lists_unique;    // 1 => no list holds an id twice: the first-occurrence test of a hop is skipped
  uint32_t novis;           // lean ADC traversal with R in registers: no visited bitset (see search_core.cuh); needs stats == null
  // queries
  const float* queries;     // [nq][q_ld]
  uint32_t q_ld;
  uint32_t nq;
  uint32_t entry;
  uint32_t k;               // results written per query
  uint32_t ef;              // already max(ef, k)
  int32_t metric;
  float prune_ratio;        // leann.rs:991-1056 (0 => identity)
  int32_t strategy;
  uint64_t prune_seed;      // Proportional: seed of the draw stream
  const uint32_t* deg_counts;  // Proportional: graph.degree_counts (out-degree per node)
  // scratch (per resident warp slot)
  uint32_t* visited;        // [slots][vis_words]
  uint32_t vis_words;
  uint2* r_global;          // [slots][ef] when R does not fit shared memory
  uint2* ties_global;       // [slots][ef] spill area of the tie list (search_core.cuh)
  uint32_t u_cap;           // capacity of the unvisited list (>= max degree, multiple of 32)
  // outputs
  uint64_t* out_ids;        // [nq][k] (may be null)
  uint32_t* out_ids32;      // [nq][k] (may be null; build path)
  float* out_dist;          // [nq][k]
  uint32_t* out_count;      // [nq]
  ShardOut shard;           // shard-exchange records (api_shard.cu)
  isl_search_stats* stats;  // [nq] or null
  unsigned int* work_counter;
  unsigned int* error_flag; // internal invariant guard (never set: the tie list cannot overflow its kTieCap + ef entries)
  // two-level search: PQ ADC traversal + exact rerank (docs/leann-specification.md:223-269)
  const float* luts;        // [nq][pq_m*pq_ksub] squared-L2 tables (pq.rs:307-338); null => built per query in shared memory
  const float* pq_codebooks; // [pq_m][pq_ksub][pq_ld_sub] (only read when luts is null)
  uint32_t pq_dsub, pq_ld_sub;
  const uint8_t* codes8;    // [n][pq_m] when pq_ksub <= 256
  const uint16_t* codes16;  // [n][pq_m] otherwise
  uint32_t pq_m, pq_ksub;
  float rerank_ratio;
  uint32_t aq_cap;          // capacity of the approximate queue (entries)
  uint2* aq_global;         // [slots][aq_cap] when the queue does not fit shared memory
  uint32_t lut_smem_floats; // pq_m*pq_ksub when the table is staged in shared memory, else 0
  uint32_t aq_smem_entries; // aq_cap when the queue lives in shared memory, else 0
  // HNSW (hnsw.rs): per-query entry points, queries taken from the vector table, layer row pools
  const uint32_t* entries;   // entry node per query (null => `entry`): [nq], or indexed by node id when query_ids is set
  const uint32_t* query_ids; // [nq] query i is vectors[query_ids[i]] (null => `queries`)
  const uint32_t* row_map;   // node -> first adjacency row in a compact pool (null => row = node)
  uint32_t row_add;          // added to row_map[node] (layer - 1)
  // recompute (MODE 2 split in two launches around the encoder): phase 1 = ADC traversal only, the
  // ef survivors are written to surv_ids / surv_cnt; phase 2 = exact rerank only of cand lists
  // against rows addressed through row_of_id (recomputed embeddings live in a compact matrix)
  uint32_t phase;              // 0 = both in one launch
  uint32_t rerank_limit;       // phase 1: hand over at most this many survivors, best table distance first (0 = all ef)
  uint32_t* surv_ids;          // [nq][ef]   (phase 1 out, phase 2 in)
  uint32_t* surv_cnt;          // [nq]
  const uint32_t* row_of_id;   // node id -> row of `vectors` / `sqnorms` (null => identity)
  const uint32_t* node_levels; // [n] top layer of every node; a node has no list above it (hnsw.rs:108-110)
  uint32_t layer;            // layer being searched (only read when node_levels is set)
};

struct SearchPlan {
  size_t smem = 0;        // dynamic shared memory per one-warp CTA
  int ctas_per_sm = 0;    // resident warps (= queries in flight) per SM
  uint32_t grid = 0;      // resident warp slots on the device; scratch is sized for this many
  bool r_in_smem = true;  // result array in shared memory (else L2-resident global memory)
  int acc = 0;            // accumulation kind (dist_pass.cuh)
  bool two_level = false;
  int mode = 0;           // 0 exact, 1 two-level (AQ promotion), 2 ADC traversal + exact rerank, 3 ADC traversal only
  int ks = 0;             // register-bag kernel: 128 = the table fold specialised for ksub == 128, 0 = any ksub
  int nr = 0;             // MODE 3: the register-bag kernel (adc_traverse.cuh), nr entries per lane (0 = search_core.cuh, shared / global memory)
  bool novis_ok = false;  // MODE 3: the launch may run without the visited bitset (SearchArgs::novis)
  uint32_t lut_smem_floats = 0;  // PQ table staged in shared memory (0 => read from global/L2)
  uint32_t aq_smem_entries = 0;  // approximate queue in shared memory (0 => global/L2)
  uint32_t aq_cap = 0;
};

// Chooses the kernel variant, opts into the shared-memory size and reports the slot count.
isl_status plan_search(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan);
isl_status plan_search_adc(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, uint32_t pq_m,
                           uint32_t pq_ksub, int sms, SearchPlan* plan);
// The two launches of "ADC traversal + exact rerank": the lean traversal (MODE 3, survivors out) and the
// rerank of candidate lists (MODE 2, SearchArgs::phase = 2; no PQ table in shared memory).
isl_status plan_search_adc_traverse(uint32_t ef, uint32_t u_cap, uint32_t pq_m, uint32_t pq_ksub, bool one_byte_codes, int sms,
                                    SearchPlan* plan);
isl_status plan_search_rerank(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, int sms, SearchPlan* plan);
isl_status plan_search_two_level(int32_t metric, uint32_t ld, uint32_t ef, uint32_t u_cap, uint32_t pq_m,
                                 uint32_t pq_ksub, uint32_t aq_cap, int sms, SearchPlan* plan);
// Enqueues the search.  args.visited / args.r_global must cover plan.grid slots and
// *args.work_counter / *args.error_flag must be zero.
isl_status launch_search(const SearchPlan& plan, const SearchArgs& args, cudaStream_t st);

}  // namespace isl
