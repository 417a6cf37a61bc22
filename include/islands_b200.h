/*
 * islands_b200.h — C ABI of the B200-native LEANN / HNSW search hot path.
 *
 * This is the drop-in boundary for the `src/core` vector-index API of panbanda/islands
 * (reference paths below are relative to the reference repository root).  Every entry
 * point is `extern "C"`, takes plain pointers and sizes, and returns an `isl_status`
 * that mirrors the reference's `CoreError` (src/core/error.rs:9-62).  No torch types,
 * no C++ types, no exceptions cross this boundary.
 *
 * Conventions kept from the reference:
 *   - empty index  -> search returns ISL_OK with zero results (leann.rs:875-877, hnsw.rs:459-461)
 *   - ef := max(ef, k)                                       (leann.rs:890, hnsw.rs:500)
 *   - results ascending by distance, at most k per query     (leann.rs:984-987, :895)
 *   - ties are broken by (distance, id) — the order of the reference's own heaps
 *     (leann.rs:701-702, :907-908) applied to the final sort as well
 *   - search takes a const handle (reference: &self) and is re-entrant
 *   - ids are u64 at the ABI (reference type); on the device they are u32 (n < 2^31)
 *
 * Host pointers are copied to / from the device inside the call.  Entry points with a
 * `_dev` suffix take device pointers on the handle's CUDA device and do no host copies.
 * Every call runs on an internal stream of its own (searches lease one per call, so several
 * threads can search one handle at the same time) and returns after that stream has been
 * synchronised: outputs are complete on return.  Device INPUTS of a `_dev` call may still be
 * being written on the caller's stream: the call first waits (on the device, cudaStreamWaitEvent)
 * for everything enqueued so far on the caller's stream — the legacy default stream unless
 * isl_set_caller_stream() named another one for the calling thread.
 *
 * There is NO CPU fallback: if no CUDA device is usable every compute entry point
 * returns ISL_CUDA_ERROR and isl_last_error() says why.
 */
#ifndef ISLANDS_B200_H
#define ISLANDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISL_ABI_VERSION 2
#define ISL_NO_ENTRY (-1)
#define ISL_INVALID_ID UINT64_MAX

/* Mirrors CoreError (src/core/error.rs:9-62). */
typedef enum isl_status {
  ISL_OK = 0,
  ISL_DIM_MISMATCH = 1,     /* CoreError::DimensionMismatch{expected,actual} */
  ISL_EMPTY_COLLECTION = 2, /* CoreError::EmptyCollection */
  ISL_INVALID_CONFIG = 3,   /* CoreError::InvalidConfig(String) */
  ISL_INDEX_NOT_BUILT = 4,  /* CoreError::IndexNotBuilt */
  ISL_NODE_NOT_FOUND = 5,   /* CoreError::NodeNotFound(u64) */
  ISL_PQ_ERROR = 6,         /* CoreError::PQError(String) */
  ISL_SERIALIZATION = 7,    /* CoreError::Serialization / Deserialization */
  ISL_CUDA_ERROR = 8,       /* no reference analogue: device failure, never a fallback */
  ISL_INVALID_ARGUMENT = 9  /* null pointer / out-of-range argument at the ABI */
} isl_status;

/* DistanceMetric (src/core/distance.rs:9-19); values = serde variant index. */
typedef enum isl_metric {
  ISL_METRIC_COSINE = 0,
  ISL_METRIC_EUCLIDEAN = 1,
  ISL_METRIC_DOT = 2,
  ISL_METRIC_MANHATTAN = 3
} isl_metric;

/* PruningStrategy (src/core/leann.rs:168-178). */
typedef enum isl_pruning_strategy {
  ISL_PRUNE_GLOBAL = 0,
  ISL_PRUNE_LOCAL = 1,
  ISL_PRUNE_PROPORTIONAL = 2 /* the reference draws from thread_rng (leann.rs:1043); here from the seeded stream below */
} isl_pruning_strategy;

/* PruningStrategy::Proportional keeps candidate j of a hop when gen::<f32>() < prob_j * num_to_keep
 * (leann.rs:1017-1053).  The reference draws from thread_rng, which no one can reproduce; this library (and its
 * oracle) draws from a counter-based stream instead: draw c (0, 1, 2, ... in the order the reference's loop consumes
 * them) of query q of a call (its row in the query matrix) with isl_leann_config::prune_seed is
 *   mix(z) = splitmix64's finaliser (z ^= z >> 30; z *= 0xBF58476D1CE4E5B9; z ^= z >> 27; z *= 0x94D049BB133111EB; z ^= z >> 31)
 *   h = mix(mix(prune_seed + 0x9E3779B97F4A7C15 * (q + 1)) + 0x9E3779B97F4A7C15 * (c + 1));   u = (h >> 40) * 2^-24
 * — a uniform f32 in [0, 1) with 24 random bits, the shape of rand's gen::<f32>(). */

/* LeannConfig (src/core/leann.rs:322-371), same field order. */
typedef struct isl_leann_config {
  uint64_t m;
  uint64_t m0;
  uint64_t ef_construction;
  double ml;
  uint64_t max_layers;
  int32_t metric; /* isl_metric */
  uint64_t ef_search;
  uint64_t beam_width;
  float prune_ratio;
  int32_t pruning_strategy; /* isl_pruning_strategy */
  int32_t high_degree_pruning;
  float hub_percentile;
  int32_t is_compact;
  int32_t is_recompute;
  uint64_t prune_seed; /* not a reference field: seed of the Proportional draw stream (see isl_pruning_strategy) */
} isl_leann_config;

/* HnswConfig (src/core/hnsw.rs:15-28). */
typedef struct isl_hnsw_config {
  uint64_t m;
  uint64_t m0;
  uint64_t ef_construction;
  double ml;
  int32_t metric;
  uint64_t max_layers;
} isl_hnsw_config;

/* PQConfig (src/core/pq.rs:13-22); seed: Option<u64> = (has_seed != 0 ? Some(seed) : None). */
typedef struct isl_pq_config {
  uint64_t num_subquantizers;
  uint64_t num_centroids;
  uint64_t training_iterations;
  uint64_t seed;
  int32_t has_seed;
} isl_pq_config;

/* Per-query traversal counters (the reference counts `embeddings_computed`, leann.rs:920,950).
 * n_hop   = expanded candidates (loop iterations that reached the neighbour fetch)
 * n_edge  = neighbour ids read
 * n_dist  = exact distances evaluated (entry included)
 * n_adc   = PQ table distances evaluated (two-level search only)
 * n_rerank= exact distances evaluated for promoted nodes (two-level search only; subset of n_dist) */
typedef struct isl_search_stats {
  uint64_t n_hop;
  uint64_t n_edge;
  uint64_t n_dist;
  uint64_t n_adc;
  uint64_t n_rerank;
} isl_search_stats;

/* What LeannIndex::build did (construction roofline): the traversal counters summed over the efConstruction
 * searches of all inserts (same meaning as isl_search_stats), the edges of the finished graph, and CUDA-event
 * times: search_ms = the search launches only, rounds_ms = search + selection + sort + reverse-edge kernels. */
typedef struct isl_build_stats {
  uint64_t n_hop;
  uint64_t n_edge;
  uint64_t n_dist;
  uint64_t edges;
  uint64_t rounds;
  float search_ms;
  float rounds_ms;
} isl_build_stats;

/* Shape of the recompute encoder: BERT (the reference's CandleEmbedder loads a BERT-family model,
 * src/core/embedding/candle_provider.rs:230-300).  head dimension is fixed at 64. */
typedef struct isl_encoder_config {
  uint32_t vocab_size;
  uint32_t hidden_size;
  uint32_t num_layers;
  uint32_t num_heads;
  uint32_t intermediate_size;
  uint32_t max_position;
  uint32_t type_vocab_size;
  float layer_norm_eps;
  int32_t normalize; /* EmbeddingConfig::normalize: L2-normalise the pooled vector */
  int32_t precision; /* ISL_ENCODER_BF16 (default) or ISL_ENCODER_BF16X3 */
} isl_encoder_config;
/* GEMM operand precision of the recompute encoder.  BF16: activations and weights rounded to bf16 (f32 accumulate).
 * BF16X3: split precision — every operand is carried as hi + lo bf16 and each product as hi.hi + hi.lo + lo.hi on the
 * same tensor-core kernel (3x the FLOPs), activations stay f32 between the GEMMs: embeddings agree with an f32 forward
 * to ~1e-6, for callers that need rank-level agreement with f32 embeddings (recall within 0.002). */
#define ISL_ENCODER_BF16 0
#define ISL_ENCODER_BF16X3 1

typedef struct isl_index isl_index; /* LeannIndex + CsrGraph + resident vectors (leann.rs:193-208, :493-500) */
typedef struct isl_pq isl_pq;       /* ProductQuantizer (pq.rs:116-129) */
typedef struct isl_hnsw isl_hnsw;   /* HnswGraph (hnsw.rs:151-164) */
typedef struct isl_encoder isl_encoder; /* CandleEmbedder's model (candle_provider.rs) for on-demand recompute */
typedef struct isl_shard isl_shard; /* one rank's membership in a sharded index: communicator + exchange buffers */

/* One entry of a per-shard result list as it travels between GPUs: 16 bytes. */
typedef struct isl_shard_record {
  float dist;
  uint32_t reserved; /* 0 */
  uint64_t id;       /* global id (local id + the shard's id_base); ISL_INVALID_ID pads short lists */
} isl_shard_record;

/* ---- library ---------------------------------------------------------------------- */
int isl_abi_version(void);
/* Thread-local message of the last failing call on this thread (String payloads of CoreError). */
const char* isl_last_error(void);
/* Numeric payload of the last failing call on this thread: ISL_DIM_MISMATCH -> (*a, *b) = (expected, actual);
 * ISL_NODE_NOT_FOUND -> *a = the node id; otherwise zeros.  Lets a binding rebuild the full CoreError variant. */
void isl_last_error_detail(uint64_t* a, uint64_t* b);
/* Number of usable CUDA devices (0 => every compute call fails with ISL_CUDA_ERROR). */
int isl_device_count(void);
/* Launch bookkeeping for benchmarks: kernels launched by this library since the last reset. */
uint64_t isl_kernel_launch_count(void);
void isl_kernel_launch_count_reset(void);
/* Names the CUDA stream (a cudaStream_t; NULL = the legacy default stream) on which the calling thread
 * produces the device buffers it hands to `_dev` entry points.  Thread-local. */
isl_status isl_set_caller_stream(void* cuda_stream);

/* ---- configs (leann.rs:373-461, hnsw.rs:37-85, pq.rs:24-65) ---------------------- */
isl_status isl_leann_config_default(isl_leann_config* out);  /* paper_default(): m=30,m0=60,efC=128 */
isl_status isl_leann_config_fast(isl_leann_config* out);     /* leann.rs:406-416 */
isl_status isl_leann_config_accurate(isl_leann_config* out); /* leann.rs:419-429 */
isl_status isl_leann_config_validate(const isl_leann_config* cfg);
isl_status isl_hnsw_config_default(isl_hnsw_config* out);
isl_status isl_hnsw_config_validate(const isl_hnsw_config* cfg);
isl_status isl_pq_config_default(isl_pq_config* out);
isl_status isl_pq_config_validate(const isl_pq_config* cfg, uint64_t dimension);
uint64_t isl_pq_config_bytes_per_vector(const isl_pq_config* cfg); /* pq.rs:58-64 */

/* ---- distances (src/core/distance.rs) --------------------------------------------- */
/* Distance::calculate (distance.rs:37-52): one pair; len_a != len_b -> ISL_DIM_MISMATCH. */
isl_status isl_distance_calculate(int32_t metric, const float* a, uint64_t len_a, const float* b,
                                  uint64_t len_b, float* out);
/* Distance::calculate_squared (distance.rs:54-66). */
isl_status isl_distance_calculate_squared(int32_t metric, const float* a, uint64_t len_a,
                                          const float* b, uint64_t len_b, float* out);
/* Distance::batch_calculate (distance.rs:32-34): rows is [n_rows][dim] row-major. */
isl_status isl_distance_batch(int32_t metric, const float* query, const float* rows,
                              uint64_t n_rows, uint32_t dim, float* out);
isl_status isl_distance_batch_dev(int32_t metric, const float* d_query, const float* d_rows,
                                  uint64_t n_rows, uint32_t dim, float* d_out);
/* normalize_vector (distance.rs:125-132) applied to each of n_rows rows in place. */
isl_status isl_normalize_rows(float* rows, uint64_t n_rows, uint32_t dim);

/* ---- LEANN index (src/core/leann.rs) ----------------------------------------------- */
/* Adopt an existing CSR graph ("identical graphs" entry): CsrGraph fields (leann.rs:193-208)
 * node_offsets [n+1], neighbors [node_offsets[n]], levels [n] (may be NULL -> zeros),
 * entry_point (ISL_NO_ENTRY = None).  `vectors` [n][dim] are the embeddings the
 * InMemoryEmbeddingProvider (leann.rs:104-159) would return; they are made resident in HBM. */
isl_status isl_index_from_csr(const isl_leann_config* cfg, uint32_t dim, uint64_t n,
                              const uint64_t* node_offsets, const uint64_t* neighbors,
                              const uint64_t* levels, int64_t entry_point, const float* vectors,
                              isl_index** out);
/* LeannIndex::build (leann.rs:560-631).  levels_or_null: explicit per-node levels (the reference
 * draws them from thread_rng, leann.rs:549-554); NULL -> drawn from `seed` with the same formula.
 * batch = 1 reproduces the reference's sequential insertion order exactly; batch > 1 inserts
 * `batch` nodes against one graph snapshot (GPU-parallel construction). */
isl_status isl_index_build(const isl_leann_config* cfg, uint32_t dim, uint64_t n,
                           const float* vectors, const uint64_t* levels_or_null, uint64_t seed,
                           uint32_t batch, isl_index** out);
isl_status isl_index_build_dev(const isl_leann_config* cfg, uint32_t dim, uint64_t n,
                               const float* d_vectors, const uint64_t* levels_or_null,
                               uint64_t seed, uint32_t batch, isl_index** out);
/* Counters and timings of the construction behind this handle (zeros when the graph was adopted, not built). */
isl_status isl_index_last_build_stats(const isl_index* idx, isl_build_stats* out);
void isl_index_free(isl_index* idx);
uint64_t isl_index_len(const isl_index* idx);            /* LeannIndex::len */
uint32_t isl_index_dimension(const isl_index* idx);      /* LeannIndex::dimension (0 = None) */
uint64_t isl_index_num_edges(const isl_index* idx);      /* graph.neighbors.len() */
int64_t isl_index_entry_point(const isl_index* idx);     /* graph.entry_point */
uint64_t isl_index_max_level(const isl_index* idx);      /* graph.max_level */
uint64_t isl_index_storage_bytes(const isl_index* idx);  /* CsrGraph::storage_bytes (leann.rs:296-301) */
/* Copy the CSR arrays out; any pointer may be NULL.  node_offsets [n+1], neighbors [num_edges],
 * levels [n], degree_counts [n]. */
isl_status isl_index_export_csr(const isl_index* idx, uint64_t* node_offsets, uint64_t* neighbors,
                                uint64_t* levels, uint64_t* degree_counts);
/* CsrGraph::get_neighbors (leann.rs:225-233): returns count, writes up to cap ids;
 * node_id >= n -> ISL_NODE_NOT_FOUND (reference: None). */
isl_status isl_index_get_neighbors(const isl_index* idx, uint64_t node_id, uint64_t* out,
                                   uint64_t cap, uint64_t* out_count);

/* CsrGraph::set_neighbors (leann.rs:256-293): replace the list of node_id (graph arrays rebuilt when the
 * length changes, as in the reference); node_id >= n is ignored as in the reference (leann.rs:258-260), a
 * neighbour id >= n -> ISL_NODE_NOT_FOUND.
 * Mutating (&mut self in the reference): must not run concurrently with searches on the handle. */
isl_status isl_index_set_neighbors(isl_index* idx, uint64_t node_id, const uint64_t* neighbors, uint64_t count);

/* Batched LeannIndex::search_with_params (leann.rs:868-896) over nq queries [nq][dim].
 * out_ids/out_dist are [nq][k], padded with ISL_INVALID_ID / +inf; out_count [nq];
 * stats_or_null [nq].  query_dim != index dimension -> ISL_DIM_MISMATCH. */
isl_status isl_index_search(const isl_index* idx, const float* queries, uint64_t nq,
                            uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                            float* out_dist, uint32_t* out_count, isl_search_stats* stats_or_null);
isl_status isl_index_search_dev(const isl_index* idx, const float* d_queries, uint64_t nq,
                                uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                                float* d_out_dist, uint32_t* d_out_count,
                                isl_search_stats* d_stats_or_null);
/* LeannIndex::search (leann.rs:858-865): ef = config.ef_search. */
isl_status isl_index_search_default(const isl_index* idx, const float* queries, uint64_t nq,
                                    uint32_t query_dim, uint32_t k, uint64_t* out_ids,
                                    float* out_dist, uint32_t* out_count);
/* Duration in milliseconds (CUDA events on the handle's stream) of the search kernel of the
 * last isl_index_search* call on this handle, and the algorithmic HBM bytes it moved. */
isl_status isl_index_last_search_timing(const isl_index* idx, float* kernel_ms,
                                        uint64_t* kernel_launches);

/* ---- Product quantizer (src/core/pq.rs) -------------------------------------------- */
isl_status isl_pq_new(uint32_t dimension, const isl_pq_config* cfg, isl_pq** out); /* pq.rs:133-149 */
void isl_pq_free(isl_pq* pq);
isl_status isl_pq_set_metric(isl_pq* pq, int32_t metric);                          /* with_metric, pq.rs:152-155 */
int32_t isl_pq_is_trained(const isl_pq* pq);
uint64_t isl_pq_num_subquantizers(const isl_pq* pq);
float isl_pq_compression_ratio(const isl_pq* pq);                                   /* pq.rs:168-172 */
/* ProductQuantizer::train (pq.rs:175-218): vectors [n][dimension]. */
isl_status isl_pq_train(isl_pq* pq, const float* vectors, uint64_t n, uint32_t dim);
/* Install codebooks trained elsewhere: [m][ksub][dsub]; marks the quantizer trained. */
isl_status isl_pq_set_codebooks(isl_pq* pq, const float* codebooks, uint64_t num_centroids);
isl_status isl_pq_get_codebooks(const isl_pq* pq, float* out, uint64_t* out_num_centroids);
/* Batched ProductQuantizer::encode (pq.rs:221-244): codes [n][m] u16. */
isl_status isl_pq_encode(const isl_pq* pq, const float* vectors, uint64_t n, uint32_t dim,
                         uint16_t* out_codes);
/* Batched decode (pq.rs:247-271): codes [n][codes_per_vector]; out [n][dimension]. */
isl_status isl_pq_decode(const isl_pq* pq, const uint16_t* codes, uint64_t n,
                         uint64_t codes_per_vector, float* out);
/* build_distance_tables (pq.rs:307-338): tables [m][ksub]. */
isl_status isl_pq_build_tables(const isl_pq* pq, const float* query, uint32_t dim, float* out_tables);
/* table_distance (pq.rs:341-348) for n code rows against one table set. */
isl_status isl_pq_table_distance(const isl_pq* pq, const float* tables, const uint16_t* codes,
                                 uint64_t n, float* out);
/* asymmetric_distance (pq.rs:275-304) for n code rows against one query. */
isl_status isl_pq_asymmetric_distance(const isl_pq* pq, const float* query, uint32_t dim,
                                      const uint16_t* codes, uint64_t n, float* out);

#ifdef ISL_TEST_HOOKS
/* Test hook (not part of the product ABI; compiled into the library, declared only for the tests): the generator
 * isl_pq_train draws from when config.seed is set — `StdRng::seed_from_u64(seed)` of rand 0.8.5 (ChaCha12;
 * pq.rs:190-193), restated in csrc/std_rng.h.  A scripted sequence of draws: kinds[i] = 0 next_u32, 1 next_u64
 * (gen::<usize>()), 2 gen::<f32>() (bit pattern), 3 SliceRandom::choose index over `bound` elements.  Host code only. */
isl_status isl_std_rng_draw(uint64_t seed, const uint8_t* kinds, uint64_t count, uint64_t bound, uint64_t* out);
/* Test hook: the bfloat16 rounding the ADC traversal applies to its table entries (isl_index_search_adc_rerank below;
 * csrc/common.cuh bf16_round_bits), element by element.  Host code only. */
isl_status isl_adc_table_round(const float* in, uint64_t count, float* out);
/* Test hook: throws a C++ exception inside an entry point (0 = std::bad_alloc, 1 = std::length_error, 2 = a non-std
 * type).  Every isl_status entry point catches at the boundary: the call returns ISL_INVALID_ARGUMENT with a message,
 * nothing unwinds into the caller. */
isl_status isl_test_raise(int32_t kind);
#endif

/* ---- two-level search (docs/leann-specification.md:223-269; no reference code) ------ */
/* Attach PQ codes [n][m] (u16 at the ABI) for ADC-carried traversal; rerank_ratio = `a`. */
isl_status isl_index_attach_pq(isl_index* idx, const isl_pq* pq, const uint16_t* codes);
isl_status isl_index_search_two_level(const isl_index* idx, const float* queries, uint64_t nq,
                                      uint32_t query_dim, uint32_t k, uint32_t ef,
                                      float rerank_ratio, uint64_t* out_ids, float* out_dist,
                                      uint32_t* out_count, isl_search_stats* stats_or_null);

/* "PQ ADC traversal + exact rerank": the best-first search of leann.rs:899-988 runs entirely on
 * table distances (pq.rs:341-348; same admission / termination / tie rules with adc as the key),
 * then the ef surviving candidates get their exact distance (distance.rs) and are returned sorted by
 * (distance, id).  Traversal reads m code bytes per visited node instead of 4*dim.
 * Table entries of the TRAVERSAL: every entry of build_distance_tables (pq.rs:307-338) is rounded to bfloat16
 * (round to nearest even on the f32 bit pattern, NaN -> quiet NaN) before it is folded; table_distance is then the
 * f32 left fold over the subquantizers of those entries and a square root (pq.rs:341-348 otherwise unchanged).  The
 * rounding error (2^-9 relative per entry) is two orders of magnitude below the quantisation error of the codes it
 * ranks, and the per-query table is 2 bytes per entry in shared memory — twice the resident queries per SM.  Only
 * the traversal's ranking uses these values: isl_pq_build_tables / isl_pq_table_distance / isl_pq_asymmetric_distance and the
 * two-level search keep the f32 tables, and every returned distance is exact.  The oracle's twin
 * (orc_leann_search_adc_rerank) follows the same rule; like the mode itself it is defined here, not by the reference.
 * With stats_or_null == NULL (one-byte codes, m = 16 / 32, ef <= 2048, n < 2^25) the traversal keeps no
 * visited set — a node may be scored more than once, which cannot change the result —;
 * with statistics it keeps the exact bitset, and n_adc counts distinct nodes scored.  Ids and distances
 * are the same either way. */
isl_status isl_index_search_adc_rerank(const isl_index* idx, const float* queries, uint64_t nq,
                                       uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                       float* out_dist, uint32_t* out_count,
                                       isl_search_stats* stats_or_null);

/* ---- HNSW graph (src/core/hnsw.rs) -------------------------------------------------- */
/* HnswGraph::new (hnsw.rs:167-178): empty graph; the dimension is fixed by the first insert. */
isl_status isl_hnsw_new(const isl_hnsw_config* cfg, isl_hnsw** out);
void isl_hnsw_free(isl_hnsw* g);
uint64_t isl_hnsw_len(const isl_hnsw* g);           /* HnswGraph::len */
uint32_t isl_hnsw_dimension(const isl_hnsw* g);     /* HnswGraph::dimension (0 = None) */
int64_t isl_hnsw_entry_point(const isl_hnsw* g);    /* entry_point (ISL_NO_ENTRY = None) */
uint64_t isl_hnsw_max_level(const isl_hnsw* g);
/* `count` calls of HnswGraph::insert (hnsw.rs:214-329) for vectors [count][dim]; ids are
 * len()..len()+count-1 and *out_first_id is the first.  levels_or_null: explicit level per vector
 * (the reference draws them from thread_rng, hnsw.rs:206-211); NULL -> drawn from `seed` with the
 * same formula.  batch = 1 is the reference's sequential insertion; batch > 1 inserts up to `batch`
 * nodes against one graph snapshot (GPU-parallel construction).  dim != dimension() ->
 * ISL_DIM_MISMATCH. */
isl_status isl_hnsw_insert_batch(isl_hnsw* g, const float* vectors, uint64_t count, uint32_t dim,
                                 const uint64_t* levels_or_null, uint64_t seed, uint32_t batch,
                                 uint64_t* out_first_id);
isl_status isl_hnsw_insert_batch_dev(isl_hnsw* g, const float* d_vectors, uint64_t count, uint32_t dim,
                                     const uint64_t* levels_or_null, uint64_t seed, uint32_t batch,
                                     uint64_t* out_first_id);
/* HnswNode::level / neighbors_at (hnsw.rs:90-125) of get_node(node_id) (hnsw.rs:201-203). */
isl_status isl_hnsw_node_level(const isl_hnsw* g, uint64_t node_id, uint64_t* out_level);
isl_status isl_hnsw_get_neighbors(const isl_hnsw* g, uint64_t node_id, uint64_t layer, uint64_t* out,
                                  uint64_t cap, uint64_t* out_count);
/* All lists of one layer: out_degrees [len] (-1 = the node does not reach the layer),
 * out_neighbors [len][m0 or m] padded with ISL_INVALID_ID; either pointer may be NULL. */
isl_status isl_hnsw_export_layer(const isl_hnsw* g, uint64_t layer, int64_t* out_degrees,
                                 uint64_t* out_neighbors);
/* Batched HnswGraph::search (hnsw.rs:458-504): greedy descent from the entry point through the
 * upper layers, then best-first search of layer 0 with ef := max(ef, k). */
isl_status isl_hnsw_search(const isl_hnsw* g, const float* queries, uint64_t nq, uint32_t query_dim,
                           uint32_t k, uint32_t ef, uint64_t* out_ids, float* out_dist,
                           uint32_t* out_count);
isl_status isl_hnsw_search_dev(const isl_hnsw* g, const float* d_queries, uint64_t nq,
                               uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                               float* d_out_dist, uint32_t* d_out_count);
isl_status isl_hnsw_last_search_timing(const isl_hnsw* g, float* kernel_ms);

/* ---- on-demand embedding recomputation (src/core/embedding/candle_provider.rs:353-507) ------- */
/* BERT-base shape (110M parameters): vocab 30522, hidden 768, 12 layers, 12 heads, FFN 3072. */
isl_status isl_encoder_config_default(isl_encoder_config* out);
isl_status isl_encoder_new(const isl_encoder_config* cfg, isl_encoder** out);
void isl_encoder_free(isl_encoder* enc);
uint32_t isl_encoder_dimension(const isl_encoder* enc);          /* EmbeddingProvider::dimension */
uint64_t isl_encoder_num_parameters(const isl_encoder* enc);
/* Random-init weights of the configured shape: matrices and embeddings ~ N(0, stddev) from a
 * counter-based generator keyed by `seed`, biases 0, LayerNorm weight 1 / bias 0. */
isl_status isl_encoder_init_random(isl_encoder* enc, uint64_t seed, float stddev);
/* Parameters by their Hugging Face BERT names ("embeddings.word_embeddings.weight",
 * "encoder.layer.3.attention.self.query.weight", ...), f32 row-major [out][in]. */
isl_status isl_encoder_set_parameter(isl_encoder* enc, const char* name, const float* data, uint64_t count);
isl_status isl_encoder_get_parameter(const isl_encoder* enc, const char* name, float* out, uint64_t count);
/* embed_texts_raw after tokenisation: token_ids [batch][seq_len] (padded with 0), lengths [batch] =
 * number of attended tokens per row (attention_mask = 1 for the first lengths[b] positions);
 * out [batch][hidden]: BERT forward (bf16 tensor-core GEMMs, f32 accumulate) -> masked mean pooling
 * with clamp(sum_mask, 1e-9) -> L2 normalisation with clamp(norm, 1e-12) when cfg.normalize. */
isl_status isl_encoder_embed(isl_encoder* enc, const int32_t* token_ids, const int32_t* lengths,
                             uint64_t batch, uint32_t seq_len, float* out);
isl_status isl_encoder_embed_dev(isl_encoder* enc, const int32_t* d_token_ids, const int32_t* d_lengths,
                                 uint64_t batch, uint32_t seq_len, float* d_out);
/* CUDA-event duration of the last embed call and the FLOPs it performed (GEMMs + attention). */
isl_status isl_encoder_last_timing(const isl_encoder* enc, float* ms, double* flops);
/* The encoder's dense contraction on its own (tcgen05): out = act(A[m][k] * W[n][k]^T + bias)
 * (+ residual), A / W / residual / out_bf16 are bf16 on the device; k and n multiples of 64. */
isl_status isl_gemm_bf16_dev(const void* d_a_bf16, const void* d_w_bf16, uint32_t m, uint32_t n, uint32_t k,
                             const float* d_bias, const void* d_residual_bf16, int32_t gelu,
                             void* d_out_bf16, float* d_out_f32);

/* ---- search with on-demand recompute (leann.rs:82-99 EmbeddingProvider seam, :947-950) ---------- */
/* Attach the recompute provider: node i's embedding is encoder(token_ids[i][0..seq_len), lengths[i]).
 * enc == NULL detaches.  The encoder must outlive the attachment; its dimension must equal the
 * index dimension (else ISL_DIM_MISMATCH). */
isl_status isl_index_set_recompute(isl_index* idx, isl_encoder* enc, const int32_t* token_ids,
                                   const int32_t* lengths, uint32_t seq_len);
/* Free the resident f32 embeddings (LEANN's storage saving): only the recompute search remains. */
isl_status isl_index_drop_vectors(isl_index* idx);
/* PQ ADC traversal -> recompute the distinct ef-survivors of the batch with the bf16 encoder ->
 * exact rerank (reference-order f32 distances) against the recomputed embeddings. */
isl_status isl_index_search_adc_recompute(const isl_index* idx, const float* queries, uint64_t nq,
                                          uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                          float* out_dist, uint32_t* out_count,
                                          isl_search_stats* stats_or_null);
/* LeannIndex::search_with_params with the attached encoder as the EmbeddingProvider, exactly as the reference runs it
 * (leann.rs:899-988): every hop asks the provider for the embeddings of its unvisited neighbours
 * (compute_embeddings_batch, :947-950) and scores them with the exact metric; no stored vectors and no PQ are involved.
 * The hops of all queries of the batch advance in lockstep, so one encoder pass serves the whole frontier of the batch
 * (docs/leann-specification.md:364-394).  Results equal isl_index_search over an index that stores the encoder's
 * outputs, bit for bit.  Cost: one encoder forward over the batch's frontier per hop. */
isl_status isl_index_search_recompute(const isl_index* idx, const float* queries, uint64_t nq, uint32_t query_dim,
                                      uint32_t k, uint32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                      isl_search_stats* stats_or_null);
/* "PQ ADC traversal + exact rerank" and its recompute form: give an exact distance (and, with recompute, an
 * encoder pass) only to the `limit` survivors with the best table distance instead of all ef — the
 * traversal stays wide, the expensive half shrinks (the role of the rerank ratio `a` of
 * docs/leann-specification.md:223-269 for a traversal that runs entirely on table distances).
 * The effective limit is max(limit, k); 0 (default) reranks every survivor. */
isl_status isl_index_set_rerank_limit(isl_index* idx, uint32_t limit);
/* Hub-embedding cache (docs/leann-specification.md:661-690 `HubCache`): keep the embeddings of the `count`
 * nodes with the highest in-degree (ties: smaller id) resident; the recompute search skips them.  They are
 * computed with the attached encoder, so results are bit-identical with and without the cache.
 * count == 0 drops the cache; attaching another provider drops it too. */
isl_status isl_index_set_hub_cache(isl_index* idx, uint64_t count);
/* Cached nodes, and how many distinct survivors of the last recompute search were served from the cache. */
isl_status isl_index_hub_cache_info(const isl_index* idx, uint64_t* cached_nodes, uint64_t* last_hits);
/* Nodes recomputed by the last recompute search and the CUDA-event time of its three stages. */
isl_status isl_index_last_recompute(const isl_index* idx, uint64_t* unique_nodes, float* traverse_ms,
                                    float* encoder_ms, float* rerank_ms);

/* ---- to_bytes / from_bytes (leann.rs:1059-1066, pq.rs:351-358, hnsw.rs:507-514) -------------- */
/* The reference's `bincode::serialize` layout (bincode 1.x defaults; islands_b200/csrc/bincode.h).
 * out == NULL only reports the length in *out_len.  LeannIndex bytes hold the graph only
 * (leann.rs:1058), so isl_index_from_bytes also takes the embeddings [n][dim] the provider returns.
 * Malformed input -> ISL_SERIALIZATION (CoreError::Deserialization). */
isl_status isl_index_to_bytes(const isl_index* idx, uint8_t* out, uint64_t cap, uint64_t* out_len);
isl_status isl_index_from_bytes(const uint8_t* bytes, uint64_t len, const float* vectors, uint32_t dim,
                                isl_index** out);
isl_status isl_pq_to_bytes(const isl_pq* pq, uint8_t* out, uint64_t cap, uint64_t* out_len);
isl_status isl_pq_from_bytes(const uint8_t* bytes, uint64_t len, isl_pq** out);
isl_status isl_hnsw_to_bytes(const isl_hnsw* g, uint8_t* out, uint64_t cap, uint64_t* out_len);
isl_status isl_hnsw_from_bytes(const uint8_t* bytes, uint64_t len, isl_hnsw** out);
/* The configuration a handle carries (what from_bytes deserialised). */
isl_status isl_index_get_config(const isl_index* idx, isl_leann_config* out);
isl_status isl_pq_get_config(const isl_pq* pq, isl_pq_config* out);
uint32_t isl_pq_dimension(const isl_pq* pq);
isl_status isl_hnsw_get_config(const isl_hnsw* g, isl_hnsw_config* out);
/* One stored vector: HnswNode::vector (hnsw.rs:93-95) / InMemoryEmbeddingProvider::compute_embedding
 * (leann.rs:141-150); out [dimension].  Used by SearchConfig::include_vectors (search.rs:160-164). */
isl_status isl_hnsw_get_vector(const isl_hnsw* g, uint64_t node_id, float* out);
isl_status isl_index_get_vector(const isl_index* idx, uint64_t node_id, float* out);

/* ---- island / shard merge (search.rs:211-237, indexer/service.rs:775-801) ---------- */
/* Per query, merge `parts` lists of k (dist,id) pairs laid out [parts][nq][k] into the k best
 * by (dist, id); ISL_INVALID_ID entries are ignored. */
isl_status isl_merge_topk(const uint64_t* ids, const float* dist, uint32_t parts, uint64_t nq,
                          uint32_t k, uint64_t* out_ids, float* out_dist, uint32_t* out_count);
isl_status isl_merge_topk_dev(const uint64_t* d_ids, const float* d_dist, uint32_t parts,
                              uint64_t nq, uint32_t k, uint64_t* d_out_ids, float* d_out_dist,
                              uint32_t* d_out_count);

/* ... and the merge of `parts` record lists laid out [parts][nq][k] (what follows the exchange). */
isl_status isl_merge_packed_dev(const isl_shard_record* d_records, uint32_t parts, uint64_t nq, uint32_t k,
                                uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_count);

/* ---- sharded search (indexer/service.rs:777-801, search.rs:211-237) ------------------------------------
 * The index is split by node range or by island, one shard per GPU / process.  Every rank calls
 * isl_index_search_sharded with the SAME queries; each searches its shard, the per-shard top-k lists travel as
 * isl_shard_record arrays in ONE exchange and every rank ends with the same merged top-k (by (dist, id), global
 * ids = local id + id_base).  Search kernel, exchange and merge kernel run on one stream without a host
 * synchronisation in between. */
/* ncclGetUniqueId into out (128 bytes): call on one rank, hand the bytes to the others by any host channel. */
isl_status isl_shard_unique_id(void* out, uint64_t cap);
/* ncclCommInitRank on the current CUDA device.  Collective over the `world` ranks. */
isl_status isl_shard_init(int rank, int world, const void* nccl_uid, isl_shard** out);
void isl_shard_free(isl_shard* sh);
int isl_shard_rank(const isl_shard* sh);
int isl_shard_world(const isl_shard* sh);
/* Exchange by peer stores instead of ncclAllGather (ranks of one node): every rank's gather buffer is mapped
 * into the others (CUDA IPC), the search kernel stores each finished query's records straight into all of them
 * over NVLink, and a flag handshake replaces the collective.  Collective call; max_records >= nq * k of any
 * later search (larger searches fall back to NCCL). */
isl_status isl_shard_enable_peer_exchange(isl_shard* sh, uint64_t max_records);
/* Collective: every rank of the communicator must call it, with the same nq / k / queries.  id_base = first
 * global id of this rank's shard.  An empty shard (n == 0) takes part with an empty list. */
isl_status isl_index_search_sharded(const isl_index* idx, isl_shard* sh, uint64_t id_base, const float* queries,
                                    uint64_t nq, uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                    float* out_dist, uint32_t* out_count);
isl_status isl_index_search_sharded_dev(const isl_index* idx, isl_shard* sh, uint64_t id_base, const float* d_queries,
                                        uint64_t nq, uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                                        float* d_out_dist, uint32_t* d_out_count);
/* The same for the PQ searches: mode ISL_SHARD_ADC_RERANK = isl_index_search_adc_rerank on every shard,
 * ISL_SHARD_ADC_RECOMPUTE = isl_index_search_adc_recompute (graph + codes + token rows per shard, encoder per GPU:
 * BASELINE configs[4]); the exact-rerank launch writes the exchange records. */
#define ISL_SHARD_EXACT 0
#define ISL_SHARD_ADC_RERANK 1
#define ISL_SHARD_ADC_RECOMPUTE 2
isl_status isl_index_search_sharded_adc(const isl_index* idx, isl_shard* sh, uint64_t id_base, int32_t mode, const float* queries,
                                        uint64_t nq, uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                        float* out_dist, uint32_t* out_count);
/* CUDA-event durations of the three stages of the last sharded search on this rank. */
isl_status isl_shard_last_timing(const isl_shard* sh, float* search_ms, float* exchange_ms, float* merge_ms);
/* The halves on their own: this shard's result records [nq][k] without an exchange ... */
isl_status isl_index_search_packed_dev(const isl_index* idx, uint64_t id_base, const float* d_queries, uint64_t nq,
                                       uint32_t query_dim, uint32_t k, uint32_t ef, isl_shard_record* d_records);

#ifdef __cplusplus
}
#endif
#endif /* ISLANDS_B200_H */
