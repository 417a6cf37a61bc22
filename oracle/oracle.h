/*
 * oracle.h — CPU restatement of the reference's LEANN/HNSW/PQ hot path (TEST INFRASTRUCTURE).
 *
 * This library is the parity CHECKER.  Only tests/, __graft_entry__.smoke() and the
 * `cpu_baseline` / `--impl reference` legs of bench.py may load it.  The product
 * (libislands_b200.so) never links, loads or calls anything in oracle/.
 *
 * The reference (panbanda/islands v1.5.0) is Rust and cannot be compiled in this image
 * (no cargo/rustc); each function below restates the Rust it cites, with the same
 * operation order: f32, separately rounded multiply and add (no FMA), sums as left folds
 * from 0.0f, IEEE sqrt and divide.  Compile with -ffp-contract=off (see Makefile).
 *
 * Pinning: the reference's own tests hold only a handful of known-answer values for this
 * path (distance.rs:150-229,354-363; pq.rs:505-520,671-677,787-809; leann.rs:1091-1103,
 * 1178-1204; search.rs:311-324); tests/test_oracle_kat.py checks all of them.  Search
 * results, built graphs and PQ codes have no golden vectors in the reference: for those
 * the oracle is "pinned by restatement only".  The two-level (PQ + rerank) search has no
 * reference code at all (docs/leann-specification.md:223-269 is pseudocode): PARITY UNPINNED.
 */
#ifndef ISLANDS_ORACLE_H
#define ISLANDS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/islands_b200.h" /* config / stats struct layouts only */

#ifdef __cplusplus
extern "C" {
#endif

/* distance.rs:37-122 */
float orc_distance(int32_t metric, const float* a, const float* b, uint64_t d);
float orc_distance_squared(int32_t metric, const float* a, const float* b, uint64_t d);
void orc_distance_batch(int32_t metric, const float* q, const float* rows, uint64_t n, uint32_t d,
                        float* out);
/* distance.rs:125-132 */
void orc_normalize(float* v, uint64_t d);

/* leann.rs:549-554 (formula only; the reference's RNG is thread_rng, so u is an input). */
uint64_t orc_level_from_uniform(double u, double ml, uint64_t max_layers);
/* Deterministic level stream used by both oracle and product when no levels are given:
 * u_i = (splitmix64(seed + i) >> 11) * 2^-53 (0 is mapped to 2^-53), level = formula above. */
void orc_draw_levels(uint64_t seed, uint64_t n, double ml, uint64_t max_layers, uint64_t* out);

/* leann.rs:868-988 (+ :991-1056 pruning) on an explicit CSR graph, batched over queries.
 * threads <= 1 -> the reference's sequential map (search.rs:179-181). Returns isl_status. */
int32_t orc_leann_search(const isl_leann_config* cfg, const float* vectors, uint64_t n, uint32_t d,
                         const uint64_t* offsets, const uint64_t* nbrs, int64_t entry,
                         const float* queries, uint64_t nq, uint32_t k, uint32_t ef,
                         uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                         isl_search_stats* stats_or_null, int32_t threads);

/* leann.rs:560-631 with :634-658, :661-749, :761-833.  levels [n] explicit.
 * out_offsets [n+1]; out_nbrs capacity n*m0; returns number of edges via out_num_edges. */
int32_t orc_leann_build(const isl_leann_config* cfg, const float* vectors, uint64_t n, uint32_t d,
                        const uint64_t* levels, uint64_t* out_offsets, uint64_t* out_nbrs,
                        uint64_t* out_num_edges, int64_t* out_entry, uint64_t* out_max_level);

/* Batched construction model used by the GPU build when batch > 1 (no reference analogue;
 * the definition lives in DESIGN.md §build).  Same outputs as orc_leann_build. */
int32_t orc_leann_build_batched(const isl_leann_config* cfg, const float* vectors, uint64_t n,
                                uint32_t d, const uint64_t* levels, uint32_t batch,
                                uint64_t* out_offsets, uint64_t* out_nbrs, uint64_t* out_num_edges,
                                int64_t* out_entry, uint64_t* out_max_level, int32_t threads);

/* pq.rs:86-106, 221-348.  codebooks [m][ksub][dsub]. */
void orc_pq_encode(int32_t metric, const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                   const float* vectors, uint64_t n, uint16_t* out_codes);
int32_t orc_pq_decode(const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                      const uint16_t* codes, uint64_t n, float* out);
void orc_pq_build_tables(const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                         const float* query, float* out_tables);
void orc_pq_table_distance(const float* tables, uint32_t m, uint32_t ksub, const uint16_t* codes,
                           uint64_t n, float* out);
void orc_pq_asymmetric_distance(const float* codebooks, uint32_t m, uint32_t ksub, uint32_t dsub,
                                const float* query, const uint16_t* codes, uint64_t n, float* out);
/* pq.rs:175-218, 362-463 with a splitmix64-based RNG in place of StdRng (ChaCha12 is not
 * restated): statistical parity only.  out_codebooks [m][min(ksub,n)][dsub]. */
int32_t orc_pq_train(int32_t metric, const float* vectors, uint64_t n, uint32_t d, uint32_t m,
                     uint32_t ksub, uint32_t iterations, uint64_t seed, float* out_codebooks,
                     uint32_t* out_ksub);

/* Two-level search as defined in DESIGN.md (docs/leann-specification.md:223-269). */
int32_t orc_leann_search_two_level(const isl_leann_config* cfg, const float* vectors, uint64_t n,
                                   uint32_t d, const uint64_t* offsets, const uint64_t* nbrs,
                                   int64_t entry, const float* codebooks, uint32_t m, uint32_t ksub,
                                   const uint16_t* codes, const float* queries, uint64_t nq,
                                   uint32_t k, uint32_t ef, float rerank_ratio, uint64_t* out_ids,
                                   float* out_dist, uint32_t* out_count,
                                   isl_search_stats* stats_or_null, int32_t threads);

/* The bfloat16 rounding of the ADC traversal's table entries, element by element (so that it can be checked on its own). */
void orc_adc_table_round(const float* in, uint64_t count, float* out);
/* "PQ ADC traversal + exact rerank" (definition: include/islands_b200.h isl_index_search_adc_rerank). */
int32_t orc_leann_search_adc_rerank(const isl_leann_config* cfg, const float* vectors, uint64_t n, uint32_t d,
                                    const uint64_t* offsets, const uint64_t* nbrs, int64_t entry,
                                    const float* codebooks, uint32_t m, uint32_t ksub, const uint16_t* codes,
                                    const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                    float* out_dist, uint32_t* out_count, isl_search_stats* stats_or_null,
                                    int32_t threads,
                                    uint32_t rerank_limit /* 0 = every survivor */);

/* search.rs:211-237 under the (dist,id) rule: lists [parts][nq][k]. */
void orc_merge_topk(const uint64_t* ids, const float* dist, uint32_t parts, uint64_t nq, uint32_t k,
                    uint64_t* out_ids, float* out_dist, uint32_t* out_count);

/* hnsw.rs:214-329, 332-402, 405-446, 458-504: multi-layer HNSW, levels explicit. */
typedef struct orc_hnsw orc_hnsw;
orc_hnsw* orc_hnsw_new(const isl_hnsw_config* cfg, uint32_t d);
void orc_hnsw_free(orc_hnsw* g);
int32_t orc_hnsw_insert(orc_hnsw* g, const float* v, uint64_t level, uint64_t* out_id);
/* Round model of insert for batched construction (see oracle.cpp); batch = 1 == orc_hnsw_insert. */
int32_t orc_hnsw_insert_batch(orc_hnsw* g, const float* vectors, uint64_t count, const uint64_t* levels,
                              uint32_t batch, int32_t threads);
int64_t orc_hnsw_node_level(const orc_hnsw* g, uint64_t id);
uint64_t orc_hnsw_len(const orc_hnsw* g);
int64_t orc_hnsw_entry_point(const orc_hnsw* g);
uint64_t orc_hnsw_max_level(const orc_hnsw* g);
/* neighbours of `id` at `layer`; returns count (or -1 when node/layer is absent). */
int64_t orc_hnsw_neighbors(const orc_hnsw* g, uint64_t id, uint64_t layer, uint64_t* out,
                           uint64_t cap);
int32_t orc_hnsw_search(const orc_hnsw* g, const float* queries, uint64_t nq, uint32_t k,
                        uint32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                        int32_t threads);

/* search.rs:99-102 */
float orc_to_similarity(float score);

/* StdRng (rand 0.8.5) restatement, test hooks: one ChaCha block; a scripted sequence of draws
 * (kind 0 = next_u32, 1 = next_u64, 2 = f32 bits, 3 = choose(bound)). */
void orc_chacha_block(const uint32_t* key8, uint64_t counter, uint64_t stream, int32_t rounds, uint32_t* out16);
void orc_std_rng_draw(uint64_t seed, const uint8_t* kinds, uint64_t count, uint64_t bound, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
