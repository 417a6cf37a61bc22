// kernels.h — launchers of the standalone kernels (distance batch, norms, merge, PQ).
#pragma once

#include "common.cuh"

namespace isl {

// K1: out[i] = metric(query, rows[i]) in reference order (distance.rs:32-34, 37-122).
// rows [n_rows][ld] (ld % 4 == 0, 16B aligned), query [ld] zero padded.
// squared => Distance::calculate_squared (distance.rs:54-66).
isl_status launch_distance_batch(int32_t metric, bool squared, const float* d_query,
                                 const float* d_rows, uint64_t n_rows, uint32_t d, uint32_t ld,
                                 float* d_out, int sms, cudaStream_t st);
// Σ y*y per row, left fold (distance.rs:79) — the cosine denominator of every resident node.
isl_status launch_row_sqnorms(const float* d_rows, uint64_t n_rows, uint32_t d, uint32_t ld,
                              float* d_out, int sms, cudaStream_t st);
// normalize_vector (distance.rs:125-132) per row, in place.
isl_status launch_normalize_rows(float* d_rows, uint64_t n_rows, uint32_t d, uint32_t ld, int sms,
                                 cudaStream_t st);
// Copies [n][d] -> [n][ld] with zero padding (device to device).
isl_status launch_pad_rows(const float* d_src, uint32_t d, float* d_dst, uint32_t ld, uint64_t n,
                           cudaStream_t st);
// u64 -> u32 id narrowing with range check (flag set when any id >= limit).
isl_status launch_narrow_ids(const uint64_t* d_src, uint32_t* d_dst, uint64_t count, uint64_t limit,
                             unsigned int* d_flag, cudaStream_t st);

// CSR -> fixed-stride rows padded with 0xffffffff.
isl_status launch_pad_adjacency(const uint64_t* d_offsets, const uint32_t* d_nbrs, uint64_t n, uint32_t stride,
                                uint32_t* d_out, cudaStream_t st);

// d_deg[i] = offsets[i + 1] - offsets[i] (graph.degree_counts).
isl_status launch_degree_counts(const uint64_t* d_offsets, uint64_t n, uint32_t* d_deg, cudaStream_t st);

// Sets *d_flag (pre-zeroed) to 1 when some neighbour list holds an id twice.
isl_status launch_list_duplicates(const uint64_t* d_offsets, const uint32_t* d_nbrs, uint64_t n, unsigned int* d_flag,
                                  cudaStream_t st);

// K8: per-query merge of [parts][nq][k] lists by (dist,id) (search.rs:211-237).
isl_status launch_merge_topk(const uint64_t* d_ids, const float* d_dist, uint32_t parts, uint64_t nq,
                             uint32_t k, uint64_t* d_out_ids, float* d_out_dist,
                             uint32_t* d_out_count, cudaStream_t st);

// ---- PQ (pq.rs) ------------------------------------------------------------------------
struct PqDev {
  const float* codebooks;  // [m][ksub][ld_sub], rows zero padded to ld_sub (multiple of 4)
  uint32_t m, ksub, dsub, ld_sub;
  int32_t metric;          // metric used by encode (pq.rs:239); tables are always squared L2
};
// K4 encode: codes[i][j] = argmin_c metric(sub_j(v_i), centroid_jc), strict `<` (pq.rs:86-106).
isl_status launch_pq_encode(const PqDev& pq, const float* d_vectors, uint32_t ld, uint64_t n,
                            uint16_t* d_codes, int sms, cudaStream_t st);
// decode: out[i] = concat_j centroid[j][codes[i][j]] (pq.rs:247-271); flag set on code >= ksub.
isl_status launch_pq_decode(const PqDev& pq, const uint16_t* d_codes, uint64_t n, float* d_out,
                            unsigned int* d_flag, cudaStream_t st);
// K3 tables: tables[q][j][c] = Σ_t (q_jt - c_jct)^2 (pq.rs:307-338) for nq queries [nq][q_ld].
isl_status launch_pq_tables(const PqDev& pq, const float* d_queries, uint32_t q_ld, uint64_t nq,
                            float* d_tables, int sms, cudaStream_t st);
// table_distance (pq.rs:341-348) for n code rows against one table set.
isl_status launch_pq_table_distance(const PqDev& pq, const float* d_tables, const uint16_t* d_codes,
                                    uint64_t n, float* d_out, cudaStream_t st);
// asymmetric_distance (pq.rs:275-304) for n code rows against one query.
isl_status launch_pq_asymmetric(const PqDev& pq, const float* d_query, const uint16_t* d_codes,
                                uint64_t n, float* d_out, unsigned int* d_flag, cudaStream_t st);

}  // namespace isl
