/* Compiled as C99 (-pedantic) and run by tests/test_cpp_host.py: include/islands_b200.h is a plain C header — what a
 * cgo / bindgen / ctypes consumer binds — and the library links from C without a C++ runtime on the caller's side.
 * Host-only calls: defaults of the three configurations (leann.rs:386-403, hnsw.rs:37-48, pq.rs:24-34), validation
 * errors with their messages (leann.rs:432-460), the dimension check that precedes any device work
 * (distance.rs:39-44) with its payload, NULL handles. */
#include <stdio.h>
#include <string.h>

#include "islands_b200.h"

#define EXPECT(c)                                             \
  do {                                                        \
    if (!(c)) {                                               \
      printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c);      \
      return 1;                                               \
    }                                                         \
  } while (0)

int main(void) {
  isl_leann_config lc;
  isl_hnsw_config hc;
  isl_pq_config pc;
  uint64_t a = 0, b = 0;
  float out = 0.0f;
  const float x[3] = {1.0f, 2.0f, 3.0f}, y[2] = {1.0f, 2.0f};

  EXPECT(isl_abi_version() >= 2);
  EXPECT(isl_leann_config_default(&lc) == ISL_OK && lc.m == 30 && lc.m0 == 60 && lc.ef_construction == 128);
  EXPECT(lc.metric == ISL_METRIC_COSINE && lc.pruning_strategy == ISL_PRUNE_GLOBAL && lc.high_degree_pruning == 1);
  EXPECT(isl_hnsw_config_default(&hc) == ISL_OK && hc.m == 16 && hc.m0 == 32 && hc.ef_construction == 200);
  EXPECT(isl_pq_config_default(&pc) == ISL_OK && pc.num_subquantizers == 8 && pc.num_centroids == 256 && pc.has_seed == 0);
  EXPECT(isl_pq_config_bytes_per_vector(&pc) == 8);
  lc.m0 = 10; /* M0 < M */
  EXPECT(isl_leann_config_validate(&lc) == ISL_INVALID_CONFIG && strcmp(isl_last_error(), "M0 must be >= M") == 0);
  EXPECT(isl_leann_config_validate(NULL) == ISL_INVALID_ARGUMENT);
  EXPECT(isl_distance_calculate(ISL_METRIC_EUCLIDEAN, x, 3, y, 2, &out) == ISL_DIM_MISMATCH);
  isl_last_error_detail(&a, &b);
  EXPECT(a == 3 && b == 2); /* DimensionMismatch { expected: 3, actual: 2 } */
  EXPECT(isl_index_len(NULL) == 0 && isl_hnsw_len(NULL) == 0);
  isl_index_free(NULL);
  isl_hnsw_free(NULL);
  isl_pq_free(NULL);
  printf("OK\n");
  return 0;
}
