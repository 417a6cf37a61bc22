// search_core.cuh — batched best-first graph search, one query per warp (one-warp CTAs).
//
// Restates the loop of LeannIndex::search_layer_recompute (src/core/leann.rs:899-988) and
// search_layer_with_adjacency (leann.rs:692-749) with data structures chosen for a warp:
//
//   * results R: the reference's max-heap of (dist,id) bounded to ef is kept as an ASCENDING
//     SORTED ARRAY of (dist,id|expanded-bit) — shared memory when it fits, L2-resident global
//     memory otherwise.  The heap and the sorted array hold the same set because admission
//     (`|R| < ef || d < worst.dist`, leann.rs:956-960) and eviction (pop max, :966-968) only
//     depend on the order of the keys.
//   * candidates C: the reference's unbounded min-heap is NOT materialised.  A candidate can
//     still be expanded only while `!(|R| >= ef && d > worst.dist)` (leann.rs:924-928).  Every
//     admitted, unexpanded node with d < worst.dist is in R; admitted nodes that were evicted
//     from R stay expandable only while d == worst.dist — those rare entries are kept in a small
//     `ties` list.  pop_min(C) is therefore "first unexpanded entry of R, else the smallest live
//     tie"; the search ends when neither exists — exactly when the reference pops a stale
//     candidate or drains C.
//   * visited: exact bitset in global memory, one per resident warp; neighbours are marked
//     visited BEFORE pruning (leann.rs:933-937), duplicates within a list keep the first.
//   * exactly ONE candidate is expanded per iteration (multi-pop beams change ids).
//   * distances: dist_pass.cuh (lane per candidate, reference-order fold); admission is then
//     replayed sequentially in CSR order by the whole warp.
#pragma once

#include "row_stream.cuh"
#include "search.h"

namespace isl {

constexpr uint32_t kTieCap = 64;

// one entry of the ADC traversal's table: bfloat16-rounded (common.cuh)
__device__ __forceinline__ float adc_table_entry(float x) { return __uint_as_float(bf16_round_bits(__float_as_uint(x))); }
constexpr uint32_t kExpandedBit = 0x80000000u;

// lean = the ADC-traversal-only kernel (MODE 3): no row staging ring, no query vector.
template <int CH, int STAGES>
__host__ __device__ constexpr size_t search_smem_bytes(uint32_t ld, uint32_t ef_smem, uint32_t u_cap,
                                                       uint32_t lut_floats = 0, uint32_t aq_entries = 0, bool lean = false) {
  return (lean ? 0 : (size_t)STAGES * StageGeom<CH>::STAGE_FLOATS * 4 + (size_t)ld * 4) + (size_t)ef_smem * 8 +
         (size_t)u_cap * 8 + (size_t)kTieCap * 8 + (size_t)STAGES * 8 + (size_t)lut_floats * 4 +
         (size_t)aq_entries * 8;
}
// the lean kernel with R in registers appends the admitted-id cache
__host__ __device__ constexpr size_t search_smem_bytes_idc() { return (size_t)kIdcEntries * 2; }

template <bool R_SMEM>
struct RView {
  uint2* p;
  __device__ __forceinline__ uint2 ld(uint32_t i) const {
    if (R_SMEM) return p[i];
    return __ldcg(p + i);
  }
  __device__ __forceinline__ void st(uint32_t i, uint2 v) const {
    if (R_SMEM)
      p[i] = v;
    else
      __stcg(p + i, v);
  }
};

// leann.rs:991-1016 (Global / Local).  f32 arithmetic as in the reference.
__device__ __forceinline__ uint32_t prune_keep(float prune_ratio, int strategy, uint32_t n_cands,
                                               uint32_t r_len, uint32_t ef) {
  if (prune_ratio == 0.0f || n_cands == 0) return n_cands;
  const float fn = (float)n_cands;
  uint32_t keep;
  if (strategy == ISL_PRUNE_GLOBAL) {
    const float ratio = __fdiv_rn((float)r_len, (float)ef);
    keep = (uint32_t)ceilf(__fmul_rn(fn, __fsub_rn(1.0f, __fmul_rn(ratio, prune_ratio))));
  } else {
    keep = (uint32_t)ceilf(__fmul_rn(fn, __fsub_rn(1.0f, prune_ratio)));
  }
  if (keep < 1) keep = 1;
  return keep < n_cands ? keep : n_cands;
}

// PruningStrategy::Proportional (leann.rs:1017-1053) over list[0 .. n_cands): candidate j is kept when its draw is
// below degree_j / total_degree * num_to_keep; the loop stops once num_to_keep are kept; nothing kept => the first
// candidate.  The draws come from the seeded stream of common.cuh, consumed in list order (`draw_ctr` is the query's
// running draw counter).  The kept ids are compacted to the front of `list`; returns how many.  All lanes call.
__device__ __forceinline__ uint32_t prune_proportional(float prune_ratio, uint32_t* list, uint32_t n_cands, const uint32_t* deg_counts,
                                                       uint64_t seed, uint32_t query, uint64_t* draw_ctr) {
  const uint32_t lane = lane_id();
  uint32_t num_to_keep = (uint32_t)ceilf(__fmul_rn((float)n_cands, __fsub_rn(1.0f, prune_ratio)));  // leann.rs:1001-1003
  if (num_to_keep < 1) num_to_keep = 1;
  unsigned long long total = 0;
  for (uint32_t i = lane; i < n_cands; i += 32) total += __ldg(deg_counts + list[i]);
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (total == 0) return num_to_keep < n_cands ? num_to_keep : n_cands;  // leann.rs:1029-1031
  const float totf = __ull2float_rn(total), keepf = (float)num_to_keep;
  uint32_t w = 0;
  uint64_t consumed = n_cands;
  for (uint32_t b = 0; b < n_cands; b += 32) {
    const uint32_t i = b + lane;
    uint32_t id = 0;
    bool sel = false;
    if (i < n_cands) {
      id = list[i];
      const float prob = __fdiv_rn((float)__ldg(deg_counts + id), totf);
      sel = prune_draw(seed, query, *draw_ctr + i) < __fmul_rn(prob, keepf);
    }
    uint32_t bal = __ballot_sync(0xffffffffu, sel);
    const uint32_t need = num_to_keep - w;
    bool done = false;
    if ((uint32_t)__popc(bal) >= need) {  // the need-th kept candidate ends the reference's loop (selected.len() >= num_to_keep)
      const uint32_t last = __fns(bal, 0, need);
      bal &= last == 31 ? 0xffffffffu : ((1u << (last + 1)) - 1u);
      consumed = (uint64_t)b + last + 1;
      done = true;
    }
    __syncwarp();  // every lane has read its id before the compacted prefix is written (positions <= i)
    if (bal & (1u << lane)) list[w + __popc(bal & ((1u << lane) - 1u))] = id;
    w += __popc(bal);
    __syncwarp();
    if (done) break;
  }
  *draw_ctr += consumed;
  return w ? w : 1u;  // nothing kept: the first candidate, still at list[0] (leann.rs:1049-1051)
}

// ---- R as an unsorted bag in registers (exact traversal, ef <= 1024) ------------------------------------------------
// The reference needs three things from its two heaps (leann.rs:899-988): the worst entry of R (admission `d < worst`,
// eviction), the closest unexpanded entry (the next candidate) and, at the end, the k best in order.  None needs R
// sorted while the search runs: entry i lives in row i / 32 of lane i % 32 and never moves; the worst entry is an argmax
// over the bag (per-lane max over its NR rows, redux.sync, a vote for the owner), paid once per ADMITTED node, which then
// overwrites the evicted slot; the next candidate is an argmin over the unexpanded entries, once per hop; the results are
// k argmin rounds at the end.  A sorted array in shared memory pays a binary search (log2 ef dependent LDS) and a shift
// of ef/2 entries per admission: at ef = 512 that was a third of the kernel's time (0.70 of the HBM peak against 0.87 at
// ef = 104).  Keys: kd = order-preserving image of the f32 distance (NaN folded onto the greatest word: OrderedFloat),
// ki = id << 1 | expanded; an empty slot is kd = 0 (below every real key), ki = 0xffffffff ("expanded").
__device__ __forceinline__ uint32_t dist_key(float d) {
  if (d != d) return 0xffffffffu;
  const uint32_t u = __float_as_uint(d);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_dist(uint32_t k) {
  if (k == 0xffffffffu) return __uint_as_float(0x7fffffffu);
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct XBagPos {
  uint32_t row, lane;  // warp-uniform
};

// argmax of (distance, id) over the bag; called when it holds at least one entry.  Warp-uniform result.
template <int NR>
__device__ __forceinline__ void xbag_argmax(const uint32_t (&kd)[NR], const uint32_t (&ki)[NR], uint32_t* out_kd, uint32_t* out_ki, XBagPos* pos) {
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) m = max(m, kd[j]);
  const uint32_t top = __reduce_max_sync(FULL, m);
  uint32_t hit = 0, sel_ki = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (kd[j] == top) {
      hit |= 1u << j;
      sel_ki = ki[j];
    }
  }
  const uint32_t bal = __ballot_sync(FULL, hit != 0);
  const bool dup_in_lane = __any_sync(FULL, (hit & (hit - 1)) != 0);
  const bool multi = ((bal & (bal - 1)) != 0) | dup_in_lane;
  uint32_t owner = __ffs(bal) - 1;
  uint32_t row = 31 - __clz(hit);
  if (multi) {  // exact distance tie: the greatest id is the worst entry (ki orders like the id)
    bool have = false;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      if (kd[j] == top && (!have || ki[j] > sel_ki)) {
        have = true;
        row = j;
        sel_ki = ki[j];
      }
    }
    const uint32_t top_ki = __reduce_max_sync(FULL, have ? sel_ki : 0u);
    owner = __ffs(__ballot_sync(FULL, have && sel_ki == top_ki)) - 1;
  }
  pos->row = __shfl_sync(FULL, row, owner);
  pos->lane = owner;
  *out_kd = top;
  *out_ki = __shfl_sync(FULL, sel_ki, owner);
}

// argmin of (distance, id) over the unexpanded entries; false when there is none.  Warp-uniform result.
template <int NR>
__device__ __forceinline__ bool xbag_argmin_unexpanded(const uint32_t (&kd)[NR], const uint32_t (&ki)[NR], uint32_t* out_kd, uint32_t* out_id,
                                                       XBagPos* pos) {
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t m = 0xffffffffu, un = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (!(ki[j] & 1u)) {
      un |= 1u << j;
      m = min(m, kd[j]);
    }
  }
  if (!__any_sync(FULL, un != 0)) return false;
  const uint32_t low = __reduce_min_sync(FULL, m);
  uint32_t hit = 0, sel_ki = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (((un >> j) & 1u) && kd[j] == low) {
      hit |= 1u << j;
      sel_ki = ki[j];
    }
  }
  const uint32_t bal = __ballot_sync(FULL, hit != 0);
  const bool dup_in_lane = __any_sync(FULL, (hit & (hit - 1)) != 0);
  const bool multi = ((bal & (bal - 1)) != 0) | dup_in_lane;
  uint32_t owner = __ffs(bal) - 1;
  uint32_t row = 31 - __clz(hit);
  if (multi) {  // exact distance tie: the smallest id first
    bool have = false;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      if (((un >> j) & 1u) && kd[j] == low && (!have || ki[j] < sel_ki)) {
        have = true;
        row = j;
        sel_ki = ki[j];
      }
    }
    const uint32_t low_ki = __reduce_min_sync(FULL, have ? sel_ki : 0xffffffffu);
    owner = __ffs(__ballot_sync(FULL, have && sel_ki == low_ki)) - 1;
  }
  pos->row = __shfl_sync(FULL, row, owner);
  pos->lane = owner;
  *out_kd = low;
  *out_id = __shfl_sync(FULL, sel_ki, owner) >> 1;
  return true;
}

// TWO = two-level search: unvisited neighbours are scored with the PQ table distance
// (pq.rs:341-348) into an approximate queue AQ ordered by (adc,id); after every expansion the
// ceil(a*|AQ|) best entries (at least one) leave AQ, get their exact distance and go through the
// same admission as the exact search.  Since |AQ| <= max_degree / a, AQ is a sorted array too.
//
// MODE 2 (ADC) = "PQ ADC traversal + exact rerank": the whole best-first search runs on the
// table distances (same admission / termination rules with adc in place of the exact distance),
// then the ef surviving candidates get their exact distance and are re-sorted by (dist, id).
// Traversal traffic drops from 4d bytes to m bytes per visited node.
//
// MODE 3 = the traversal half of MODE 2 as its own lean kernel (no staging ring, no query vector:
// 2-3x the resident warps), survivors written out for a MODE 2 / phase 2 rerank launch (the register-bag
// kernel of adc_traverse.cuh serves one-byte codes with ef <= 512; this one the rest).
// NR > 0 (MODE 0 only): R is the unsorted register bag above, NR entries per lane, ef <= 32 * NR.
template <int ACC, int CH, int STAGES, bool R_SMEM, int MODE, int NR = 0>
__global__ void __launch_bounds__(32) leann_search_kernel(const SearchArgs a) {
  constexpr bool TWO = MODE == 1;
  constexpr bool ADC = MODE == 2 || MODE == 3;
  constexpr bool LEAN = MODE == 3;
  constexpr bool RREG = NR > 0;  // R = unsorted bag in registers
  static_assert(!RREG || MODE == 0, "the register bag serves the exact traversal");
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using G = StageGeom<CH>;
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* q_smem = stage + (LEAN ? 0 : STAGES * G::STAGE_FLOATS);
  uint2* r_smem = reinterpret_cast<uint2*>(q_smem + (LEAN ? 0 : a.ld));
  uint32_t* u_list = reinterpret_cast<uint32_t*>(r_smem + ((R_SMEM && !RREG) ? a.ef : 0));
  float* u_nb = reinterpret_cast<float*>(u_list + a.u_cap);  // squared norms of the rows in u_list
  uint2* ties = reinterpret_cast<uint2*>(u_nb + a.u_cap);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ties + kTieCap);
  float* lut_smem = reinterpret_cast<float*>(bars + STAGES);
  uint2* aq_smem = reinterpret_cast<uint2*>(lut_smem + (MODE != 0 ? a.lut_smem_floats : 0));
  uint16_t* idc = reinterpret_cast<uint16_t*>(aq_smem + (TWO ? a.aq_smem_entries : 0));  // only laid out for the lean kernel with R in registers or shared memory
  // Lean ADC traversal without a visited set.  Scoring a node again can never change R: while R is not
  // full every scored node is admitted, afterwards a node that was rejected (d >= worst) or evicted
  // (key above the worst one) is rejected again because the worst distance only decreases, and a node
  // that is still in R is found at its own insertion position (equal key) and dropped.  So the per-query
  // bitset (n / 8 bytes zeroed per query, one DRAM read-modify-write per scored node: two thirds of the
  // kernel's DRAM transactions at 1M nodes, profiles/) is replaced by that duplicate test plus a small
  // direct-mapped cache of admitted ids that filters most repeats before they are replayed.  Ids,
  // survivors, n_hop and n_edge are unchanged; n_adc (distinct nodes scored) is not defined in this
  // mode, which is therefore only used when no statistics are requested.
  const bool novis = LEAN && R_SMEM && a.novis != 0;  // R in registers or in shared memory (the host only sets it then)

  const uint32_t lane = lane_id();
  const uint32_t slot = blockIdx.x;
  uint2* AQ = nullptr;
  if (TWO) AQ = a.aq_smem_entries ? aq_smem : a.aq_global + (size_t)slot * a.aq_cap;
  const bool aq_in_smem = TWO && a.aq_smem_entries != 0;
  auto aq_ld = [&](uint32_t i) -> uint2 { return aq_in_smem ? AQ[i] : __ldcg(AQ + i); };
  auto aq_st = [&](uint32_t i, uint2 v) {
    if (aq_in_smem)
      AQ[i] = v;
    else
      __stcg(AQ + i, v);
  };
  uint32_t* vis = a.visited + (size_t)slot * a.vis_words;
  RView<R_SMEM> R{R_SMEM ? r_smem : a.r_global + (size_t)slot * a.ef};
  const uint32_t ef = a.ef;
  // ties: the first kTieCap entries live in shared memory, the rest spill to a per-slot global array of ef
  // entries.  Live ties all carry the current worst distance and were in R together when the first of them
  // was evicted, so there are never more than ef - 1 of them: kTieCap + ef entries cannot overflow once the
  // stale ones are dropped (the reference's unbounded heap, leann.rs:924-928, needs no more either).
  uint2* ties_spill = a.ties_global + (size_t)slot * ef;
  auto tie_ld = [&](uint32_t i) -> uint2 { return i < kTieCap ? ties[i] : __ldcg(ties_spill + (i - kTieCap)); };
  auto tie_st = [&](uint32_t i, uint2 v) {
    if (i < kTieCap)
      ties[i] = v;
    else
      __stcg(ties_spill + (i - kTieCap), v);
  };
  // register bag (RREG): entry i = row i / 32 of lane i % 32
  uint32_t kd[RREG ? NR : 1];
  uint32_t ki[RREG ? NR : 1];

  RowRing<STAGES> ring;
  ring.stage = stage;
  ring.bars = bars;
  ring.phase_bits = 0;
  ring.islot = 0;
  ring.cslot = 0;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    mbar_fence_init();
  }
  __syncwarp();

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(a.work_counter, 1u);
    qi = __shfl_sync(0xffffffffu, qi, 0);
    if (qi >= a.nq) break;

    // ---- per-query setup -------------------------------------------------------------
    {
      // a query is either row qi of the query matrix or (construction) a row of the vector table
      if (!LEAN) {
        const float4* src = reinterpret_cast<const float4*>(
            a.query_ids ? a.vectors + (size_t)__ldg(a.query_ids + qi) * a.ld : a.queries + (size_t)qi * a.q_ld);
        float4* dst = reinterpret_cast<float4*>(q_smem);
        for (uint32_t i = lane; i < a.ld / 4; i += 32) dst[i] = src[i];
      }
      if (novis) {
        uint32_t* c2 = reinterpret_cast<uint32_t*>(idc);
        for (uint32_t i = lane; i < kIdcEntries / 2; i += 32) c2[i] = 0xffffffffu;
      } else if (!(ADC && a.phase == 2)) {  // the rerank-only launch never touches the visited set
        uint4* v4 = reinterpret_cast<uint4*>(vis);
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (uint32_t i = lane; i < a.vis_words / 4; i += 32) __stcg(v4 + i, z);
      }
    }
    __threadfence();
    __syncwarp();
    const float na = (!LEAN && a.metric == ISL_METRIC_COSINE) ? smem_sqnorm_fold(q_smem, a.d) : 0.0f;

    uint32_t r_len = 0, first_unexp = 0, n_ties = 0, aq_len = 0;
    // capacity of R for the shared / global array: ef during a traversal; k for the exact rerank of the ADC survivors,
    // which only has to deliver the k best by (distance, id) — see the rerank below
    uint32_t cap = ef;
    bool rerank_topk = false;
    uint32_t tie_next = kTieCap;  // the tie list is compacted (stale entries dropped) when it reaches this length
    float wst_d = 0.0f;    // register bag: the worst entry once R is full (distance, key words, position)
    uint32_t wst_kd = 0xffffffffu, wst_ki = 0xffffffffu;
    XBagPos wst_pos{0, 0};
    if constexpr (RREG) {
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        kd[j] = 0u;
        ki[j] = 0xffffffffu;
      }
    }
    uint64_t n_hop = 0, n_edge = 0, n_dist = 0, n_adc = 0, n_rerank = 0;
    uint64_t draw_ctr = 0;  // Proportional pruning: draws consumed by this query so far
    const float* lut = nullptr;
    if (MODE != 0 && !(ADC && a.phase == 2)) {
      if (a.luts == nullptr) {
        // build_distance_tables (pq.rs:307-338) for this query straight into shared memory:
        // LUT[j][c] = sum_t (q_jt - c_jct)^2, left fold, one centroid per lane.  ~200k fold steps per
        // query against the L2-resident codebooks: a few percent of a traversal, and it saves the
        // separate tables kernel plus 32 KB of table write + read per query.
        const float* qv = a.queries + (size_t)qi * a.q_ld;
        const uint32_t nvec = a.pq_ld_sub >> 2;
        constexpr int CC = 4;  // centroids per lane in flight: independent fold chains hide the latency
        for (uint32_t j = 0; j < a.pq_m; ++j) {
          const float* qs = qv + (size_t)j * a.pq_dsub;
          for (uint32_t c0 = 0; c0 < a.pq_ksub; c0 += 32 * CC) {
            float acc[CC];
            const float4* row[CC];
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
              acc[cc] = 0.0f;
              const uint32_t c = min(c0 + cc * 32 + lane, a.pq_ksub - 1);
              row[cc] = reinterpret_cast<const float4*>(a.pq_codebooks + ((size_t)j * a.pq_ksub + c) * a.pq_ld_sub);
            }
            for (uint32_t v = 0; v < nvec; ++v) {
              float4 y[CC];
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) y[cc] = __ldg(row[cc] + v);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t t = v * 4 + e;
                if (t < a.pq_dsub) {
                  const float qe = __ldg(qs + t);
#pragma unroll
                  for (int cc = 0; cc < CC; ++cc) {
                    const float ye = e == 0 ? y[cc].x : (e == 1 ? y[cc].y : (e == 2 ? y[cc].z : y[cc].w));
                    const float diff = __fsub_rn(qe, ye);
                    acc[cc] = __fadd_rn(acc[cc], __fmul_rn(diff, diff));
                  }
                }
              }
            }
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
              const uint32_t c = c0 + cc * 32 + lane;
              if (c < a.pq_ksub) lut_smem[j * a.pq_ksub + c] = ADC ? adc_table_entry(acc[cc]) : acc[cc];
            }
          }
        }
        __syncwarp();
        lut = lut_smem;
      } else {
        const float* g = a.luts + (size_t)qi * a.pq_m * a.pq_ksub;
        if (a.lut_smem_floats) {
          for (uint32_t i = lane; i < a.lut_smem_floats; i += 32) lut_smem[i] = ADC ? adc_table_entry(__ldg(g + i)) : __ldg(g + i);
          __syncwarp();
          lut = lut_smem;
        } else {
          lut = g;
        }
      }
    }

    // ---- R access ---------------------------------------------------------------------------
    auto r_get = [&](uint32_t i) __attribute__((always_inline)) -> uint2 { return R.ld(i); };  // sorted array only
    // ---- insert into R -------------------------------------------------------------------------
    // shared/global R: sorted array — binary search, warp shift, store.  register bag: see above.
    auto r_insert = [&](float dnew, uint32_t idnew) __attribute__((always_inline)) {
      uint32_t pos;
      const bool full = RREG ? (r_len == ef) : (r_len == cap);
      uint2 evicted = make_uint2(0, 0);
      if constexpr (RREG) {
        // bag: the new entry goes into the next free slot, or takes the slot of the worst entry (pop max,
        // leann.rs:966-968); the new worst entry is an argmax over the bag
        const uint32_t nkd = dist_key(dnew), nki = idnew << 1;
        if (full) evicted = make_uint2(__float_as_uint(wst_d), (wst_ki >> 1) | ((wst_ki & 1u) ? kExpandedBit : 0u));
        const uint32_t srow = full ? wst_pos.row : (r_len >> 5);
        const bool me = lane == (full ? wst_pos.lane : (r_len & 31));
#pragma unroll
        for (int j = 0; j < NR; ++j)
          if ((uint32_t)j == srow && me) {
            kd[j] = nkd;
            ki[j] = nki;
          }
        if (full || r_len + 1 == ef) {
          xbag_argmax<NR>(kd, ki, &wst_kd, &wst_ki, &wst_pos);
          wst_d = key_dist(wst_kd);
        }
        pos = 0xffffffffu;  // no position in a bag: first_unexp is not used
      } else {
        uint32_t lo = 0, hi = r_len;  // first position whose key is not < new key
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          const uint2 e = R.ld(mid);
          if (key_lt(__uint_as_float(e.x), e.y & ~kExpandedBit, dnew, idnew))
            lo = mid + 1;
          else
            hi = mid;
        }
        pos = lo;
        if (novis) {  // a node that is already in R has an equal key: it sits exactly at the lower bound
          if (pos < r_len) {
            const uint2 e = R.ld(pos);
            if ((e.y & ~kExpandedBit) == idnew) return;  // scored twice, admitted once
          }
          if (lane == 0) idc[idnew & (kIdcEntries - 1)] = (uint16_t)(idnew >> kIdcBits);
        }
        if (full) evicted = R.ld(cap - 1);
        const int top = full ? (int)cap - 1 : (int)r_len;
        for (int t = top; t > (int)pos; t -= 32) {
          const int i = t - (int)lane;
          const bool act = i > (int)pos;
          uint2 e = make_uint2(0, 0);
          if (act) e = R.ld(i - 1);
          __syncwarp();
          if (act) R.st(i, e);
          __syncwarp();
        }
        if (lane == 0) R.st(pos, make_uint2(__float_as_uint(dnew), idnew));
        __syncwarp();
      }
      if (!full) r_len++;
      if (pos <= first_unexp) first_unexp = pos;
      if (full && !rerank_topk && !(evicted.y & kExpandedBit)) {
        // An evicted, unexpanded node stays expandable while its distance equals the worst
        // distance in R (leann.rs:924-928 uses a strict `>`).
        const float wd = RREG ? wst_d : __uint_as_float(R.ld(cap - 1).x), edist = __uint_as_float(evicted.x);
        if (!of_lt(wd, edist)) {
          if (n_ties == tie_next) {  // drop stale ties first (lane 0 moves the survivors down, in order)
            uint32_t kept = 0;
            for (uint32_t i = 0; i < n_ties; ++i) {
              const uint2 t = tie_ld(i);
              if (!of_lt(wd, __uint_as_float(t.x))) {
                __syncwarp();
                if (lane == 0) tie_st(kept, t);
                kept++;
              }
            }
            n_ties = kept;
            tie_next = min(kTieCap + ef, max(kTieCap, 2 * kept));
            __syncwarp();
          }
          if (n_ties >= kTieCap + ef) {  // cannot happen (see above); kept as a loud guard
            if (lane == 0) atomicExch(a.error_flag, 1u);
          } else {
            if (lane == 0) tie_st(n_ties, evicted);
            n_ties++;
            __syncwarp();
          }
        }
      }
    };

    // ---- exact distances of u_list[0 .. total) and their sequential admission ----------------
    // Streaming and fold: row_stream.cuh.  Admission replays the reference's per-neighbour loop
    // (leann.rs:953-970) in list order over the lanes that can still be admitted.
    auto admit_values = [&](bool ok, float dn, uint32_t cid) __attribute__((always_inline)) {
      if constexpr (!RREG) {
        if (rerank_topk) {
          // exact rerank: R keeps the k best by the FULL (distance, id) key — what sorting all survivors and taking k yields
          uint2 w = make_uint2(0, 0);
          if (r_len > 0) w = R.ld(r_len - 1);
          uint32_t mask = __ballot_sync(0xffffffffu, ok && (r_len < cap || key_lt(dn, cid, __uint_as_float(w.x), w.y & ~kExpandedBit)));
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float dj = __shfl_sync(0xffffffffu, dn, j);
            const uint32_t idj = __shfl_sync(0xffffffffu, cid, j);
            bool add = r_len < cap;
            if (!add) {
              const uint2 e = R.ld(cap - 1);
              add = key_lt(dj, idj, __uint_as_float(e.x), e.y & ~kExpandedBit);
            }
            if (add) r_insert(dj, idj);
          }
          return;
        }
      }
      float worst = 0.0f;
      if (r_len > 0) worst = RREG ? wst_d : __uint_as_float(R.ld(r_len - 1).x);
      uint32_t mask = __ballot_sync(0xffffffffu, ok && (r_len < ef || dn < worst));
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const float dj = __shfl_sync(0xffffffffu, dn, j);
        const uint32_t idj = __shfl_sync(0xffffffffu, cid, j);
        bool add = r_len < ef;
        if (!add) add = dj < (RREG ? wst_d : __uint_as_float(R.ld(ef - 1).x));  // raw f32 `<` (leann.rs:959)
        if (add) r_insert(dj, idj);
      }
    };
    auto admit_group = [&](uint32_t base, uint32_t cnt, float acc) {
      float dn = 0.0f;
      uint32_t cid = 0;
      cp_async_wait<0>();  // the squared norms requested before the stream started
      __syncwarp();
      if (lane < cnt) {
        cid = u_list[base + lane];
        const float nb = (a.metric == ISL_METRIC_COSINE) ? u_nb[base + lane] : 0.0f;
        dn = finalize_distance(a.metric, acc, na, nb);
      }
      admit_values(lane < cnt, dn, cid);
    };
    // table_distance (pq.rs:341-348) of node `nid`: left fold over the subquantizers, then sqrt
    auto adc_of = [&](uint32_t nid) -> float {
      float sacc = 0.0f;
      // the ADC traversal folds bfloat16-rounded table entries (common.cuh; rounding is idempotent, so entries that were
      // rounded when the table was staged pass through unchanged); the two-level search keeps the f32 table
      if (a.codes8) {
        const uint8_t* cd = a.codes8 + (size_t)nid * a.pq_m;
        for (uint32_t j = 0; j < a.pq_m; ++j) {
          const float e = lut[j * a.pq_ksub + cd[j]];
          sacc = __fadd_rn(sacc, ADC ? adc_table_entry(e) : e);
        }
      } else {
        const uint16_t* cd = a.codes16 + (size_t)nid * a.pq_m;
        for (uint32_t j = 0; j < a.pq_m; ++j) {
          const float e = lut[j * a.pq_ksub + cd[j]];
          sacc = __fadd_rn(sacc, ADC ? adc_table_entry(e) : e);
        }
      }
      return __fsqrt_rn(sacc);
    };
    auto score_and_admit = [&](uint32_t total) {
      if (a.metric == ISL_METRIC_COSINE) {  // 4-byte async gathers: land long before the group is admitted
        for (uint32_t i = lane; i < total; i += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(u_nb + i)),
                       "l"(a.sqnorms + (a.row_of_id ? __ldg(a.row_of_id + u_list[i]) : u_list[i])));
        cp_async_commit();
      }
      stream_rows_fold<ACC, STAGES, CH>(ring, a.vectors, a.ld, a.d, u_list, total, q_smem, admit_group, a.row_of_id);
    };

    const bool traverse = !(ADC && a.phase == 2);
    // ---- entry point (leann.rs:911-916; per query for the HNSW layer searches, hnsw.rs:501) ----
    const uint32_t entry = a.entries ? __ldg(a.entries + (a.query_ids ? __ldg(a.query_ids + qi) : qi)) : a.entry;
    if (traverse) {
      if (lane == 0) {
        u_list[0] = entry;
        if (!novis) atomicOr(vis + (entry >> 5), 1u << (entry & 31));
      }
      __syncwarp();
      if (ADC) {
        const float d0 = adc_of(entry);
        admit_values(lane == 0, d0, entry);
        n_adc = 1;
      } else {
        score_and_admit(1);
        n_dist = 1;
      }
    }

    // ---- main loop (leann.rs:922-972) ------------------------------------------------------
    for (; traverse;) {
      uint32_t cur = 0;
      bool have_cur = false;
      if constexpr (RREG) {
        // the closest unexpanded entry of the bag; mark it expanded
        uint32_t ckd;
        XBagPos cp;
        if (xbag_argmin_unexpanded<NR>(kd, ki, &ckd, &cur, &cp)) {
          have_cur = true;
          const bool me = lane == cp.lane;
#pragma unroll
          for (int j = 0; j < NR; ++j)
            if ((uint32_t)j == cp.row && me) ki[j] |= 1u;
          // the cached copy of the worst entry must see the flag too: an expanded worst entry that is evicted
          // later with a distance equal to the new worst one must not come back through the ties list
          if (r_len == ef && cp.row == wst_pos.row && cp.lane == wst_pos.lane) wst_ki |= 1u;
        }
      } else if (first_unexp < r_len) {
        have_cur = true;
        const uint2 e = r_get(first_unexp);
        cur = e.y;
        uint32_t nxt = r_len;
        __syncwarp();
        if (lane == 0) R.st(first_unexp, make_uint2(e.x, e.y | kExpandedBit));
        __syncwarp();
        // advance to the next unexpanded entry
        for (uint32_t b = first_unexp + 1; b < r_len; b += 32) {
          const uint32_t i = b + lane;
          const bool un = i < r_len && !(R.ld(i).y & kExpandedBit);
          const uint32_t bal = __ballot_sync(0xffffffffu, un);
          if (bal) {
            nxt = b + __ffs(bal) - 1;
            break;
          }
        }
        first_unexp = nxt;
      }
      if (!have_cur) {
        // smallest live tie, if any.  Every lane scans the whole list (broadcast loads): the result is then provably
        // warp-uniform.  A lane-strided scan + shuffle reduction hands `cur` back through shuffles, which the compiler
        // must treat as divergent — it then wraps every collective of the search loop in convergence barriers
        // (BSSY/BSYNC 29 -> 370 in the SASS of the ADC traversal kernel, 63 -> 80 registers, -13 % speed).
        const float wd = RREG ? wst_d : __uint_as_float(R.ld(r_len - 1).x);
        int best = -1;
        for (uint32_t i = 0; i < n_ties; ++i) {
          const uint2 t = tie_ld(i);
          if (r_len >= ef && of_lt(wd, __uint_as_float(t.x))) continue;  // stale
          if (best < 0 || key_lt(__uint_as_float(t.x), t.y, __uint_as_float(tie_ld(best).x), tie_ld(best).y))
            best = (int)i;
        }
        if (best < 0) break;
        cur = tie_ld(best).y;
        __syncwarp();
        if (lane == 0) tie_st(best, tie_ld(n_ties - 1));
        n_ties--;
        __syncwarp();
      }

      // neighbour list of `cur`
      // Three layouts: CSR (offsets), fixed-stride rows with a live degree array (graph under
      // construction), fixed-stride rows padded with 0xffffffff (finished index: one dependent
      // load less per hop, and the row address of the NEXT candidate is known early enough to
      // be prefetched into L2).
      uint64_t start;
      uint32_t deg;
      bool sentinel = false;
      if (a.offsets) {
        start = __ldg(a.offsets + cur);
        deg = (uint32_t)(__ldg(a.offsets + cur + 1) - start);
      } else {
        // upper HNSW layers keep their rows in a compact pool: row = row_map[node] + row_add
        const bool has_layer = !a.node_levels || __ldg(a.node_levels + cur) >= a.layer;
        const uint32_t row = (a.row_map && has_layer) ? __ldg(a.row_map + cur) + a.row_add : cur;
        start = (uint64_t)row * a.adj_stride;
        if (!has_layer) {
          deg = 0;  // neighbors_at(layer) is None above the node's own level (hnsw.rs:108-110, :362)
        } else if (a.degrees) {
          deg = __ldg(a.degrees + row);
        } else {
          deg = a.adj_stride;
          sentinel = true;
          // the list of the candidate that is next as things stand goes to L2 while this one is scored
          uint32_t nxt = 0;
          bool have_nxt = false;
          if constexpr (RREG) {
            uint32_t nkd;
            XBagPos np;
            have_nxt = xbag_argmin_unexpanded<NR>(kd, ki, &nkd, &nxt, &np);
          } else if (first_unexp < r_len) {
            have_nxt = true;
            nxt = r_get(first_unexp).y & ~kExpandedBit;
          }
          if (have_nxt && lane * 32 < a.adj_stride) {
            const uint32_t* pf = a.nbrs + (size_t)nxt * a.adj_stride + lane * 32;
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(pf));
          }
        }
      }
      n_hop++;
      if (!sentinel) n_edge += deg;

      if (ADC && a.codes8 && (a.pq_m == 16 || a.pq_m == 32) && a.lut_smem_floats) {
        // ADC hop with every latency in flight at once: 64 list positions per pass (two per lane);
        // their ids are loaded, then the visited-bit atomics AND the code rows of all positions are
        // issued together (codes of already visited nodes are wasted bytes, ~15 %, but the test and
        // the gather no longer wait for each other), then table distances, then admission in list
        // order of the positions whose bit was clear.
        const uint32_t nv = a.pq_m >> 4;  // 16-byte pieces per code row
        if (novis) __syncwarp();          // lane 0's writes to the admitted-id cache are read by every lane below
        for (uint32_t b = 0; b < deg; b += 64) {
          uint32_t nid[2];
          bool valid[2], chk[2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t i = b + r * 32 + lane;
            nid[r] = i < deg ? __ldg(a.nbrs + start + i) : 0xffffffffu;
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            valid[r] = nid[r] != 0xffffffffu;
            if (sentinel) n_edge += __popc(__ballot_sync(FULL, valid[r]));
            chk[r] = valid[r] && nid[r] < a.n;
            if (!a.lists_unique && !novis) {  // first of its value in this half
              const uint32_t same = __match_any_sync(FULL, nid[r]);
              chk[r] = chk[r] && lane == (uint32_t)(__ffs(same) - 1);
            }
          }
          if (!a.lists_unique && !novis && b + 32 < deg) {  // a value of the second half that already occurs in the first is not a first occurrence
            for (uint32_t t = 0; t < 32; ++t) {
              const uint32_t v0 = __shfl_sync(FULL, nid[0], t);
              if (nid[1] == v0) chk[1] = false;
            }
          }
          uint32_t old[2] = {0xffffffffu, 0xffffffffu};
          uint4 cw[2][2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (novis) {  // an id found in the admitted-id cache was admitted before: never again
              if (chk[r] && idc[nid[r] & (kIdcEntries - 1)] == (uint16_t)(nid[r] >> kIdcBits)) chk[r] = false;
              old[r] = 0;
            } else if (chk[r]) {
              old[r] = atomicOr(vis + (nid[r] >> 5), 1u << (nid[r] & 31));
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              cw[r][v] = make_uint4(0, 0, 0, 0);
              if (chk[r] && (uint32_t)v < nv) cw[r][v] = __ldg(reinterpret_cast<const uint4*>(a.codes8 + (size_t)nid[r] * a.pq_m) + v);
            }
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float sacc = 0.0f;  // table_distance (pq.rs:341-348): left fold over the subquantizers
            const float* lj = lut_smem;  // the staged table: addressed as shared memory (LDS), one row per subquantizer
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              if ((uint32_t)v < nv) {
                const uint32_t w[4] = {cw[r][v].x, cw[r][v].y, cw[r][v].z, cw[r][v].w};
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const uint32_t code = __byte_perm(w[k >> 2], 0, 0x4440 + (k & 3));
                  sacc = __fadd_rn(sacc, lj[code]);
                  lj += a.pq_ksub;
                }
              }
            }
            const float adc = __fsqrt_rn(sacc);
            const bool unv = chk[r] && !(old[r] & (1u << (nid[r] & 31)));
            n_adc += __popc(__ballot_sync(FULL, unv));
            admit_values(unv, adc, nid[r]);
          }
        }
        continue;
      }

      // unvisited neighbours, in list order (leann.rs:933-937)
      uint32_t ucnt = 0;
      // Lists without repeated ids (a finished index, scanned once): 64 positions per pass, two per lane, so that the
      // visited-bit atomics of both halves — DRAM read-modify-writes on big shards — are in flight together instead of
      // one dependent round trip per 32 positions.  (With repeated ids the second half must see the bits the first
      // half set, so that the FIRST occurrence is the one that is kept: the sequential loop below.)
      for (uint32_t b = 0; a.lists_unique && b < deg; b += 64) {
        uint32_t nid2[2], old2[2];
        bool ok2[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t i = b + r * 32 + lane;
          nid2[r] = i < deg ? __ldg(a.nbrs + start + i) : 0xffffffffu;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const bool valid = nid2[r] != 0xffffffffu;  // past the degree, or the padding of a fixed-stride row
          if (sentinel) n_edge += __popc(__ballot_sync(0xffffffffu, valid));
          ok2[r] = valid && nid2[r] < a.n;
          old2[r] = 0xffffffffu;
          if (ok2[r]) old2[r] = atomicOr(vis + (nid2[r] >> 5), 1u << (nid2[r] & 31));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const bool unv = ok2[r] && !(old2[r] & (1u << (nid2[r] & 31)));
          const uint32_t bal = __ballot_sync(0xffffffffu, unv);
          if (unv) u_list[ucnt + __popc(bal & ((1u << lane) - 1))] = nid2[r];
          ucnt += __popc(bal);
        }
      }
      for (uint32_t b = 0; !a.lists_unique && b < deg; b += 32) {
        const uint32_t i = b + lane;
        bool valid = i < deg;
        uint32_t nid = 0xffffffffu;
        if (valid) nid = __ldg(a.nbrs + start + i);
        if (sentinel) {
          valid = valid && nid != 0xffffffffu;
          n_edge += __popc(__ballot_sync(0xffffffffu, valid));
        }
        const uint32_t same = __match_any_sync(0xffffffffu, nid);
        const bool first = lane == (uint32_t)(__ffs(same) - 1);
        bool unv = false;
        if (valid && first && nid < a.n) {
          const uint32_t bit = 1u << (nid & 31);
          const uint32_t old = atomicOr(vis + (nid >> 5), bit);
          unv = !(old & bit);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, unv);
        if (unv) u_list[ucnt + __popc(bal & ((1u << lane) - 1))] = nid;
        ucnt += __popc(bal);
      }
      __syncwarp();
      if (ucnt == 0 && (!TWO || aq_len == 0)) continue;  // leann.rs:939-941

      if (ADC) {
        // traversal on table distances only: one lane per unvisited neighbour, list order
        n_adc += ucnt;
        for (uint32_t b = 0; b < ucnt; b += 32) {
          const uint32_t i = b + lane;
          float adc = 0.0f;
          uint32_t nid = 0;
          if (i < ucnt) {
            nid = u_list[i];
            adc = adc_of(nid);
          }
          admit_values(i < ucnt, adc, nid);
        }
        continue;
      }
      if (!TWO) {
        const uint32_t keep = (a.strategy == ISL_PRUNE_PROPORTIONAL && a.prune_ratio != 0.0f)
                                  ? prune_proportional(a.prune_ratio, u_list, ucnt, a.deg_counts, a.prune_seed, qi, &draw_ctr)
                                  : prune_keep(a.prune_ratio, a.strategy, ucnt, r_len, ef);  // :944
        n_dist += keep;
        score_and_admit(keep);
        continue;
      }

      // ---- two-level: ADC for the frontier, then promote the best of AQ ---------------------
      n_adc += ucnt;
      for (uint32_t b = 0; b < ucnt; b += 32) {
        const uint32_t i = b + lane;
        float adc = 0.0f;
        uint32_t nid = 0;
        if (i < ucnt) {
          nid = u_list[i];
          adc = adc_of(nid);
        }
        const uint32_t cntb = min(32u, ucnt - b);
        for (uint32_t t = 0; t < cntb; ++t) {  // sorted insert of each (adc,id) into AQ
          const float dnew = __shfl_sync(0xffffffffu, adc, t);
          const uint32_t idnew = __shfl_sync(0xffffffffu, nid, t);
          uint32_t lo = 0, hi = aq_len;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint2 e = aq_ld(mid);
            if (key_lt(__uint_as_float(e.x), e.y, dnew, idnew))
              lo = mid + 1;
            else
              hi = mid;
          }
          for (int tt = (int)aq_len; tt > (int)lo; tt -= 32) {
            const int ii = tt - (int)lane;
            const bool act = ii > (int)lo;
            uint2 e = make_uint2(0, 0);
            if (act) e = aq_ld(ii - 1);
            __syncwarp();
            if (act) aq_st(ii, e);
            __syncwarp();
          }
          if (lane == 0) aq_st(lo, make_uint2(__float_as_uint(dnew), idnew));
          aq_len++;
          __syncwarp();
        }
      }
      uint32_t promote = (uint32_t)ceilf(__fmul_rn((float)aq_len, a.rerank_ratio));
      if (promote < 1) promote = 1;
      if (promote > aq_len) promote = aq_len;
      for (uint32_t i = lane; i < promote; i += 32) u_list[i] = aq_ld(i).y;
      __syncwarp();
      for (uint32_t b = 0; b + promote < aq_len; b += 32) {  // close the gap left by the promoted prefix
        const uint32_t i = b + lane;
        uint2 e = make_uint2(0, 0);
        const bool act = i + promote < aq_len;
        if (act) e = aq_ld(i + promote);
        __syncwarp();
        if (act) aq_st(i, e);
        __syncwarp();
      }
      aq_len -= promote;
      n_dist += promote;
      n_rerank += promote;
      score_and_admit(promote);
    }

    if (ADC && (LEAN || a.phase == 1)) {
      // traversal-only launch: hand the ef survivors (ascending adc order) to the rerank / recompute step;
      // with a rerank limit only the best of them (isl_index_set_rerank_limit)
      const uint32_t n_surv = a.rerank_limit ? min(r_len, a.rerank_limit) : r_len;
      for (uint32_t i = lane; i < n_surv; i += 32) a.surv_ids[(size_t)qi * ef + i] = R.ld(i).y & ~kExpandedBit;
      if (lane == 0) {
        a.surv_cnt[qi] = n_surv;
        if (a.stats) {
          isl_search_stats s;
          s.n_hop = n_hop;
          s.n_edge = n_edge;
          s.n_dist = 0;
          s.n_adc = n_adc;
          s.n_rerank = 0;
          a.stats[qi] = s;
        }
      }
      __syncwarp();
      continue;
    }
    if (ADC) {
      // exact rerank of the ef survivors: their ids move to u_list, R is rebuilt from the exact
      // distances (capacity ef >= their number, so every one is admitted) and ends up sorted by
      // (dist, id).
      uint32_t total = r_len;
      if (a.phase == 2) {
        total = __ldg(a.surv_cnt + qi);
        for (uint32_t i = lane; i < total; i += 32) u_list[i] = __ldg(a.surv_ids + (size_t)qi * ef + i);
      } else {
        for (uint32_t i = lane; i < total; i += 32) u_list[i] = R.ld(i).y & ~kExpandedBit;
      }
      __syncwarp();
      r_len = 0;
      first_unexp = 0;
      n_ties = 0;
      tie_next = kTieCap;
      n_dist = total;
      n_rerank = total;
      // Sorting all survivors by (distance, id) and taking k == keeping a sorted array of the k best by that key: a
      // survivor only enters R when its key is below the k-th best so far (a handful of inserts per group instead of one
      // sorted insert into ef entries per survivor).
      cap = a.k < ef ? a.k : ef;
      rerank_topk = true;
      score_and_admit(total);
    }

    // ---- results: R is already sorted by (dist,id); take(k) (leann.rs:895) -----------------
    const uint32_t cnt = r_len < a.k ? r_len : a.k;
    if constexpr (RREG) {
      // the k best of the bag in (distance, id) order: k argmin rounds over the entries not yet taken (every flag is
      // cleared first; a taken entry is flagged again); lane i % 32 keeps result i until its row of 32 is complete
#pragma unroll
      for (int j = 0; j < NR; ++j)
        if ((uint32_t)(j * 32) + lane < r_len) ki[j] &= ~1u;
    }
    for (uint32_t i0 = 0; i0 < a.k; i0 += 32) {
      const uint32_t i = i0 + lane;
      uint32_t id = 0xffffffffu;
      float dist = __int_as_float(0x7f800000);
      if constexpr (RREG) {
        for (uint32_t t = i0; t < min(i0 + 32, cnt); ++t) {
          uint32_t tkd, tid;
          XBagPos tp;
          xbag_argmin_unexpanded<NR>(kd, ki, &tkd, &tid, &tp);  // t < r_len: there is one
          const bool me = lane == tp.lane;
#pragma unroll
          for (int j = 0; j < NR; ++j)
            if ((uint32_t)j == tp.row && me) ki[j] |= 1u;
          if (lane == (t & 31)) {
            id = tid;
            dist = key_dist(tkd);
          }
        }
      } else if (i < cnt) {
        const uint2 e = R.ld(i);
        id = e.y & ~kExpandedBit;
        dist = __uint_as_float(e.x);
      }
      if (i >= a.k) continue;
      const size_t o = (size_t)qi * a.k + i;
      if (a.out_ids) a.out_ids[o] = i < cnt ? (uint64_t)id : ISL_INVALID_ID;
      if (a.out_ids32) a.out_ids32[o] = id;
      if (a.out_dist) a.out_dist[o] = dist;
      if (a.shard.packed || a.shard.n_peer) {  // shard exchange record: {dist bits, 0, global id lo, global id hi}
        const uint64_t gid = i < cnt ? (uint64_t)id + a.shard.id_base : ISL_INVALID_ID;
        const uint4 rec = make_uint4(__float_as_uint(dist), 0u, (uint32_t)gid, (uint32_t)(gid >> 32));
        if (a.shard.packed) a.shard.packed[o] = rec;
        for (uint32_t r = 0; r < a.shard.n_peer; ++r) a.shard.peer[r][o] = rec;  // peer stores over NVLink
      }
    }
    if (lane == 0) {
      if (a.out_count) a.out_count[qi] = cnt;
      if (a.stats) {
        if (ADC && a.phase == 2) {  // the traversal launch wrote the other counters
          a.stats[qi].n_dist = n_dist;
          a.stats[qi].n_rerank = n_rerank;
        } else {
          isl_search_stats s;
          s.n_hop = n_hop;
          s.n_edge = n_edge;
          s.n_dist = n_dist;
          s.n_adc = n_adc;
          s.n_rerank = n_rerank;
          a.stats[qi] = s;
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace isl
