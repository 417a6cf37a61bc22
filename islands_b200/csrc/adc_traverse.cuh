// adc_traverse.cuh — the lean ADC traversal (the traversal half of "PQ ADC traversal + exact rerank") for one-byte
// codes with m = 16 / 32 subquantizers, the per-query table in shared memory and ef <= 512: one query per warp,
// the result set R in REGISTERS as an UNSORTED bag.
//
// The loop is the reference's best-first search (leann.rs:899-988) with the table distance (pq.rs:341-348) as the
// key; what changes against search_core.cuh is only how R is kept.  The reference needs three things from its two
// heaps: the worst entry of R (admission `d < worst`, eviction), the closest unexpanded entry (the next candidate),
// and at the very end the survivors.  None of them needs R sorted:
//   * worst entry  = argmax over the bag: per-lane IMNMX over its NR entries, redux.sync, a ballot to find the owner;
//     paid once per ADMITTED candidate (the new entry simply overwrites the evicted slot);
//   * next candidate = argmin over the unexpanded entries, the same reduction, once per hop;
//   * survivors    = the bag as it stands (the rerank launch admits all of them, order does not matter); only a
//     rerank limit needs the `limit` best, found by a rank count in shared memory once per query.
// The sorted register array of round 1 paid a 64-bit compare + ballot per row to find the position and a two-shuffle
// rotate per row to make room — ~150 instructions per admitted candidate, a third of all issued instructions
// (profiles/r01_adc_traverse_novis_ncu_full.txt).  Results are bit-identical: same admission rule, same eviction
// (the greatest (dist, id) key), same pop order (smallest (dist, id) key among the unexpanded), same tie list.
//
// Keys: kd = bits of the table distance (a square root: never negative, so the bit patterns order like the values;
// NaN is folded onto 0x7fc00000, greatest, OrderedFloat's rule), ki = id << 1 | expanded; slot i (row i / 32 of lane
// i % 32) is occupied iff i < r_len.  Reductions run on kd; the id decides only among entries that share the extreme kd.
// (Measured and rejected: every lane keeping its NR entries sorted, so that the worst entry is the greatest column
// top — fewer instructions per admission but a longer dependent chain: 7.85 ms against 7.16 ms at ef = 192.)
#pragma once

#include "search_core.cuh"

namespace isl {

template <int NR>
struct RegBag {
  uint32_t kd[NR];
  uint32_t ki[NR];
};

// Location of an entry of the bag: row (register index) and lane.  Warp-uniform.
struct BagPos {
  uint32_t row, lane;
};

// argmax of (kd, id) over the occupied slots (slot i = row i / 32 of lane i % 32 is occupied iff i < r_len).
// All lanes call; the result is warp-uniform.  Fast path: exactly one entry carries the greatest kd (the rule unless
// distances tie exactly); otherwise the id decides among the entries that share it.
template <int NR>
__device__ __forceinline__ void bag_argmax(const RegBag<NR>& b, uint32_t r_len, uint32_t* out_kd, uint32_t* out_ki, BagPos* pos) {
  const uint32_t lane = lane_id();
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) m = max(m, (uint32_t)(j * 32) + lane < r_len ? b.kd[j] : 0u);
  const uint32_t top = __reduce_max_sync(0xffffffffu, m);
  uint32_t row = 0, sel_i = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const bool eq = b.kd[j] == top && (uint32_t)(j * 32) + lane < r_len;
    if (eq && (cnt == 0 || (b.ki[j] >> 1) > (sel_i >> 1))) {  // within the lane: the greatest id among its matches
      row = j;
      sel_i = b.ki[j];
    }
    cnt += eq ? 1u : 0u;
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, cnt != 0);
  uint32_t owner = __ffs(bal) - 1;
  if (bal & (bal - 1)) {  // several lanes hold the greatest kd: the greatest id wins (ids are unique)
    const uint32_t top_id = __reduce_max_sync(0xffffffffu, cnt ? (sel_i >> 1) : 0u);
    owner = __ffs(__ballot_sync(0xffffffffu, cnt != 0 && (sel_i >> 1) == top_id)) - 1;
  }
  pos->row = __shfl_sync(0xffffffffu, row, owner);
  pos->lane = owner;
  *out_kd = top;
  *out_ki = __shfl_sync(0xffffffffu, sel_i, owner);
}

// argmin of (kd, id) over the occupied, unexpanded slots.  Returns false when there is none.
template <int NR>
__device__ __forceinline__ bool bag_argmin_unexpanded(const RegBag<NR>& b, uint32_t r_len, uint32_t* out_id, BagPos* pos) {
  const uint32_t lane = lane_id();
  uint32_t m = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const bool un = (uint32_t)(j * 32) + lane < r_len && !(b.ki[j] & 1u);
    m = min(m, un ? b.kd[j] : 0xffffffffu);
  }
  // a real entry never carries kd == 0xffffffff (NaN patterns are folded to 0x7fc00000)
  const uint32_t low = __reduce_min_sync(0xffffffffu, m);
  if (low == 0xffffffffu) return false;
  uint32_t row = 0, sel_id = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const bool eq = b.kd[j] == low && (uint32_t)(j * 32) + lane < r_len && !(b.ki[j] & 1u);
    if (eq && (b.ki[j] >> 1) < sel_id) {  // within the lane: the smallest id among its matches
      row = j;
      sel_id = b.ki[j] >> 1;
    }
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, sel_id != 0xffffffffu);
  uint32_t owner = __ffs(bal) - 1;
  if (bal & (bal - 1)) {  // several lanes hold the smallest kd: the smallest id wins
    const uint32_t id = __reduce_min_sync(0xffffffffu, sel_id);
    owner = __ffs(__ballot_sync(0xffffffffu, sel_id == id)) - 1;
  }
  pos->row = __shfl_sync(0xffffffffu, row, owner);
  pos->lane = owner;
  *out_id = __shfl_sync(0xffffffffu, sel_id, owner);
  return true;
}

template <int NR>
__global__ void __launch_bounds__(32) adc_traverse_kernel(const SearchArgs a) {
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [table or, at the end of a query, the rank-sort scratch][ties][admitted-id cache]
  float* lut_smem = reinterpret_cast<float*>(smem_raw);
  const uint32_t table_bytes = max(a.lut_smem_floats * 4u, a.ef * 8u);  // the rank-sort scratch of a rerank limit reuses the table
  uint2* ties = reinterpret_cast<uint2*>(smem_raw + ((table_bytes + 15u) & ~15u));
  uint16_t* idc = reinterpret_cast<uint16_t*>(ties + kTieCap);
  const uint32_t lane = lane_id();
  const uint32_t slot = blockIdx.x;
  const uint32_t ef = a.ef;
  const bool novis = a.novis != 0;
  uint32_t* vis = a.visited + (size_t)slot * a.vis_words;
  uint2* ties_spill = a.ties_global + (size_t)slot * ef;
  auto tie_ld = [&](uint32_t i) -> uint2 { return i < kTieCap ? ties[i] : __ldcg(ties_spill + (i - kTieCap)); };
  auto tie_st = [&](uint32_t i, uint2 v) {
    if (i < kTieCap)
      ties[i] = v;
    else
      __stcg(ties_spill + (i - kTieCap), v);
  };

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(a.work_counter, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= a.nq) break;

    // ---- per-query setup -------------------------------------------------------------------------
    if (novis) {
      uint32_t* c2 = reinterpret_cast<uint32_t*>(idc);
      for (uint32_t i = lane; i < kIdcEntries / 2; i += 32) c2[i] = 0xffffffffu;
    } else {
      uint4* v4 = reinterpret_cast<uint4*>(vis);
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (uint32_t i = lane; i < a.vis_words / 4; i += 32) __stcg(v4 + i, z);
      __threadfence();
    }
    if (a.luts == nullptr) {
      // build_distance_tables (pq.rs:307-338) straight into shared memory: LUT[j][c] = sum_t (q_jt - c_jct)^2, left
      // fold, one centroid per lane, four independent fold chains in flight
      const float* qv = a.queries + (size_t)qi * a.q_ld;
      const uint32_t nvec = a.pq_ld_sub >> 2;
      constexpr int CC = 4;
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
          if (nvec <= 8) {
            // every 16-byte piece of the four centroid rows in flight at once (up to 32 loads per lane: the bag is not
            // live yet, so the registers are free): one L2 latency per subquantizer instead of one per piece
            float4 y[CC][8];
#pragma unroll
            for (int v = 0; v < 8; ++v)
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) y[cc][v] = (uint32_t)v < nvec ? __ldg(row[cc] + v) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              if ((uint32_t)v < nvec) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t t = v * 4 + e;
                  if (t < a.pq_dsub) {
                    const float qe = __ldg(qs + t);
#pragma unroll
                    for (int cc = 0; cc < CC; ++cc) {
                      const float ye = e == 0 ? y[cc][v].x : (e == 1 ? y[cc][v].y : (e == 2 ? y[cc][v].z : y[cc][v].w));
                      const float diff = __fsub_rn(qe, ye);
                      acc[cc] = __fadd_rn(acc[cc], __fmul_rn(diff, diff));
                    }
                  }
                }
              }
            }
          } else {
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
          }
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            const uint32_t c = c0 + cc * 32 + lane;
            if (c < a.pq_ksub) lut_smem[j * a.pq_ksub + c] = acc[cc];
          }
        }
      }
    } else {
      const float* g = a.luts + (size_t)qi * a.pq_m * a.pq_ksub;
      for (uint32_t i = lane; i < a.lut_smem_floats; i += 32) lut_smem[i] = __ldg(g + i);
    }
    __syncwarp();

    RegBag<NR> R;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      R.kd[j] = 0xffffffffu;
      R.ki[j] = 0xffffffffu;
    }
    uint32_t r_len = 0, n_ties = 0, tie_next = kTieCap;
    uint32_t w_kd = 0xffffffffu, w_ki = 0xffffffffu;  // the worst entry (only meaningful once R is full) ...
    BagPos w_pos{0, 0};                               // ... and where it sits
    float w_d = 0.0f;
    uint64_t n_hop = 0, n_edge = 0, n_adc = 0;

    // ---- admission of one scored node (leann.rs:953-970), warp-uniform arguments ----------------------
    auto admit_one = [&](float d, uint32_t id) __attribute__((always_inline)) {
      const uint32_t nkd = (d != d) ? 0x7fc00000u : __float_as_uint(d);
      const uint32_t nki = id << 1;
      if (novis) {
        // no visited set: a node that is already in R was scored before and is admitted once
        bool same = false;
#pragma unroll
        for (int j = 0; j < NR; ++j) same = same || ((R.ki[j] ^ nki) < 2u);
        if (__any_sync(FULL, same)) return;
        if (lane == 0) idc[id & (kIdcEntries - 1)] = (uint16_t)(id >> kIdcBits);
      }
      if (r_len < ef) {
        const uint32_t row = r_len >> 5, ln = r_len & 31;
#pragma unroll
        for (int j = 0; j < NR; ++j)
          if ((uint32_t)j == row && lane == ln) {
            R.kd[j] = nkd;
            R.ki[j] = nki;
          }
        r_len++;
        if (r_len == ef) {
          bag_argmax<NR>(R, r_len, &w_kd, &w_ki, &w_pos);
          w_d = __uint_as_float(w_kd);
        }
        return;
      }
      // full: the new entry takes the slot of the worst one (pop max, leann.rs:966-968)
      const uint32_t e_kd = w_kd, e_ki = w_ki;
#pragma unroll
      for (int j = 0; j < NR; ++j)
        if ((uint32_t)j == w_pos.row && lane == w_pos.lane) {
          R.kd[j] = nkd;
          R.ki[j] = nki;
        }
      bag_argmax<NR>(R, r_len, &w_kd, &w_ki, &w_pos);
      w_d = __uint_as_float(w_kd);
      if (!(e_ki & 1u) && !of_lt(w_d, __uint_as_float(e_kd))) {
        // an evicted, unexpanded node stays expandable while its distance equals the worst distance in R
        // (leann.rs:924-928 uses a strict `>`); see search_core.cuh for the capacity argument
        if (n_ties == tie_next) {
          uint32_t kept = 0;
          for (uint32_t i = 0; i < n_ties; ++i) {
            const uint2 t = tie_ld(i);
            if (!of_lt(w_d, __uint_as_float(t.x))) {
              __syncwarp();
              if (lane == 0) tie_st(kept, t);
              kept++;
            }
          }
          n_ties = kept;
          tie_next = min(kTieCap + ef, max(kTieCap, 2 * kept));
          __syncwarp();
        }
        if (n_ties >= kTieCap + ef) {
          if (lane == 0) atomicExch(a.error_flag, 1u);
        } else {
          if (lane == 0) tie_st(n_ties, make_uint2(e_kd, e_ki >> 1));
          n_ties++;
          __syncwarp();
        }
      }
    };
    // admission of up to 32 scored nodes in list order: lanes that can still be admitted, replayed sequentially
    auto admit_values = [&](bool ok, float dn, uint32_t cid) __attribute__((always_inline)) {
      uint32_t mask = __ballot_sync(FULL, ok && (r_len < ef || dn < w_d));
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const float dj = __shfl_sync(FULL, dn, j);
        const uint32_t idj = __shfl_sync(FULL, cid, j);
        if (r_len < ef || dj < w_d) admit_one(dj, idj);  // raw f32 `<` (leann.rs:959)
      }
    };
    auto adc_of = [&](uint32_t nid) -> float {  // table_distance (pq.rs:341-348) of one node
      float sacc = 0.0f;
      const uint8_t* cd = a.codes8 + (size_t)nid * a.pq_m;
      for (uint32_t j = 0; j < a.pq_m; ++j) sacc = __fadd_rn(sacc, lut_smem[j * a.pq_ksub + cd[j]]);
      return __fsqrt_rn(sacc);
    };

    // ---- entry point (leann.rs:911-916) -----------------------------------------------------------
    {
      const uint32_t entry = a.entry;
      if (!novis && lane == 0) atomicOr(vis + (entry >> 5), 1u << (entry & 31));
      const float d0 = adc_of(entry);
      admit_values(lane == 0, d0, entry);
      n_adc = 1;
    }

    // ---- main loop (leann.rs:922-972) -------------------------------------------------------------
    const uint32_t nv = a.pq_m >> 4;  // 16-byte pieces per code row
    for (;;) {
      uint32_t cur;
      BagPos cp;
      if (bag_argmin_unexpanded<NR>(R, r_len, &cur, &cp)) {
#pragma unroll
        for (int j = 0; j < NR; ++j)
          if ((uint32_t)j == cp.row && lane == cp.lane) R.ki[j] |= 1u;
        if (r_len == ef && cp.row == w_pos.row && cp.lane == w_pos.lane) w_ki |= 1u;  // the cached copy of the worst entry sees the flag too
      } else {
        // smallest live tie, if any; every lane scans the whole list so that the result is provably warp-uniform
        int best = -1;
        for (uint32_t i = 0; i < n_ties; ++i) {
          const uint2 t = tie_ld(i);
          if (r_len >= ef && of_lt(w_d, __uint_as_float(t.x))) continue;  // stale
          if (best < 0 || key_lt(__uint_as_float(t.x), t.y, __uint_as_float(tie_ld(best).x), tie_ld(best).y)) best = (int)i;
        }
        if (best < 0) break;
        cur = tie_ld(best).y;
        __syncwarp();
        if (lane == 0) tie_st(best, tie_ld(n_ties - 1));
        n_ties--;
        __syncwarp();
      }

      // neighbour list of `cur`: CSR, or fixed-stride rows padded with 0xffffffff
      uint64_t start;
      uint32_t deg;
      bool sentinel = false;
      if (a.offsets) {
        start = __ldg(a.offsets + cur);
        deg = (uint32_t)(__ldg(a.offsets + cur + 1) - start);
      } else {
        start = (uint64_t)cur * a.adj_stride;
        deg = a.adj_stride;
        sentinel = true;
      }
      n_hop++;
      if (!sentinel) n_edge += deg;
      if (novis) __syncwarp();  // lane 0's writes to the admitted-id cache are read by every lane below

      // 64 list positions per pass (two per lane): ids, then visited test AND code rows of all positions in flight
      // together, then table distances, then admission in list order
      for (uint32_t b = 0; b < deg; b += 64) {
        uint32_t nid[2];
        bool chk[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t i = b + r * 32 + lane;
          nid[r] = i < deg ? __ldg(a.nbrs + start + i) : 0xffffffffu;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const bool valid = nid[r] != 0xffffffffu;
          if (sentinel) n_edge += __popc(__ballot_sync(FULL, valid));
          chk[r] = valid && nid[r] < a.n;
          if (!a.lists_unique && !novis) {  // first of its value in this half
            const uint32_t same = __match_any_sync(FULL, nid[r]);
            chk[r] = chk[r] && lane == (uint32_t)(__ffs(same) - 1);
          }
        }
        if (!a.lists_unique && !novis && b + 32 < deg) {  // a value of the second half that already occurs in the first
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
          const float* lj = lut_smem;
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
    }

    // ---- hand the survivors to the rerank / recompute step ----------------------------------------------
    // Without a rerank limit their order does not matter (the rerank admits all of them into an empty R of the
    // same capacity); with a limit the `limit` best by (table distance, id) are selected by a rank count.
    uint32_t n_surv = r_len;
    if (a.rerank_limit && a.rerank_limit < r_len) {
      n_surv = a.rerank_limit;
      __syncwarp();
      uint2* sk = reinterpret_cast<uint2*>(smem_raw);  // the table is no longer needed
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const uint32_t idx = j * 32 + lane;
        if (idx < r_len) sk[idx] = make_uint2(R.kd[j], R.ki[j] >> 1);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const uint32_t idx = j * 32 + lane;
        if (idx < r_len) {
          const uint64_t mine = ((uint64_t)R.kd[j] << 32) | (R.ki[j] >> 1);
          uint32_t rank = 0;
          for (uint32_t t = 0; t < r_len; ++t) {
            const uint2 o = sk[t];
            rank += (((uint64_t)o.x << 32) | o.y) < mine ? 1u : 0u;
          }
          if (rank < n_surv) a.surv_ids[(size_t)qi * ef + rank] = R.ki[j] >> 1;
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const uint32_t idx = j * 32 + lane;
        if (idx < r_len) a.surv_ids[(size_t)qi * ef + idx] = R.ki[j] >> 1;
      }
    }
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
  }
}

__host__ __device__ constexpr size_t adc_traverse_smem_bytes(uint32_t lut_floats, uint32_t ef) {
  const size_t table = lut_floats * 4u > ef * 8u ? lut_floats * 4u : ef * 8u;
  return ((table + 15u) & ~(size_t)15u) + (size_t)kTieCap * 8 + (size_t)kIdcEntries * 2;
}

}  // namespace isl
