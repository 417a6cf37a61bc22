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
// Keys: the distance part of kd = bits of the table distance + 1 (a square root: never negative, so the bit patterns
// order like the values; NaN is folded onto 0x7fc00000, greatest, OrderedFloat's rule); slot i (row i / 32 of lane
// i % 32) is occupied iff i < r_len, empty slots are neutral for both reductions (see RegBag).  Reductions run on kd;
// the id decides only among entries that share the extreme distance.
// The per-query table holds bfloat16-rounded entries, 2 bytes each (common.cuh: bf16_round_bits; the definition of the
// ADC traversal in include/islands_b200.h): 8 KB instead of 16 KB at m = 32, ksub = 128, so 21 queries are resident per
// SM instead of 12 (profiles/r02_adc_bag_ncu.txt: issue slots 47 % -> 64 % busy).
// (Measured and rejected: every lane keeping its NR entries sorted, so that the worst entry is the greatest column
// top — fewer instructions per admission but a longer dependent chain: 7.85 ms against 7.16 ms at ef = 192.)
#pragma once

#include "search_core.cuh"

namespace isl {

// Bag entry: kd = (bits of the table distance + 1) | expanded << 31, ki = id << 1 | expanded.  An EMPTY slot is
// kd = 0x80000000 ("expanded", distance part 0): it loses every max over the distance parts (a real entry is >= 1) and
// every unsigned min over the raw words (an unexpanded entry is < 0x80000000), so neither reduction needs a validity
// test.  The flag is kept twice on purpose: in kd it makes the argmin over the unexpanded entries one IMNMX per row; the
// copy in ki travels with the id of the worst entry.  (Carrying it in a variable of its own made ptxas treat the branch
// on it as divergent and wrap every collective of the loop in BSSY / WARPSYNC / BRA.DIV: 4656 -> 6472 SASS instructions.)
constexpr uint32_t kBagExpanded = 0x80000000u;
constexpr uint32_t kBagDistMask = 0x7fffffffu;

template <int NR>
struct RegBag {
  uint32_t kd[NR];
  uint32_t ki[NR];
};

// Location of an entry of the bag: row (register index) and lane.  Warp-uniform.
struct BagPos {
  uint32_t row, lane;
};

// argmax of (distance, id) over the bag.  All lanes call; the result is warp-uniform.  Fast path: exactly one entry
// carries the greatest distance (the rule unless distances tie exactly); otherwise the greatest id among them wins.
// out_bits = bits of that distance, out_ki = id << 1 | expanded of the entry.
template <int NR>
__device__ __forceinline__ void bag_argmax(const RegBag<NR>& b, uint32_t* out_bits, uint32_t* out_ki, BagPos* pos) {
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) m = max(m, b.kd[j] & kBagDistMask);
  const uint32_t top = __reduce_max_sync(FULL, m);
  uint32_t hit = 0, sel_ki = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if ((b.kd[j] & kBagDistMask) == top) {
      hit |= 1u << j;
      sel_ki = b.ki[j];
    }
  }
  const uint32_t bal = __ballot_sync(FULL, hit != 0);
  const bool dup_in_lane = __any_sync(FULL, (hit & (hit - 1)) != 0);
  const bool multi = ((bal & (bal - 1)) != 0) | dup_in_lane;
  uint32_t owner = __ffs(bal) - 1;
  uint32_t row = 31 - __clz(hit);  // the lane's only match on the fast path (meaningful in the owner)
  if (multi) {  // several entries share the greatest distance: the greatest id wins (ids are unique in the bag)
    bool have = false;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      if ((b.kd[j] & kBagDistMask) == top && (!have || b.ki[j] > sel_ki)) {  // ki orders like the id
        have = true;
        row = j;
        sel_ki = b.ki[j];
      }
    }
    const uint32_t top_ki = __reduce_max_sync(FULL, have ? sel_ki : 0u);
    owner = __ffs(__ballot_sync(FULL, have && sel_ki == top_ki)) - 1;
  }
  pos->row = __shfl_sync(FULL, row, owner);
  pos->lane = owner;
  *out_bits = top - 1;
  *out_ki = __shfl_sync(FULL, sel_ki, owner);
}

// argmin of (distance, id) over the unexpanded entries.  Returns false when there is none.
template <int NR>
__device__ __forceinline__ bool bag_argmin_unexpanded(const RegBag<NR>& b, uint32_t* out_id, BagPos* pos) {
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t m = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < NR; ++j) m = min(m, b.kd[j]);
  const uint32_t low = __reduce_min_sync(FULL, m);
  if (low >= kBagExpanded) return false;  // every entry is expanded (or empty)
  uint32_t hit = 0, sel_ki = 0;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (b.kd[j] == low) {
      hit |= 1u << j;
      sel_ki = b.ki[j];
    }
  }
  const uint32_t bal = __ballot_sync(FULL, hit != 0);
  const bool dup_in_lane = __any_sync(FULL, (hit & (hit - 1)) != 0);
  const bool multi = ((bal & (bal - 1)) != 0) | dup_in_lane;
  uint32_t owner = __ffs(bal) - 1;
  uint32_t row = 31 - __clz(hit);
  if (multi) {  // several unexpanded entries share the smallest distance: the smallest id wins
    bool have = false;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      if (b.kd[j] == low && (!have || b.ki[j] < sel_ki)) {
        have = true;
        row = j;
        sel_ki = b.ki[j];
      }
    }
    const uint32_t low_ki = __reduce_min_sync(FULL, have ? sel_ki : 0xffffffffu);
    owner = __ffs(__ballot_sync(FULL, have && sel_ki == low_ki)) - 1;
  }
  pos->row = __shfl_sync(FULL, row, owner);
  pos->lane = owner;
  *out_id = __shfl_sync(FULL, sel_ki, owner) >> 1;  // unexpanded: the flag bit is clear
  return true;
}

// Resident CTAs per SM the kernel is compiled for: 2-byte table entries leave room for 21 queries per SM (8 KB table +
// tie list + id cache), which needs <= 96 registers per thread; the large bags (ef up to 512) keep 128.
template <int NR>
struct BagOccupancy {
  static constexpr int kMinCtas = NR <= 8 ? 21 : 16;
};

// table_distance (pq.rs:341-348) of one node from its code row (two 16-byte pieces): the f32 left fold over the
// subquantizers of the bfloat16 table entries, then sqrt.  KS = 128: ksub is exactly 128, so the row offsets of the table
// are immediates, and all four codes of a word are doubled at once (w + w: every code is < 128, no carry crosses a byte)
// so that the extracted byte IS the byte offset of the 2-byte entry — PRMT, LDS.U16, shift, FADD per subquantizer.
// KS = 0: any ksub <= 256.
template <int KS>
__device__ __forceinline__ float bag_table_distance(const uint16_t* lut16, const uint4 (&cw)[2], uint32_t nv, uint32_t ksub) {
  float sacc = 0.0f;
  if constexpr (KS == 128) {
    const unsigned char* base = reinterpret_cast<const unsigned char*>(lut16);
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if ((uint32_t)v < nv) {
        const uint32_t w[4] = {cw[v].x << 1, cw[v].y << 1, cw[v].z << 1, cw[v].w << 1};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint32_t off = __byte_perm(w[k >> 2], 0, 0x4440 + (k & 3));
          const uint16_t e = *reinterpret_cast<const uint16_t*>(base + off + (v * 16 + k) * (KS * 2));
          sacc = __fadd_rn(sacc, __uint_as_float((uint32_t)e << 16));
        }
      }
    }
  } else {
    const uint16_t* lj = lut16;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if ((uint32_t)v < nv) {
        const uint32_t w[4] = {cw[v].x, cw[v].y, cw[v].z, cw[v].w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint32_t code = __byte_perm(w[k >> 2], 0, 0x4440 + (k & 3));
          sacc = __fadd_rn(sacc, __uint_as_float((uint32_t)lj[code] << 16));
          lj += ksub;
        }
      }
    }
  }
  return __fsqrt_rn(sacc);
}

template <int NR, int KS>
__global__ void __launch_bounds__(32, BagOccupancy<NR>::kMinCtas) adc_traverse_kernel(const SearchArgs a) {
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [table (bf16 bits, 2 bytes per entry) or, at the end of a query, the rank-sort scratch][ties][admitted-id cache]
  uint16_t* lut16 = reinterpret_cast<uint16_t*>(smem_raw);
  const uint32_t table_bytes = max(a.lut_smem_floats * 2u, a.ef * 8u);  // the rank-sort scratch of a rerank limit reuses the table
  uint2* ties = reinterpret_cast<uint2*>(smem_raw + ((table_bytes + 15u) & ~15u));
  uint16_t* idc = reinterpret_cast<uint16_t*>(ties + kTieCap);
  const uint32_t lane = lane_id();
  const uint32_t slot = blockIdx.x;
  const uint32_t ef = a.ef;
  const bool novis = a.novis != 0;
  const bool want_stats = a.stats != nullptr;
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
      // fold, one centroid per lane, four independent fold chains and three 16-byte pieces of each row in flight; the
      // finished sum is rounded to bfloat16 (common.cuh: bf16_round_bits)
      const float* qv = a.queries + (size_t)qi * a.q_ld;
      const uint32_t nvec = a.pq_ld_sub >> 2;
      constexpr int CC = 4, VC = 3;
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
          for (uint32_t v0 = 0; v0 < nvec; v0 += VC) {
            float4 y[CC][VC];
#pragma unroll
            for (int v = 0; v < VC; ++v)
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) y[cc][v] = v0 + v < nvec ? __ldg(row[cc] + v0 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < VC; ++v) {
              if (v0 + v < nvec) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t t = (v0 + v) * 4 + e;
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
          }
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            const uint32_t c = c0 + cc * 32 + lane;
            if (c < a.pq_ksub) lut16[j * a.pq_ksub + c] = (uint16_t)(bf16_round_bits(__float_as_uint(acc[cc])) >> 16);
          }
        }
      }
    } else {
      const float* g = a.luts + (size_t)qi * a.pq_m * a.pq_ksub;
      for (uint32_t i = lane; i < a.lut_smem_floats; i += 32) lut16[i] = (uint16_t)(bf16_round_bits(__float_as_uint(__ldg(g + i))) >> 16);
    }
    __syncwarp();

    RegBag<NR> R;
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      R.kd[j] = kBagExpanded;
      R.ki[j] = 0xffffffffu;
    }
    uint32_t r_len = 0, n_ties = 0, tie_next = kTieCap;
    uint32_t w_bits = 0xffffffffu, w_ki = 0xffffffffu;  // the worst entry (only meaningful once R is full) ...
    BagPos w_pos{0, 0};                                 // ... and where it sits
    float w_d = 0.0f;
    uint64_t n_hop = 0, n_edge = 0, n_adc = 0;

    // ---- admission of one scored node (leann.rs:953-970), warp-uniform arguments ----------------------
    auto admit_one = [&](float d, uint32_t id) __attribute__((always_inline)) {
      const uint32_t nbits = (d != d) ? 0x7fc00000u : __float_as_uint(d);
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
        const uint32_t row = r_len >> 5;
        const bool me = lane == (r_len & 31);
#pragma unroll
        for (int j = 0; j < NR; ++j)
          if ((uint32_t)j == row && me) {
            R.kd[j] = nbits + 1u;
            R.ki[j] = nki;
          }
        r_len++;
        if (r_len == ef) {
          bag_argmax<NR>(R, &w_bits, &w_ki, &w_pos);
          w_d = __uint_as_float(w_bits);
        }
        return;
      }
      // full: the new entry takes the slot of the worst one (pop max, leann.rs:966-968)
      const uint32_t e_bits = w_bits, e_ki = w_ki;
      const bool me = lane == w_pos.lane;
#pragma unroll
      for (int j = 0; j < NR; ++j)
        if ((uint32_t)j == w_pos.row && me) {
          R.kd[j] = nbits + 1u;
          R.ki[j] = nki;
        }
      bag_argmax<NR>(R, &w_bits, &w_ki, &w_pos);
      w_d = __uint_as_float(w_bits);
      if (!(e_ki & 1u) && !of_lt(w_d, __uint_as_float(e_bits))) {
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
          if (lane == 0) tie_st(n_ties, make_uint2(e_bits, e_ki >> 1));
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
    auto lut_at = [&](uint32_t i) -> float { return __uint_as_float((uint32_t)lut16[i] << 16); };
    auto adc_of = [&](uint32_t nid) -> float {  // table_distance (pq.rs:341-348) of one node
      float sacc = 0.0f;
      const uint8_t* cd = a.codes8 + (size_t)nid * a.pq_m;
      for (uint32_t j = 0; j < a.pq_m; ++j) sacc = __fadd_rn(sacc, lut_at(j * a.pq_ksub + cd[j]));
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
      if (bag_argmin_unexpanded<NR>(R, &cur, &cp)) {
        const bool me = lane == cp.lane;
#pragma unroll
        for (int j = 0; j < NR; ++j)
          if ((uint32_t)j == cp.row && me) {
            R.kd[j] |= kBagExpanded;
            R.ki[j] |= 1u;
          }
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
          if (sentinel && want_stats) n_edge += __popc(__ballot_sync(FULL, valid));
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
          const float adc = bag_table_distance<KS>(lut16, cw[r], nv, a.pq_ksub);
          const bool unv = chk[r] && !(old[r] & (1u << (nid[r] & 31)));
          if (want_stats) n_adc += __popc(__ballot_sync(FULL, unv));
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
        if (idx < r_len) sk[idx] = make_uint2(R.kd[j] & kBagDistMask, R.ki[j] >> 1);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const uint32_t idx = j * 32 + lane;
        if (idx < r_len) {
          const uint64_t mine = ((uint64_t)(R.kd[j] & kBagDistMask) << 32) | (R.ki[j] >> 1);
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

__host__ __device__ constexpr size_t adc_traverse_smem_bytes(uint32_t lut_entries, uint32_t ef) {
  const size_t table = lut_entries * 2u > ef * 8u ? lut_entries * 2u : ef * 8u;
  return ((table + 15u) & ~(size_t)15u) + (size_t)kTieCap * 8 + (size_t)kIdcEntries * 2;
}

}  // namespace isl
