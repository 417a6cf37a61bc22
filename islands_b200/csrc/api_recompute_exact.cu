// api_recompute_exact.cu — LeannIndex::search_with_params with a RECOMPUTING provider, hop by hop:
// the reference's loop (leann.rs:899-988) asks the provider for the embeddings of every hop's unvisited
// neighbours (`provider.compute_embeddings_batch(&to_compute)`, :947-950) and scores them exactly.  Here the provider
// is the attached encoder over the nodes' token rows, and the hops of all queries of a batch run in lockstep so that
// one encoder pass serves the whole frontier of the batch (docs/leann-specification.md:364-394, "dynamic batching"):
//
//   round r:  expand kernel   one warp per query: pop the next candidate (first unexpanded entry of R, else the
//                             smallest live tie), fetch its list, mark the unvisited neighbours, apply the frontier
//                             pruning (leann.rs:991-1016) -> to_compute[q][..]
//             distinct nodes of all to_compute lists -> token rows -> encoder -> embeddings (+ squared norms)
//             admit kernel    one warp per query: reference-order distances of its to_compute nodes against the
//                             recomputed rows, admission in list order (leann.rs:953-970)
//   until every query has drained its candidates.
//
// Per-query state (R, tie list, visited bitset, counters) lives in global memory between the rounds.  Nothing here is
// tuned: a round costs an encoder forward over thousands of sequences, the two search kernels are noise beside it.
// With an encoder whose output for a row does not depend on its batch (encoder.cu) the results are bit-identical to
// isl_index_search over an index that stores those embeddings — that is the parity test of this path.
#include <cub/cub.cuh>

#include <algorithm>

#include "api_common.h"
#include "search_core.cuh"

namespace isl {
namespace {

constexpr uint32_t kXBit = 0x80000000u;  // "expanded" flag in the id word of an R entry
constexpr uint32_t kXTie = 64;           // tie entries beyond ef (the list holds ef + kXTie; live ties never exceed ef - 1)

struct HopArgs {
  // index
  const uint64_t* offsets;  // CSR (null => padded rows)
  const uint32_t* nbrs;
  uint32_t adj_stride;
  uint32_t n, d, ld;
  uint32_t entry;
  int32_t metric;
  float prune_ratio;
  int32_t strategy;
  uint64_t prune_seed;
  const uint32_t* deg_counts;
  // batch
  const float* queries;  // [nq][q_ld]
  uint32_t q_ld, nq, ef, k, u_cap, vis_words, tie_cap;
  // per-query state
  uint2* R;            // [nq][ef] ascending (dist, id | expanded)
  uint2* ties;         // [nq][tie_cap]
  uint32_t* vis;       // [nq][vis_words]
  uint32_t* cand;      // [nq][u_cap] this round's to_compute list
  uint32_t* meta;      // [nq][8]: r_len, first_unexp, n_ties, cand_cnt, done, started, draw counter (lo, hi)
  float* q_sqnorm;     // [nq]
  isl_search_stats* stats;  // [nq]
  unsigned int* active;     // number of queries that still have a round to run
  // recomputed rows of this round
  const float* emb;         // [unique][ld]
  const float* emb_sq;      // [unique]
  const uint32_t* row_of_id;  // [n]
};

__device__ __forceinline__ uint32_t prune_keep_x(float prune_ratio, int strategy, uint32_t n_cands, uint32_t r_len, uint32_t ef) {
  if (prune_ratio == 0.0f || n_cands == 0) return n_cands;  // leann.rs:991-1016
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

// Σ q*q folded left to right (distance.rs:78) for every query, once per batch.
__global__ void query_sqnorm_kernel(const float* __restrict__ q, uint32_t q_ld, uint32_t d, uint32_t nq, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  float s = 0.0f;
  for (uint32_t t = 0; t < d; ++t) {
    const float x = q[(size_t)i * q_ld + t];
    s = __fadd_rn(s, __fmul_rn(x, x));
  }
  out[i] = s;
}

// One warp per query: advance to the next hop that has something to compute (or finish).
__global__ void __launch_bounds__(128) rc_expand_kernel(const HopArgs a) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= a.nq) return;
  uint32_t* meta = a.meta + (size_t)q * 8;
  uint32_t* cand = a.cand + (size_t)q * a.u_cap;
  uint32_t* vis = a.vis + (size_t)q * a.vis_words;
  uint2* R = a.R + (size_t)q * a.ef;
  uint2* ties = a.ties + (size_t)q * a.tie_cap;
  if (meta[4]) {  // done
    if (lane == 0) meta[3] = 0;
    return;
  }
  if (!meta[5]) {  // first round: the entry point (leann.rs:911-916)
    if (lane == 0) {
      cand[0] = a.entry;
      atomicOr(vis + (a.entry >> 5), 1u << (a.entry & 31));
      meta[3] = 1;
      meta[5] = 1;
      a.stats[q].n_dist = 1;
      atomicAdd(a.active, 1u);
    }
    return;
  }
  uint32_t r_len = meta[0], first_unexp = meta[1], n_ties = meta[2];
  uint64_t draw_ctr = ((uint64_t)meta[7] << 32) | meta[6];  // Proportional pruning: draws consumed so far
  uint64_t n_hop = 0, n_edge = 0, n_dist = 0;
  uint32_t ucnt = 0, keep = 0;
  bool finished = false;
  for (;;) {  // leann.rs:922-941: pop until a hop yields unvisited neighbours
    uint32_t cur;
    if (first_unexp < r_len) {
      const uint2 e = R[first_unexp];
      cur = e.y;
      __syncwarp();
      if (lane == 0) R[first_unexp] = make_uint2(e.x, e.y | kXBit);
      __syncwarp();
      uint32_t nxt = r_len;
      for (uint32_t b = first_unexp + 1; b < r_len; b += 32) {
        const uint32_t i = b + lane;
        const bool un = i < r_len && !(R[i].y & kXBit);
        const uint32_t bal = __ballot_sync(0xffffffffu, un);
        if (bal) {
          nxt = b + __ffs(bal) - 1;
          break;
        }
      }
      first_unexp = nxt;
    } else {
      const float wd = __uint_as_float(R[r_len - 1].x);
      int best = -1;
      for (uint32_t i = 0; i < n_ties; ++i) {
        const uint2 t = ties[i];
        if (r_len >= a.ef && of_lt(wd, __uint_as_float(t.x))) continue;  // stale (leann.rs:924-928)
        if (best < 0 || key_lt(__uint_as_float(t.x), t.y, __uint_as_float(ties[best].x), ties[best].y)) best = (int)i;
      }
      if (best < 0) {
        finished = true;
        break;
      }
      cur = ties[best].y;
      __syncwarp();
      if (lane == 0) ties[best] = ties[n_ties - 1];
      n_ties--;
      __syncwarp();
    }
    uint64_t start;
    uint32_t deg;
    bool sentinel = false;
    if (a.offsets) {
      start = a.offsets[cur];
      deg = (uint32_t)(a.offsets[cur + 1] - start);
    } else {
      start = (uint64_t)cur * a.adj_stride;
      deg = a.adj_stride;
      sentinel = true;
    }
    n_hop++;
    if (!sentinel) n_edge += deg;
    ucnt = 0;
    for (uint32_t b = 0; b < deg; b += 32) {  // unvisited neighbours in list order, marked before pruning (leann.rs:933-937)
      const uint32_t i = b + lane;
      bool valid = i < deg;
      uint32_t nid = 0xffffffffu;
      if (valid) nid = a.nbrs[start + i];
      if (sentinel) {
        valid = valid && nid != 0xffffffffu;
        n_edge += __popc(__ballot_sync(0xffffffffu, valid));
      }
      const uint32_t same = __match_any_sync(0xffffffffu, nid);
      const bool first = lane == (uint32_t)(__ffs(same) - 1);
      bool unv = false;
      if (valid && first && nid < a.n) {
        const uint32_t bit = 1u << (nid & 31);
        unv = !(atomicOr(vis + (nid >> 5), bit) & bit);
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, unv);
      if (unv) cand[ucnt + __popc(bal & ((1u << lane) - 1))] = nid;
      ucnt += __popc(bal);
    }
    __syncwarp();
    if (ucnt == 0) continue;  // leann.rs:939-941
    keep = (a.strategy == ISL_PRUNE_PROPORTIONAL && a.prune_ratio != 0.0f)
               ? prune_proportional(a.prune_ratio, cand, ucnt, a.deg_counts, a.prune_seed, q, &draw_ctr)
               : prune_keep_x(a.prune_ratio, a.strategy, ucnt, r_len, a.ef);  // leann.rs:944
    n_dist += keep;
    break;
  }
  if (lane == 0) {
    meta[1] = first_unexp;
    meta[2] = n_ties;
    meta[3] = finished ? 0 : keep;
    meta[6] = (uint32_t)draw_ctr;
    meta[7] = (uint32_t)(draw_ctr >> 32);
    if (finished) {
      meta[4] = 1;
      atomicSub(a.active, 1u);
    }
    isl_search_stats s = a.stats[q];
    s.n_hop += n_hop;
    s.n_edge += n_edge;
    s.n_dist += n_dist;
    a.stats[q] = s;
  }
}

// Every node named by some to_compute list of this round.
__global__ void rc_mark_kernel(const uint32_t* __restrict__ cand, const uint32_t* __restrict__ meta, uint32_t nq, uint32_t u_cap,
                               uint32_t* __restrict__ flags) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const uint32_t cnt = meta[(size_t)q * 8 + 3];
  for (uint32_t i = lane; i < cnt; i += 32) flags[cand[(size_t)q * u_cap + i]] = 1u;
}

__global__ void rc_gather_tokens_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ rows, uint32_t n,
                                        const int32_t* __restrict__ tokens, const int32_t* __restrict__ lengths, uint32_t S,
                                        int32_t* __restrict__ out_tok, int32_t* __restrict__ out_len) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
  for (uint32_t id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; id < n; id += warps) {
    if (!flags[id]) continue;
    const uint32_t r = rows[id];
    for (uint32_t i = lane; i < S; i += 32) out_tok[(size_t)r * S + i] = tokens[(size_t)id * S + i];
    if (lane == 0) out_len[r] = lengths[id];
  }
}

// One warp per query: exact distances of this round's to_compute nodes (lane per node, reference-order fold straight
// from global memory) and their admission in list order (leann.rs:953-970).
template <int ACC>
__global__ void __launch_bounds__(128) rc_admit_kernel(const HopArgs a) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= a.nq) return;
  uint32_t* meta = a.meta + (size_t)q * 8;
  const uint32_t cnt = meta[3];
  if (cnt == 0) return;
  const uint32_t* cand = a.cand + (size_t)q * a.u_cap;
  uint2* R = a.R + (size_t)q * a.ef;
  uint2* ties = a.ties + (size_t)q * a.tie_cap;
  const float* qv = a.queries + (size_t)q * a.q_ld;
  const float na = a.q_sqnorm[q];
  const uint32_t ef = a.ef;
  uint32_t r_len = meta[0], first_unexp = meta[1], n_ties = meta[2];
  for (uint32_t b = 0; b < cnt; b += 32) {
    const uint32_t i = b + lane;
    float dn = 0.0f;
    uint32_t cid = 0;
    if (i < cnt) {
      cid = cand[i];
      const uint32_t row = a.row_of_id[cid];
      const float* y = a.emb + (size_t)row * a.ld;
      float acc = 0.0f;
      for (uint32_t t = 0; t < a.d; ++t) acc = acc_step<ACC>(acc, qv[t], y[t]);
      dn = finalize_distance(a.metric, acc, na, a.metric == ISL_METRIC_COSINE ? a.emb_sq[row] : 0.0f);
    }
    float worst = 0.0f;
    if (r_len > 0) worst = __uint_as_float(R[r_len - 1].x);
    uint32_t mask = __ballot_sync(0xffffffffu, i < cnt && (r_len < ef || dn < worst));
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const float dj = __shfl_sync(0xffffffffu, dn, j);
      const uint32_t idj = __shfl_sync(0xffffffffu, cid, j);
      bool add = r_len < ef;
      if (!add) add = dj < __uint_as_float(R[ef - 1].x);  // raw f32 `<` (leann.rs:959)
      if (!add) continue;
      // sorted insert by (dist, id); the greatest entry leaves when R is full (leann.rs:961-968)
      uint32_t lo = 0, hi = r_len;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint2 e = R[mid];
        if (key_lt(__uint_as_float(e.x), e.y & ~kXBit, dj, idj))
          lo = mid + 1;
        else
          hi = mid;
      }
      const bool full = r_len == ef;
      uint2 evicted = make_uint2(0, 0);
      if (full) evicted = R[ef - 1];
      const int top = full ? (int)ef - 1 : (int)r_len;
      for (int t = top; t > (int)lo; t -= 32) {
        const int p = t - (int)lane;
        const bool act = p > (int)lo;
        uint2 e = make_uint2(0, 0);
        if (act) e = R[p - 1];
        __syncwarp();
        if (act) R[p] = e;
        __syncwarp();
      }
      if (lane == 0) R[lo] = make_uint2(__float_as_uint(dj), idj);
      __syncwarp();
      if (!full) r_len++;
      if (lo <= first_unexp) first_unexp = lo;
      if (full && !(evicted.y & kXBit)) {  // evicted, unexpanded: expandable while its distance equals the worst one
        const float wd = __uint_as_float(R[ef - 1].x);
        if (!of_lt(wd, __uint_as_float(evicted.x))) {
          if (n_ties == a.tie_cap) {  // drop stale ties (live ones never exceed ef - 1)
            uint32_t kept = 0;
            for (uint32_t t = 0; t < n_ties; ++t) {
              const uint2 tt = ties[t];
              if (!of_lt(wd, __uint_as_float(tt.x))) {
                __syncwarp();
                if (lane == 0) ties[kept] = tt;
                kept++;
              }
            }
            n_ties = kept;
            __syncwarp();
          }
          if (lane == 0) ties[n_ties] = evicted;
          n_ties++;
          __syncwarp();
        }
      }
    }
  }
  if (lane == 0) {
    meta[0] = r_len;
    meta[1] = first_unexp;
    meta[2] = n_ties;
  }
}

// take(k) of the sorted R (leann.rs:895, :984-987).
__global__ void rc_results_kernel(const uint2* __restrict__ R, const uint32_t* __restrict__ meta, uint32_t nq, uint32_t ef, uint32_t k,
                                  uint64_t* __restrict__ out_ids, float* __restrict__ out_dist, uint32_t* __restrict__ out_count) {
  const uint32_t q = blockIdx.x;
  if (q >= nq) return;
  const uint32_t r_len = meta[(size_t)q * 8];
  const uint32_t cnt = r_len < k ? r_len : k;
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    uint64_t id = ISL_INVALID_ID;
    float d = __int_as_float(0x7f800000);
    if (i < cnt) {
      const uint2 e = R[(size_t)q * ef + i];
      id = e.y & ~kXBit;
      d = __uint_as_float(e.x);
    }
    out_ids[(size_t)q * k + i] = id;
    out_dist[(size_t)q * k + i] = d;
  }
  if (threadIdx.x == 0) out_count[q] = cnt;
}

}  // namespace
}  // namespace isl

using namespace isl;

extern "C" {

isl_status isl_index_search_recompute(const isl_index* idx, const float* queries, uint64_t nq, uint32_t query_dim, uint32_t k,
                                      uint32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* out_count,
                                      isl_search_stats* stats) try {
  bool trivial;
  ISL_TRY(search_checks(idx, queries, nq, query_dim, k, &ef, &trivial, /*need_vectors=*/false));
  if (trivial) {
    fill_empty(nq, k, out_ids, out_dist, out_count, stats);
    return ISL_OK;
  }
  if (!idx->encoder) return fail(ISL_INVALID_ARGUMENT, "no recompute encoder attached (isl_index_set_recompute)");
  if (!out_ids || !out_dist) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  const uint32_t n = (uint32_t)idx->n;
  const uint32_t vis_words = round_up((n + 31) / 32, 4);
  if ((uint64_t)nq * vis_words * 4 > (8ull << 30))
    return fail(ISL_INVALID_ARGUMENT, "recompute search: batch too large for the per-query visited sets (split the batch)");
  DeviceGuard g(idx->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  cudaStream_t st = sc->stream;
  const uint32_t u_cap = std::max<uint32_t>(32, round_up(idx->max_degree, 32));
  const uint32_t tie_cap = ef + kXTie;
  const uint32_t S = idx->tok_len;
  DevBuf<uint2> R, ties;
  DevBuf<uint32_t> vis, cand, meta, flags, rows;
  DevBuf<float> qsq, emb, emb_sq;
  DevBuf<int32_t> tok, len;
  DevBuf<uint8_t> scan_tmp;
  DevBuf<unsigned int> active;
  ISL_CUDA_TRY(R.alloc(nq * ef));
  ISL_CUDA_TRY(ties.alloc(nq * tie_cap));
  ISL_CUDA_TRY(vis.alloc(nq * vis_words));
  ISL_CUDA_TRY(cand.alloc(nq * u_cap));
  ISL_CUDA_TRY(meta.alloc(nq * 8));
  ISL_CUDA_TRY(flags.alloc((size_t)n + 1));
  ISL_CUDA_TRY(rows.alloc((size_t)n + 1));
  ISL_CUDA_TRY(qsq.alloc(nq));
  ISL_CUDA_TRY(active.alloc(1));
  ISL_TRY(ensure(sc->q_stage, nq * idx->ld));
  ISL_TRY(ensure(sc->out_ids, nq * k));
  ISL_TRY(ensure(sc->out_dist, nq * k));
  ISL_TRY(ensure(sc->out_count, nq));
  ISL_TRY(ensure(sc->out_stats, nq));
  ISL_CUDA_TRY(cudaMemsetAsync(sc->q_stage.p, 0, nq * idx->ld * 4, st));
  ISL_CUDA_TRY(cudaMemcpy2DAsync(sc->q_stage.p, (size_t)idx->ld * 4, queries, (size_t)idx->dim * 4, (size_t)idx->dim * 4, nq,
                                 cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemsetAsync(vis.p, 0, vis.bytes(), st));
  ISL_CUDA_TRY(cudaMemsetAsync(meta.p, 0, meta.bytes(), st));
  ISL_CUDA_TRY(cudaMemsetAsync(sc->out_stats.p, 0, nq * sizeof(isl_search_stats), st));
  ISL_CUDA_TRY(cudaMemsetAsync(active.p, 0, sizeof(unsigned int), st));
  query_sqnorm_kernel<<<(uint32_t)((nq + 127) / 128), 128, 0, st>>>(sc->q_stage.p, idx->ld, idx->dim, (uint32_t)nq, qsq.p);
  count_launch();

  HopArgs a{};
  a.offsets = idx->adj_stride ? nullptr : idx->offsets.p;
  a.nbrs = idx->adj_stride ? idx->adj_pad.p : idx->nbrs.p;
  a.adj_stride = idx->adj_stride;
  a.n = n;
  a.d = idx->dim;
  a.ld = idx->ld;
  a.entry = (uint32_t)idx->entry;
  a.metric = idx->cfg.metric;
  a.prune_ratio = idx->cfg.prune_ratio;
  a.strategy = idx->cfg.pruning_strategy;
  a.prune_seed = idx->cfg.prune_seed;
  a.deg_counts = idx->deg_counts.p;
  a.queries = sc->q_stage.p;
  a.q_ld = idx->ld;
  a.nq = (uint32_t)nq;
  a.ef = ef;
  a.k = k;
  a.u_cap = u_cap;
  a.vis_words = vis_words;
  a.tie_cap = tie_cap;
  a.R = R.p;
  a.ties = ties.p;
  a.vis = vis.p;
  a.cand = cand.p;
  a.meta = meta.p;
  a.q_sqnorm = qsq.p;
  a.stats = sc->out_stats.p;
  a.active = active.p;
  a.row_of_id = rows.p;
  const uint32_t qblocks = (uint32_t)((nq + 3) / 4);
  uint64_t rounds = 0, recomputed = 0;
  float encoder_ms = 0.0f;
  size_t scan_bytes = 0;
  ISL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags.p, rows.p, (int)(n + 1), st));
  ISL_CUDA_TRY(scan_tmp.alloc(scan_bytes + 16));
  const int acc = acc_kind_of_metric(idx->cfg.metric);
  for (;;) {
    rc_expand_kernel<<<qblocks, 128, 0, st>>>(a);
    ISL_CUDA_TRY(cudaMemsetAsync(flags.p, 0, ((size_t)n + 1) * 4, st));
    rc_mark_kernel<<<qblocks, 128, 0, st>>>(cand.p, meta.p, (uint32_t)nq, u_cap, flags.p);
    ISL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, flags.p, rows.p, (int)(n + 1), st));
    count_launch(4);
    uint32_t unique = 0;
    unsigned int still = 0;
    ISL_CUDA_TRY(cudaMemcpyAsync(&unique, rows.p + n, 4, cudaMemcpyDeviceToHost, st));
    ISL_CUDA_TRY(cudaMemcpyAsync(&still, active.p, 4, cudaMemcpyDeviceToHost, st));
    ISL_CUDA_TRY(cudaStreamSynchronize(st));
    if (unique == 0) {
      if (still == 0) break;  // every query has drained its candidates
      continue;               // (cannot happen: an active query always has something to compute)
    }
    rounds++;
    recomputed += unique;
    if (tok.n < (size_t)unique * S) ISL_CUDA_TRY(tok.alloc((size_t)unique * S * 2));
    if (len.n < unique) ISL_CUDA_TRY(len.alloc((size_t)unique * 2));
    if (emb.n < (size_t)unique * idx->ld) ISL_CUDA_TRY(emb.alloc((size_t)unique * idx->ld * 2));
    if (emb_sq.n < unique) ISL_CUDA_TRY(emb_sq.alloc((size_t)unique * 2));
    rc_gather_tokens_kernel<<<1184, 256, 0, st>>>(flags.p, rows.p, n, idx->node_tokens.p, idx->node_lengths.p, S, tok.p, len.p);
    count_launch();
    ISL_CUDA_TRY(cudaGetLastError());
    ISL_CUDA_TRY(cudaStreamSynchronize(st));  // the encoder runs on its own stream
    ISL_TRY(isl_encoder_embed_dev(idx->encoder, tok.p, len.p, unique, S, emb.p));  // provider.compute_embeddings_batch (leann.rs:947-950)
    float ms = 0.0f;
    isl_encoder_last_timing(idx->encoder, &ms, nullptr);
    encoder_ms += ms;
    ISL_TRY(launch_row_sqnorms(emb.p, unique, idx->dim, idx->ld, emb_sq.p, idx->sms, st));
    a.emb = emb.p;
    a.emb_sq = emb_sq.p;
    if (acc == ACC_L2)
      rc_admit_kernel<ACC_L2><<<qblocks, 128, 0, st>>>(a);
    else if (acc == ACC_L1)
      rc_admit_kernel<ACC_L1><<<qblocks, 128, 0, st>>>(a);
    else
      rc_admit_kernel<ACC_DOT><<<qblocks, 128, 0, st>>>(a);
    count_launch();
    ISL_CUDA_TRY(cudaGetLastError());
  }
  rc_results_kernel<<<(uint32_t)nq, 32, 0, st>>>(R.p, meta.p, (uint32_t)nq, ef, k, sc->out_ids.p, sc->out_dist.p, sc->out_count.p);
  count_launch();
  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, sc->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, sc->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, sc->out_count.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (stats) ISL_CUDA_TRY(cudaMemcpyAsync(stats, sc->out_stats.p, nq * sizeof(isl_search_stats), cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  {
    std::lock_guard<std::mutex> tl(idx->pool_mu);
    idx->last_recomputed = recomputed;
    idx->last_encoder_ms = encoder_ms;
    idx->last_traverse_ms = 0.0f;
    idx->last_rerank_ms = 0.0f;
    idx->last_launches = rounds;
  }
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
