// build.cu — K7: LEANN graph construction on the GPU (LeannIndex::build, leann.rs:560-631).
//
// The reference inserts nodes strictly one at a time.  Here nodes are inserted in ROUNDS: all
// nodes of a round search the graph as it stood before the round (the same best-first kernel
// as queries, ef = ef_construction, leann.rs:672-678), pick their neighbours (hub-preserving
// selection, leann.rs:761-833, or truncation to m0, :685), and then their reverse edges are
// applied per target node in ascending source id with the reference's prune-to-m0-closest
// rule (leann.rs:593-607, 634-658).  round size = min(batch, max(1, inserted/2)); batch = 1
// is exactly the reference's sequential loop.  oracle/orc_leann_build_batched restates the
// same round model on the CPU and the two are compared bit for bit.
//
// Device layout while building: fixed-stride adjacency adj[n][m0] (u32) with the distance of
// every edge cached beside it (adj_dist[n][m0]); the reference recomputes those distances in
// prune_neighbors_temp, which yields the same bits because every metric of distance.rs is
// bitwise symmetric in its arguments (x*y, (x-y)^2, |x-y| and sqrt(na*nb) all commute).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <memory>

#include "api_common.h"

namespace isl {
namespace {

constexpr uint32_t kNone32 = 0xffffffffu;

// One warp per new node: neighbour selection + reverse-edge records.
// cand_ids/cand_dist: [round][efc] sorted by (dist,id); cand_cnt [round].
__global__ void __launch_bounds__(128)
select_neighbors_kernel(const uint32_t* __restrict__ cand_ids, const float* __restrict__ cand_dist,
                        const uint32_t* __restrict__ cand_cnt, uint32_t efc, uint32_t round,
                        uint32_t first_id, uint32_t m0, int hub_pruning, float hub_percentile,
                        uint32_t* __restrict__ adj, float* __restrict__ adj_dist,
                        uint32_t* __restrict__ deg, uint64_t* __restrict__ edge_keys,
                        float* __restrict__ edge_vals) {
  extern __shared__ uint32_t smem_sel[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w = blockIdx.x * (blockDim.x >> 5) + warp;
  if (w >= round) return;
  uint32_t* s_deg = smem_sel + (size_t)warp * efc;
  const uint32_t v = first_id + w;
  const uint32_t c = cand_cnt[w];
  const uint32_t* ids = cand_ids + (size_t)w * efc;
  const float* dist = cand_dist + (size_t)w * efc;
  uint32_t* out_ids = adj + (size_t)v * m0;
  float* out_d = adj_dist + (size_t)v * m0;
  uint64_t* ek = edge_keys + (size_t)w * m0;
  float* ev = edge_vals + (size_t)w * m0;
  for (uint32_t i = lane; i < m0; i += 32) ek[i] = ~0ull;
  __syncwarp();

  auto emit = [&](uint32_t pos, uint32_t id, float d) {
    out_ids[pos] = id;
    out_d[pos] = d;
    ek[pos] = ((uint64_t)id << 32) | v;
    ev[pos] = d;
  };

  if (c <= m0 || !hub_pruning) {  // leann.rs:767-769 / :685
    const uint32_t take = c < m0 ? c : m0;
    for (uint32_t i = lane; i < take; i += 32) emit(i, ids[i], dist[i]);
    if (lane == 0) deg[v] = take;
    return;
  }
  // degrees of the candidates in the snapshot (leann.rs:774-777)
  for (uint32_t i = lane; i < c; i += 32) s_deg[i] = deg[ids[i]];
  __syncwarp();
  // threshold = hub_count-th largest degree (leann.rs:778-785)
  const uint32_t hub_count = (uint32_t)ceilf(__fmul_rn((float)c, hub_percentile));
  uint32_t thr = kNone32;
  if (hub_count > 0 && hub_count < c) {
    uint32_t found = kNone32;
    for (uint32_t i = lane; i < c; i += 32) {
      const uint32_t x = s_deg[i];
      uint32_t gt = 0, ge = 0;
      for (uint32_t j = 0; j < c; ++j) {
        gt += s_deg[j] > x;
        ge += s_deg[j] >= x;
      }
      if (gt < hub_count && hub_count <= ge) found = x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, off));
    thr = found;  // every lane that found it found the same value
  }
  // classify in candidate order (leann.rs:790-797); positions via prefix counts
  uint32_t n_hub = 0, n_reg = 0;
  const uint32_t hub_slots = max(m0 / 4, 1u);  // leann.rs:807
  // pass 1: counts
  for (uint32_t b = 0; b < c; b += 32) {
    const uint32_t i = b + lane;
    const bool is_hub = i < c && thr != kNone32 && s_deg[i] >= thr;
    const uint32_t hb = __ballot_sync(0xffffffffu, is_hub);
    const uint32_t vb = __ballot_sync(0xffffffffu, i < c);
    n_hub += __popc(hb);
    n_reg += __popc(vb & ~hb);
  }
  const uint32_t n_h1 = n_hub < hub_slots ? n_hub : hub_slots;
  // pass 2: placement
  uint32_t reg_before = 0;
  for (uint32_t b = 0; b < c; b += 32) {
    const uint32_t i = b + lane;
    const bool valid = i < c;
    const bool is_hub = valid && thr != kNone32 && s_deg[i] >= thr;
    const uint32_t hb = __ballot_sync(0xffffffffu, is_hub);
    const uint32_t vb = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      uint32_t pos;
      if (is_hub) {
        // stable sort of hubs by degree descending (leann.rs:800): rank among hubs
        const uint32_t x = s_deg[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < c; ++j) {
          const uint32_t y = s_deg[j];
          if (y >= thr && (y > x || (y == x && j < i))) rank++;
        }
        pos = rank < hub_slots ? rank : n_h1 + n_reg + (rank - hub_slots);
      } else {
        // regular nodes keep (dist,id) order (leann.rs:802: stable sort by distance)
        pos = n_h1 + reg_before + __popc((vb & ~hb) & ((1u << lane) - 1));
      }
      if (pos < m0) emit(pos, ids[i], dist[i]);
    }
    reg_before += __popc(vb & ~hb);
  }
  if (lane == 0) deg[v] = m0;
}

// Marks the first record of every target-node segment in the sorted edge list.
__global__ void segment_heads_kernel(const uint64_t* __restrict__ keys, uint32_t count,
                                     uint32_t* __restrict__ heads, uint32_t* __restrict__ n_heads) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    if (k == ~0ull) continue;
    if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(k >> 32)) heads[atomicAdd(n_heads, 1u)] = i;
  }
}

// One warp per target node u: append its incoming edges in ascending source id, pruning to the
// m0 closest whenever the list overflows (leann.rs:593-607, 634-658).
__global__ void __launch_bounds__(128)
apply_reverse_edges_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ vals,
                           uint32_t count, const uint32_t* __restrict__ heads,
                           const uint32_t* __restrict__ n_heads, uint32_t m0,
                           uint32_t* __restrict__ adj, float* __restrict__ adj_dist,
                           uint32_t* __restrict__ deg, uint8_t* __restrict__ sorted_flag) {
  extern __shared__ uint32_t smem_rev[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warps = blockDim.x >> 5;
  const uint32_t cap = m0 + 1;
  uint32_t* s_id = smem_rev + (size_t)warp * 4 * cap;
  float* s_d = reinterpret_cast<float*>(s_id + cap);
  uint32_t* t_id = s_id + 2 * cap;
  float* t_d = reinterpret_cast<float*>(s_id + 3 * cap);
  const uint32_t nh = *n_heads;
  for (uint32_t h = blockIdx.x * warps + warp; h < nh; h += gridDim.x * warps) {
    uint32_t e = heads[h];
    const uint32_t u = (uint32_t)(keys[e] >> 32);
    uint32_t d = deg[u];
    bool sorted = sorted_flag[u] != 0;
    __syncwarp();
    for (uint32_t i = lane; i < d; i += 32) {
      s_id[i] = adj[(size_t)u * m0 + i];
      s_d[i] = adj_dist[(size_t)u * m0 + i];
    }
    __syncwarp();
    for (; e < count; ++e) {
      const uint64_t k = keys[e];
      if ((uint32_t)(k >> 32) != u || k == ~0ull) break;
      const uint32_t v = (uint32_t)k;
      const float dv = vals[e];
      // `if !adjacency[nid].contains(&id)` (leann.rs:595)
      bool present = false;
      for (uint32_t i = lane; i < d; i += 32) present |= s_id[i] == v;
      if (__any_sync(0xffffffffu, present)) continue;
      if (d < m0) {
        if (lane == 0) {
          s_id[d] = v;
          s_d[d] = dv;
        }
        d++;
        __syncwarp();
        continue;
      }
      if (!sorted) {
        // first overflow: stable sort of all m0+1 entries by distance, keep m0 (leann.rs:642-657)
        if (lane == 0) {
          s_id[d] = v;
          s_d[d] = dv;
        }
        __syncwarp();
        const uint32_t tot = d + 1;
        for (uint32_t i = lane; i < tot; i += 32) {
          const float x = s_d[i];
          // rank = #{strictly closer} + #{equal (or unordered) and earlier}: a stable sort position.
          // Kept as two separate counters: the fused boolean form was observed to drop the tie
          // term under nvcc 12.9 -O3 (two equal distances then collided on one slot).
          uint32_t closer = 0, tie_before = 0;
          for (uint32_t j = 0; j < tot; ++j) {
            const float y = s_d[j];
            closer += (y < x) ? 1u : 0u;
            tie_before += (!(x < y) && !(y < x) && j < i) ? 1u : 0u;
          }
          const uint32_t rank = closer + tie_before;
          t_id[rank] = s_id[i];
          t_d[rank] = x;
        }
        __syncwarp();
        for (uint32_t i = lane; i < m0; i += 32) {
          s_id[i] = t_id[i];
          s_d[i] = t_d[i];
        }
        sorted = true;
        __syncwarp();
      } else {
        // list already sorted by distance: the new entry goes after every entry that is not
        // greater (stable), and the last one is dropped
        uint32_t le = 0;
        for (uint32_t i = lane; i < d; i += 32) le += !(dv < s_d[i]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) le += __shfl_xor_sync(0xffffffffu, le, off);
        const uint32_t pos = le;
        if (pos < m0) {
          for (int t = (int)m0 - 1; t > (int)pos; t -= 32) {
            const int i = t - (int)lane;
            const bool act = i > (int)pos;
            uint32_t xi = 0;
            float xd = 0.0f;
            if (act) {
              xi = s_id[i - 1];
              xd = s_d[i - 1];
            }
            __syncwarp();
            if (act) {
              s_id[i] = xi;
              s_d[i] = xd;
            }
            __syncwarp();
          }
          if (lane == 0) {
            s_id[pos] = v;
            s_d[pos] = dv;
          }
          __syncwarp();
        }
      }
    }
    for (uint32_t i = lane; i < d; i += 32) {
      adj[(size_t)u * m0 + i] = s_id[i];
      adj_dist[(size_t)u * m0 + i] = s_d[i];
    }
    if (lane == 0) {
      deg[u] = d;
      sorted_flag[u] = sorted ? 1 : 0;
    }
    __syncwarp();
  }
}

__global__ void compact_csr_kernel(const uint32_t* __restrict__ adj, const uint32_t* __restrict__ deg,
                                   const uint64_t* __restrict__ offsets, uint32_t m0, uint32_t n,
                                   uint32_t* __restrict__ nbrs) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t u = warp; u < n; u += warps) {
    const uint64_t o = offsets[u];
    const uint32_t d = deg[u];
    for (uint32_t i = lane; i < d; i += 32) nbrs[o + i] = adj[(size_t)u * m0 + i];
  }
}

__global__ void widen_deg_kernel(const uint32_t* __restrict__ deg, uint64_t* __restrict__ out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = deg[i];
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// LeannIndex::random_level (leann.rs:549-554) with a counter-based uniform in place of thread_rng.
inline uint64_t draw_level(uint64_t seed, uint64_t i, double ml, uint64_t max_layers) {
  uint64_t r = splitmix64(seed + i) >> 11;
  if (r == 0) r = 1;
  const double u = (double)r * (1.0 / 9007199254740992.0);
  const double lv = std::floor(-std::log(u) * ml);
  uint64_t level = lv >= 0.0 ? (lv >= 1.8446744073709552e19 ? ~0ull : (uint64_t)lv) : 0;
  return std::min<uint64_t>(level, max_layers - 1);
}

// Adds the per-query traversal counters of one round to the running totals [n_hop, n_edge, n_dist].
__global__ void accumulate_stats_kernel(const isl_search_stats* __restrict__ st, uint32_t count, unsigned long long* __restrict__ tot) {
  unsigned long long h = 0, e = 0, d = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    h += st[i].n_hop;
    e += st[i].n_edge;
    d += st[i].n_dist;
  }
  for (int off = 16; off > 0; off >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, off);
    e += __shfl_xor_sync(0xffffffffu, e, off);
    d += __shfl_xor_sync(0xffffffffu, d, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(tot + 0, h);
    atomicAdd(tot + 1, e);
    atomicAdd(tot + 2, d);
  }
}

}  // namespace

// Builds the graph for idx (vectors / sqnorms already resident) and fills host + device CSR.
isl_status build_graph(isl_index* idx, const uint64_t* levels_or_null, uint64_t seed, uint32_t batch) {
  const uint64_t n = idx->n;
  const uint32_t m0 = (uint32_t)idx->cfg.m0;
  const uint32_t efc = (uint32_t)idx->cfg.ef_construction;
  if (batch == 0) batch = 1;
  if (idx->cfg.max_layers == 0) return fail(ISL_INVALID_CONFIG, "max_layers must be > 0");
  if (m0 > 1024) return fail(ISL_INVALID_CONFIG, "m0 > 1024 is not supported by the GPU build");
  cudaStream_t st = idx->stream;

  idx->h_levels.resize(n);
  for (uint64_t i = 0; i < n; ++i)
    idx->h_levels[i] = levels_or_null ? levels_or_null[i] : draw_level(seed, i, idx->cfg.ml, idx->cfg.max_layers);

  DevBuf<uint32_t> adj, deg, cand_ids, cand_cnt, heads, n_heads;
  DevBuf<float> adj_dist, cand_dist, edge_vals, edge_vals2;
  DevBuf<uint64_t> edge_keys, edge_keys2;
  DevBuf<uint8_t> sorted_flag, cub_tmp;
  DevBuf<isl_search_stats> round_stats;   // traversal counters of the round's searches (construction roofline)
  DevBuf<unsigned long long> stat_totals; // [n_hop, n_edge, n_dist] over the whole construction
  ISL_CUDA_TRY(adj.alloc(n * m0));
  ISL_CUDA_TRY(adj_dist.alloc(n * m0));
  ISL_CUDA_TRY(deg.alloc(n));
  ISL_CUDA_TRY(sorted_flag.alloc(n));
  ISL_CUDA_TRY(cudaMemsetAsync(deg.p, 0, deg.bytes(), st));
  ISL_CUDA_TRY(cudaMemsetAsync(sorted_flag.p, 0, sorted_flag.bytes(), st));
  const uint64_t max_round = std::min<uint64_t>(batch, std::max<uint64_t>(1, n));
  ISL_CUDA_TRY(cand_ids.alloc(max_round * efc));
  ISL_CUDA_TRY(cand_dist.alloc(max_round * efc));
  ISL_CUDA_TRY(cand_cnt.alloc(max_round));
  ISL_CUDA_TRY(edge_keys.alloc(max_round * m0));
  ISL_CUDA_TRY(edge_keys2.alloc(max_round * m0));
  ISL_CUDA_TRY(edge_vals.alloc(max_round * m0));
  ISL_CUDA_TRY(edge_vals2.alloc(max_round * m0));
  ISL_CUDA_TRY(heads.alloc(max_round * m0));
  ISL_CUDA_TRY(n_heads.alloc(1));
  ISL_CUDA_TRY(round_stats.alloc(max_round));
  ISL_CUDA_TRY(stat_totals.alloc(4));
  ISL_CUDA_TRY(cudaMemsetAsync(stat_totals.p, 0, stat_totals.bytes(), st));
  std::vector<cudaEvent_t> round_ev;  // start / stop of every round's search launch
  struct EvGuard {
    std::vector<cudaEvent_t>* v;
    ~EvGuard() {
      for (cudaEvent_t e : *v) cudaEventDestroy(e);
    }
  } ev_guard{&round_ev};
  ISL_CUDA_TRY(cudaEventRecord(idx->ev0, st));
  size_t cub_bytes = 0;
  ISL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, edge_keys.p, edge_keys2.p, edge_vals.p,
                                               edge_vals2.p, (int)(max_round * m0), 0, 64, st));
  ISL_CUDA_TRY(cub_tmp.alloc(cub_bytes + 16));

  // search scratch
  SearchPlan plan;
  const uint32_t u_cap = std::max<uint32_t>(32, round_up(m0, 32));
  ISL_TRY(plan_search(idx->cfg.metric, idx->ld, efc, u_cap, idx->sms, &plan));
  const uint32_t vis_words = round_up((uint32_t)((n + 31) / 32), 4);
  const uint32_t slots = (uint32_t)std::min<uint64_t>(plan.grid, max_round);
  ISL_TRY(ensure(idx->visited, (size_t)slots * vis_words));
  if (!plan.r_in_smem) ISL_TRY(ensure(idx->r_global, (size_t)slots * efc));
  ISL_TRY(ensure(idx->ties_global, (size_t)slots * efc));

  ISL_CUDA_TRY(cudaMemsetAsync(idx->counters.p, 0, 4 * sizeof(unsigned int), st));
  int64_t entry = ISL_NO_ENTRY;
  uint64_t max_level = 0;
  uint64_t s = 0;
  const uint32_t sel_warps = 4;
  while (s < n) {
    const uint64_t round = std::min<uint64_t>(batch, std::max<uint64_t>(1, s / 2));
    const uint64_t e = std::min<uint64_t>(n, s + round);
    const uint32_t r = (uint32_t)(e - s);
    if (s > 0) {
      ISL_CUDA_TRY(cudaMemsetAsync(idx->counters.p, 0, sizeof(unsigned int), st));  // work counter only
      SearchArgs a{};
      a.vectors = idx->vectors.p;
      a.sqnorms = idx->sqnorms.p;
      a.ld = idx->ld;
      a.d = idx->dim;
      a.n = (uint32_t)n;
      a.offsets = nullptr;
      a.nbrs = adj.p;
      a.degrees = deg.p;
      a.adj_stride = m0;
      a.queries = idx->vectors.p + s * idx->ld;
      a.q_ld = idx->ld;
      a.nq = r;
      a.entry = (uint32_t)(entry >= 0 ? entry : 0);  // leann.rs:669
      a.k = efc;
      a.ef = efc;
      a.metric = idx->cfg.metric;
      a.prune_ratio = 0.0f;  // search_layer_with_adjacency applies no frontier pruning
      a.strategy = 0;
      a.visited = idx->visited.p;
      a.vis_words = vis_words;
      a.r_global = idx->r_global.p;
      a.ties_global = idx->ties_global.p;
      a.u_cap = u_cap;
      a.out_ids = nullptr;
      a.out_ids32 = cand_ids.p;
      a.out_dist = cand_dist.p;
      a.out_count = cand_cnt.p;
      a.stats = round_stats.p;
      a.work_counter = idx->counters.p;
      a.error_flag = idx->counters.p + 1;
      cudaEvent_t e0, e1;
      ISL_CUDA_TRY(cudaEventCreate(&e0));
      round_ev.push_back(e0);
      ISL_CUDA_TRY(cudaEventCreate(&e1));
      round_ev.push_back(e1);
      ISL_CUDA_TRY(cudaEventRecord(e0, st));
      ISL_TRY(launch_search(plan, a, st));
      ISL_CUDA_TRY(cudaEventRecord(e1, st));
      accumulate_stats_kernel<<<std::min<uint32_t>((r + 255) / 256, 148), 256, 0, st>>>(round_stats.p, r, stat_totals.p);
      count_launch();

      const uint32_t blocks = (r + sel_warps - 1) / sel_warps;
      select_neighbors_kernel<<<blocks, sel_warps * 32, (size_t)sel_warps * efc * 4, st>>>(
          cand_ids.p, cand_dist.p, cand_cnt.p, efc, r, (uint32_t)s, m0, idx->cfg.high_degree_pruning,
          idx->cfg.hub_percentile, adj.p, adj_dist.p, deg.p, edge_keys.p, edge_vals.p);
      count_launch();
      const uint32_t n_edges = r * m0;
      size_t tmp_bytes = cub_bytes;
      ISL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, edge_keys.p, edge_keys2.p, edge_vals.p,
                                                   edge_vals2.p, (int)n_edges, 0, 64, st));
      count_launch(3);
      ISL_CUDA_TRY(cudaMemsetAsync(n_heads.p, 0, 4, st));
      segment_heads_kernel<<<std::min<uint32_t>((n_edges + 255) / 256, 1184), 256, 0, st>>>(
          edge_keys2.p, n_edges, heads.p, n_heads.p);
      count_launch();
      const uint32_t rev_blocks = std::min<uint32_t>((n_edges + sel_warps - 1) / sel_warps, (uint32_t)idx->sms * 8);
      apply_reverse_edges_kernel<<<rev_blocks, sel_warps * 32, (size_t)sel_warps * 4 * (m0 + 1) * 4, st>>>(
          edge_keys2.p, edge_vals2.p, n_edges, heads.p, n_heads.p, m0, adj.p, adj_dist.p, deg.p, sorted_flag.p);
      count_launch();
      ISL_CUDA_TRY(cudaGetLastError());
    }
    for (uint64_t id = s; id < e; ++id) {  // leann.rs:610-613
      const uint64_t level = idx->h_levels[id];
      if (entry < 0 || level > max_level) {
        entry = (int64_t)id;
        max_level = level;
      }
    }
    s = e;
  }
  unsigned int hflags[4] = {0, 0, 0, 0};
  unsigned long long htot[4] = {0, 0, 0, 0};
  ISL_CUDA_TRY(cudaMemcpyAsync(hflags, idx->counters.p, sizeof(hflags), cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(htot, stat_totals.p, sizeof(htot), cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaEventRecord(idx->ev1, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  idx->build_stats = isl_build_stats{};
  idx->build_stats.n_hop = htot[0];
  idx->build_stats.n_edge = htot[1];
  idx->build_stats.n_dist = htot[2];
  idx->build_stats.rounds = round_ev.size() / 2;
  cudaEventElapsedTime(&idx->build_stats.rounds_ms, idx->ev0, idx->ev1);
  for (size_t i = 0; i + 1 < round_ev.size(); i += 2) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, round_ev[i], round_ev[i + 1]) == cudaSuccess) idx->build_stats.search_ms += ms;
  }

  // adjacency -> CSR (leann.rs:618-627)
  DevBuf<uint64_t> deg64;
  ISL_CUDA_TRY(deg64.alloc(n + 1));
  ISL_CUDA_TRY(idx->offsets.alloc(n + 1));
  ISL_CUDA_TRY(cudaMemsetAsync(deg64.p, 0, deg64.bytes(), st));
  widen_deg_kernel<<<std::min<uint32_t>((uint32_t)((n + 255) / 256), 1184), 256, 0, st>>>(deg.p, deg64.p, (uint32_t)n);
  count_launch();
  size_t scan_bytes = 0;
  ISL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, deg64.p, idx->offsets.p, (int)(n + 1), st));
  DevBuf<uint8_t> scan_tmp;
  ISL_CUDA_TRY(scan_tmp.alloc(scan_bytes + 16));
  ISL_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, deg64.p, idx->offsets.p, (int)(n + 1), st));
  count_launch(2);
  idx->h_offsets.resize(n + 1);
  ISL_CUDA_TRY(cudaMemcpyAsync(idx->h_offsets.data(), idx->offsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  const uint64_t num_edges = idx->h_offsets[n];
  ISL_CUDA_TRY(idx->nbrs.alloc(std::max<uint64_t>(num_edges, 1)));
  compact_csr_kernel<<<(uint32_t)std::min<uint64_t>((n + 7) / 8, 148 * 8), 256, 0, st>>>(
      adj.p, deg.p, idx->offsets.p, m0, (uint32_t)n, idx->nbrs.p);
  count_launch();
  std::vector<uint32_t> n32(num_edges);
  if (num_edges)
    ISL_CUDA_TRY(cudaMemcpyAsync(n32.data(), idx->nbrs.p, num_edges * 4, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  idx->h_nbrs.resize(num_edges);
  for (uint64_t i = 0; i < num_edges; ++i) idx->h_nbrs[i] = n32[i];
  uint32_t maxdeg = 0;
  for (uint64_t i = 0; i < n; ++i)
    maxdeg = std::max<uint32_t>(maxdeg, (uint32_t)(idx->h_offsets[i + 1] - idx->h_offsets[i]));
  idx->max_degree = maxdeg;
  idx->entry = entry;
  idx->max_level = max_level;
  idx->build_stats.edges = num_edges;
  ISL_TRY(index_make_padded_adjacency(idx));
  // construction scratch is not needed by searches (they lease their own)
  idx->visited.release();
  idx->r_global.release();
  idx->ties_global.release();
  if (hflags[1]) return fail(ISL_CUDA_ERROR, "build: internal invariant violated (tie list overflow)");
  return ISL_OK;
}

}  // namespace isl

using namespace isl;

extern "C" {

static isl_status build_common(const isl_leann_config* cfg, uint32_t dim, uint64_t n, const float* vectors,
                               bool on_device, const uint64_t* levels_or_null, uint64_t seed, uint32_t batch,
                               isl_index** out) {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  ISL_TRY(isl_leann_config_validate(cfg));
  if (n >= (1ull << 31)) return fail(ISL_INVALID_ARGUMENT, "n must be < 2^31 per index (shard larger sets)");
  if (n > 0 && (!vectors || dim == 0)) return fail(ISL_INVALID_ARGUMENT, "vectors is null or dim is 0");
  std::unique_ptr<isl_index> idx(new isl_index());
  idx->cfg = *cfg;
  idx->n = n;
  idx->dim = n ? dim : 0;  // LeannIndex::build returns early for 0 vectors (leann.rs:565-567)
  idx->ld = n ? std::max<uint32_t>(4, round_up(dim, 4)) : 0;
  ISL_TRY(index_alloc_common(idx.get()));
  if (n == 0) {
    idx->h_offsets.assign(1, 0);
    *out = idx.release();
    return ISL_OK;
  }
  ISL_CUDA_TRY(idx->vectors.alloc(n * idx->ld));
  if (on_device) {
    ISL_TRY(launch_pad_rows(vectors, dim, idx->vectors.p, idx->ld, n, idx->stream));
  } else {
    if (idx->ld != dim) ISL_CUDA_TRY(cudaMemsetAsync(idx->vectors.p, 0, idx->vectors.bytes(), idx->stream));
    ISL_CUDA_TRY(cudaMemcpy2DAsync(idx->vectors.p, (size_t)idx->ld * 4, vectors, (size_t)dim * 4, (size_t)dim * 4,
                                   n, cudaMemcpyHostToDevice, idx->stream));
  }
  ISL_CUDA_TRY(idx->sqnorms.alloc(n));
  ISL_TRY(launch_row_sqnorms(idx->vectors.p, n, dim, idx->ld, idx->sqnorms.p, idx->sms, idx->stream));
  ISL_TRY(build_graph(idx.get(), levels_or_null, seed, batch));
  *out = idx.release();
  return ISL_OK;
}

isl_status isl_index_last_build_stats(const isl_index* idx, isl_build_stats* out) try {
  if (!idx || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  *out = idx->build_stats;
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_index_build(const isl_leann_config* cfg, uint32_t dim, uint64_t n, const float* vectors,
                           const uint64_t* levels_or_null, uint64_t seed, uint32_t batch, isl_index** out) try {
  return build_common(cfg, dim, n, vectors, false, levels_or_null, seed, batch, out);
} ISL_ABI_GUARD

isl_status isl_index_build_dev(const isl_leann_config* cfg, uint32_t dim, uint64_t n, const float* d_vectors,
                               const uint64_t* levels_or_null, uint64_t seed, uint32_t batch, isl_index** out) try {
  // d_vectors may have been written on another stream: order after the legacy default stream.
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceSynchronize");
  return build_common(cfg, dim, n, d_vectors, true, levels_or_null, seed, batch, out);
} ISL_ABI_GUARD

}  // extern "C"
