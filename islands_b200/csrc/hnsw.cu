// hnsw.cu — multi-layer HNSW on the GPU: HnswGraph::insert / insert_node / prune_connections /
// search (src/core/hnsw.rs:214-329, 332-402, 405-446, 458-504) behind the C ABI (isl_hnsw_*).
//
// Device layout.  Vectors [cap][ld] f32 and their squared norms are resident like the LEANN index.
// Layer 0 keeps one fixed-stride row of m0 ids per node (adj0, with the live degree in deg0 and
// the distance of every edge cached beside it in dist0).  Layers >= 1 are sparse (a node reaches
// layer L with probability m^-L), so their rows live in one compact pool of stride m:
// row(node, L) = row_map[node] + L - 1.  node_levels[node] says which layers a node has: above
// it neighbors_at() is None (hnsw.rs:108-110) and the search kernel sees an empty list.
//
// Search = one greedy-descent kernel over the upper layers (one warp per query, the strict `<`
// scan of hnsw.rs:478-497 over each snapshot list) and then the LEANN best-first kernel
// (search_core.cuh) on layer 0 with a per-query entry point.
//
// Insert is batched in ROUNDS exactly like build.cu: the read-only half of insert_node (descent,
// per-layer search_layer, take(M)) runs for every node of a round against the graph as it stood
// before the round — one kernel launch per layer over the nodes that own that layer — and the
// mutating half (own lists, reverse edges, prune_connections, entry update) is applied afterwards
// per target list in ascending source id.  batch = 1 is the reference's sequential insert;
// oracle/orc_hnsw_insert_batch restates the round model and is compared bit for bit.
//
// prune_connections (hnsw.rs:405-446) quirk kept: the node being inserted is not in `nodes` yet
// (hnsw.rs:327), so when a neighbour's list is already full the new id is filtered out and the
// list is only re-sorted (stable) by distance.  The distances it recomputes are the cached ones:
// every metric of distance.rs is bitwise symmetric in its arguments.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <mutex>
#include <vector>

#include "api_common.h"
#include "bincode.h"
#include "row_stream.cuh"

struct isl_hnsw {
  isl_hnsw_config cfg{};
  uint32_t dim = 0, ld = 0;
  uint64_t n = 0;      // nodes inserted (ids 0..n-1, hnsw.rs:227-228)
  uint64_t cap = 0;    // node capacity of the device arrays
  uint64_t n_upper = 0, cap_upper = 0;  // rows of the upper-layer pool
  int device = 0, sms = 0;
  int64_t entry = ISL_NO_ENTRY;
  uint64_t max_level = 0;
  std::vector<uint32_t> h_levels, h_row_map;
  isl::DevBuf<float> vectors, sqnorms, dist0, distU;
  isl::DevBuf<uint32_t> adj0, deg0, adjU, degU, node_levels, row_map, cur_by_node;
  mutable std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mutable float last_kernel_ms = 0.0f;
  mutable isl::DevBuf<uint32_t> visited, q_cur, out_count;
  mutable isl::DevBuf<uint2> r_global, ties_global;
  mutable isl::DevBuf<unsigned int> counters;
  mutable isl::DevBuf<float> q_stage, out_dist;
  mutable isl::DevBuf<uint64_t> out_ids;
  ~isl_hnsw() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace isl {
namespace {

struct GreedyArgs {
  const float* vectors;
  const float* sqnorms;
  uint32_t ld, d;
  int32_t metric;
  const float* queries;  // [nq][q_ld]
  uint32_t q_ld, nq;
  uint32_t entry, top_layer;          // entry point and max_level of the graph being descended
  const uint32_t* stop_levels;        // null: descend to layer 1 (search); else item i stops at stop_levels[i] + 1 (insert)
  const uint32_t* adjU;
  const uint32_t* degU;
  uint32_t m;
  const uint32_t* row_map;
  const uint32_t* node_levels;
  uint32_t u_cap;
  uint32_t* out_cur;  // [nq]
};

// Greedy descent (hnsw.rs:263-282 for insert, :478-497 for search): per layer, repeat
// { scan the neighbour list `current` had when the pass began, in order, moving to every node
// strictly closer than the running best } until a pass changes nothing.  The scan's result is
// the first position of the list minimum, taken only if it beats the running distance, so the
// warp scores 32 neighbours at a time (one row per lane, reference-order fold) and reduces.
template <int ACC>
__global__ void __launch_bounds__(32) hnsw_greedy_kernel(const GreedyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using G = StageGeom<64>;
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* q_smem = stage + G::STAGE_FLOATS;
  uint32_t* u_list = reinterpret_cast<uint32_t*>(q_smem + a.ld);
  uint64_t* bars = reinterpret_cast<uint64_t*>(u_list + a.u_cap);
  const uint32_t lane = lane_id();

  RowRing<1> ring;
  ring.stage = stage;
  ring.bars = bars;
  ring.phase_bits = 0;
  ring.islot = 0;
  ring.cslot = 0;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_fence_init();
  }
  __syncwarp();

  for (uint32_t qi = blockIdx.x; qi < a.nq; qi += gridDim.x) {
    {
      const float4* src = reinterpret_cast<const float4*>(a.queries + (size_t)qi * a.q_ld);
      float4* dst = reinterpret_cast<float4*>(q_smem);
      for (uint32_t i = lane; i < a.ld / 4; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const float na = (a.metric == ISL_METRIC_COSINE) ? smem_sqnorm_fold(q_smem, a.d) : 0.0f;
    uint32_t cur = a.entry;
    float cur_d = 0.0f;
    bool changed = false;
    bool first = true;
    auto on_group = [&](uint32_t base, uint32_t cnt, float acc) {
      float dn = __int_as_float(0x7f800000);
      if (lane < cnt) {
        const float nb = (a.metric == ISL_METRIC_COSINE) ? __ldg(a.sqnorms + u_list[base + lane]) : 0.0f;
        dn = finalize_distance(a.metric, acc, na, nb);
      }
      if (first) {  // distance to the entry point (hnsw.rs:260, :475): taken as is, NaN included
        cur_d = __shfl_sync(0xffffffffu, dn, 0);
        return;
      }
      if (dn != dn) dn = __int_as_float(0x7f800000);  // `NaN < x` is false: can never be taken
      float best = dn;
      uint32_t bi = lane;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, off);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (od < best || (od == best && oi < bi)) {
          best = od;
          bi = oi;
        }
      }
      if (best < cur_d) {  // hnsw.rs:271 / :486
        cur_d = best;
        cur = u_list[base + bi];
        changed = true;
      }
    };
    if (lane == 0) u_list[0] = cur;
    __syncwarp();
    stream_rows_fold<ACC, 1>(ring, a.vectors, a.ld, a.d, u_list, 1, q_smem, on_group);
    first = false;
    const uint32_t stop = a.stop_levels ? __ldg(a.stop_levels + qi) + 1 : 1;
    for (uint32_t layer = a.top_layer; layer >= stop && layer >= 1; --layer) {
      for (;;) {
        changed = false;
        uint32_t deg = 0;
        if (__ldg(a.node_levels + cur) >= layer) {
          const uint32_t row = __ldg(a.row_map + cur) + layer - 1;
          deg = __ldg(a.degU + row);
          __syncwarp();
          for (uint32_t i = lane; i < deg; i += 32) u_list[i] = __ldg(a.adjU + (size_t)row * a.m + i);
          __syncwarp();
        }
        stream_rows_fold<ACC, 1>(ring, a.vectors, a.ld, a.d, u_list, deg, q_smem, on_group);
        if (!changed) break;
      }
    }
    if (lane == 0) a.out_cur[qi] = cur;
    __syncwarp();
  }
}

// One warp per (new node, layer): selected = first M of the search result (hnsw.rs:290-295),
// the node's own list (hnsw.rs:298-300), one reverse-edge record per selected neighbour that owns
// the layer (hnsw.rs:303-305), and the entry for the next layer (hnsw.rs:316-318).
__global__ void __launch_bounds__(128)
hnsw_select_kernel(const uint32_t* __restrict__ cand_ids, const float* __restrict__ cand_dist,
                   const uint32_t* __restrict__ cand_cnt, uint32_t efc, uint32_t items,
                   const uint32_t* __restrict__ item_nodes, uint32_t layer, uint32_t cap_conn,
                   uint32_t* __restrict__ adj, float* __restrict__ adj_dist, uint32_t* __restrict__ deg,
                   const uint32_t* __restrict__ row_map, const uint32_t* __restrict__ node_levels,
                   uint32_t upper_base, uint64_t* __restrict__ edge_keys, float* __restrict__ edge_vals,
                   uint32_t* __restrict__ cur_by_node) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w = blockIdx.x * (blockDim.x >> 5) + warp;
  if (w >= items) return;
  const uint32_t v = item_nodes[w];
  const uint32_t c = cand_cnt[w];
  const uint32_t take = c < cap_conn ? c : cap_conn;
  const uint32_t row = layer == 0 ? v : row_map[v] + layer - 1;
  for (uint32_t i = lane; i < cap_conn; i += 32) {
    uint64_t key = ~0ull;
    float d = 0.0f;
    if (i < take) {
      const uint32_t id = cand_ids[(size_t)w * efc + i];
      d = cand_dist[(size_t)w * efc + i];
      adj[(size_t)row * cap_conn + i] = id;
      adj_dist[(size_t)row * cap_conn + i] = d;
      if (node_levels[id] >= layer) {
        const uint32_t trow = layer == 0 ? id : upper_base + row_map[id] + layer - 1;
        key = ((uint64_t)trow << 32) | v;
      }
    }
    edge_keys[(size_t)w * cap_conn + i] = key;
    edge_vals[(size_t)w * cap_conn + i] = d;
  }
  if (lane == 0) {
    deg[row] = take;
    if (take > 0) cur_by_node[v] = cand_ids[(size_t)w * efc];
  }
}

__global__ void hnsw_segment_heads_kernel(const uint64_t* __restrict__ keys, uint32_t count,
                                          uint32_t* __restrict__ heads, uint32_t* __restrict__ n_heads) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    if (k == ~0ull) continue;
    if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(k >> 32)) heads[atomicAdd(n_heads, 1u)] = i;
  }
}

// One warp per target list: incoming edges in ascending source id are appended while the list is
// below its capacity (hnsw.rs:305); the first edge that finds it full triggers prune_connections
// (hnsw.rs:405-446), which drops the id being inserted and leaves the list stable-sorted by
// distance — and so does every later one, with no further effect.
__global__ void __launch_bounds__(128)
hnsw_apply_reverse_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ vals, uint32_t count,
                          const uint32_t* __restrict__ heads, const uint32_t* __restrict__ n_heads,
                          uint32_t upper_base, uint32_t m0, uint32_t m, uint32_t* __restrict__ adj0,
                          float* __restrict__ dist0, uint32_t* __restrict__ deg0, uint32_t* __restrict__ adjU,
                          float* __restrict__ distU, uint32_t* __restrict__ degU) {
  extern __shared__ uint32_t smem_rev[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warps = blockDim.x >> 5;
  const uint32_t cmax = m0 > m ? m0 : m;
  uint32_t* s_id = smem_rev + (size_t)warp * 4 * cmax;
  float* s_d = reinterpret_cast<float*>(s_id + cmax);
  uint32_t* t_id = s_id + 2 * cmax;
  float* t_d = reinterpret_cast<float*>(s_id + 3 * cmax);
  const uint32_t nh = *n_heads;
  for (uint32_t h = blockIdx.x * warps + warp; h < nh; h += gridDim.x * warps) {
    const uint32_t e0 = heads[h];
    const uint32_t trow = (uint32_t)(keys[e0] >> 32);
    const bool upper = trow >= upper_base;
    const uint32_t row = upper ? trow - upper_base : trow;
    const uint32_t cap_conn = upper ? m : m0;
    uint32_t* adj = (upper ? adjU : adj0) + (size_t)row * cap_conn;
    float* adist = (upper ? distU : dist0) + (size_t)row * cap_conn;
    uint32_t* degp = (upper ? degU : deg0) + row;
    uint32_t d = *degp;
    uint32_t cnt = 0;
    while (e0 + cnt < count && (uint32_t)(keys[e0 + cnt] >> 32) == trow) ++cnt;  // ~0 keys sort last
    const uint32_t room = cap_conn - d;
    const uint32_t app = cnt < room ? cnt : room;
    __syncwarp();
    for (uint32_t i = lane; i < d; i += 32) {
      s_id[i] = adj[i];
      s_d[i] = adist[i];
    }
    for (uint32_t i = lane; i < app; i += 32) {
      s_id[d + i] = (uint32_t)keys[e0 + i];
      s_d[d + i] = vals[e0 + i];
    }
    __syncwarp();
    d += app;
    if (cnt > app) {  // full list + one more: stable sort by distance (ties keep list order)
      for (uint32_t i = lane; i < d; i += 32) {
        const float x = s_d[i];
        uint32_t closer = 0, tie_before = 0;
        for (uint32_t j = 0; j < d; ++j) {
          const float y = s_d[j];
          closer += (y < x) ? 1u : 0u;
          tie_before += (!(x < y) && !(y < x) && j < i) ? 1u : 0u;
        }
        const uint32_t rank = closer + tie_before;
        t_id[rank] = s_id[i];
        t_d[rank] = x;
      }
      __syncwarp();
      for (uint32_t i = lane; i < d; i += 32) {
        adj[i] = t_id[i];
        adist[i] = t_d[i];
      }
    } else {
      for (uint32_t i = lane; i < d; i += 32) {
        adj[i] = s_id[i];
        adist[i] = s_d[i];
      }
    }
    if (lane == 0) *degp = d;
    __syncwarp();
  }
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// HnswGraph::random_level (hnsw.rs:206-211) with a counter-based uniform in place of thread_rng.
inline uint64_t draw_level(uint64_t seed, uint64_t i, double ml, uint64_t max_layers) {
  uint64_t r = splitmix64(seed + i) >> 11;
  if (r == 0) r = 1;
  const double u = (double)r * (1.0 / 9007199254740992.0);
  const double lv = std::floor(-std::log(u) * ml);
  uint64_t level = lv >= 0.0 ? (lv >= 1.8446744073709552e19 ? ~0ull : (uint64_t)lv) : 0;
  return std::min<uint64_t>(level, max_layers - 1);
}

// Grows a device array to `want` elements keeping the first `keep`.
template <class T>
isl_status grow(DevBuf<T>& b, size_t keep, size_t want, cudaStream_t st) {
  if (b.n >= want) return ISL_OK;
  DevBuf<T> nb;
  ISL_CUDA_TRY(nb.alloc(want));
  if (keep) ISL_CUDA_TRY(cudaMemcpyAsync(nb.p, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  b = std::move(nb);
  return ISL_OK;
}

template <int ACC>
isl_status launch_greedy_one(const GreedyArgs& a, uint32_t grid, size_t smem, cudaStream_t st) {
  ISL_CUDA_TRY(cudaFuncSetAttribute(hnsw_greedy_kernel<ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* the maximum, always: never lowered under a concurrent launch */));
  hnsw_greedy_kernel<ACC><<<grid, 32, smem, st>>>(a);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_greedy(const isl_hnsw* h, GreedyArgs a, cudaStream_t st) {
  if (a.nq == 0) return ISL_OK;
  a.vectors = h->vectors.p;
  a.sqnorms = h->sqnorms.p;
  a.ld = h->ld;
  a.d = h->dim;
  a.metric = h->cfg.metric;
  a.adjU = h->adjU.p;
  a.degU = h->degU.p;
  a.m = (uint32_t)h->cfg.m;
  a.row_map = h->row_map.p;
  a.node_levels = h->node_levels.p;
  a.u_cap = std::max<uint32_t>(32, round_up((uint32_t)h->cfg.m, 32));
  const size_t smem = ((size_t)StageGeom<64>::STAGE_FLOATS + a.ld + a.u_cap) * 4 + 16;
  if (smem > 227 * 1024) return fail(ISL_INVALID_ARGUMENT, "hnsw: dimension too large for the descent kernel");
  const uint32_t grid = std::min<uint32_t>(a.nq, (uint32_t)h->sms * 16);
  switch (acc_kind_of_metric(h->cfg.metric)) {
    case ACC_DOT: return launch_greedy_one<ACC_DOT>(a, grid, smem, st);
    case ACC_L2: return launch_greedy_one<ACC_L2>(a, grid, smem, st);
    default: return launch_greedy_one<ACC_L1>(a, grid, smem, st);
  }
}

// Common SearchArgs of a layer search on the HNSW handle.
void fill_layer_args(const isl_hnsw* h, uint32_t layer, SearchArgs* a) {
  a->vectors = h->vectors.p;
  a->sqnorms = h->sqnorms.p;
  a->ld = h->ld;
  a->d = h->dim;
  a->n = (uint32_t)h->n;
  a->offsets = nullptr;
  a->metric = h->cfg.metric;
  a->prune_ratio = 0.0f;
  a->strategy = 0;
  a->node_levels = h->node_levels.p;
  a->layer = layer;
  if (layer == 0) {
    a->nbrs = h->adj0.p;
    a->degrees = h->deg0.p;
    a->adj_stride = (uint32_t)h->cfg.m0;
    a->row_map = nullptr;
    a->row_add = 0;
  } else {
    a->nbrs = h->adjU.p;
    a->degrees = h->degU.p;
    a->adj_stride = (uint32_t)h->cfg.m;
    a->row_map = h->row_map.p;
    a->row_add = layer - 1;
  }
}

// dist[row][i] = metric(vec[node(row)], vec[adj[row][i]]) for every stored edge (one warp per row):
// rebuilds the cached edge distances of a deserialised graph in reference order.
template <int ACC>
__global__ void __launch_bounds__(32)
hnsw_edge_dist_kernel(const float* __restrict__ vectors, const float* __restrict__ sqnorms, uint32_t ld, uint32_t d,
                      int32_t metric, const uint32_t* __restrict__ owner, uint32_t rows, const uint32_t* __restrict__ adj,
                      const uint32_t* __restrict__ deg, uint32_t stride, float* __restrict__ out, uint32_t u_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using G = StageGeom<64>;
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* q_smem = stage + G::STAGE_FLOATS;
  uint32_t* u_list = reinterpret_cast<uint32_t*>(q_smem + ld);
  uint64_t* bars = reinterpret_cast<uint64_t*>(u_list + u_cap);
  const uint32_t lane = lane_id();
  RowRing<1> ring;
  ring.stage = stage;
  ring.bars = bars;
  ring.phase_bits = 0;
  ring.islot = 0;
  ring.cslot = 0;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_fence_init();
  }
  __syncwarp();
  for (uint32_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const uint32_t node = owner ? owner[row] : row;
    const uint32_t dg = deg[row];
    __syncwarp();
    const float4* src = reinterpret_cast<const float4*>(vectors + (size_t)node * ld);
    for (uint32_t i = lane; i < ld / 4; i += 32) reinterpret_cast<float4*>(q_smem)[i] = src[i];
    for (uint32_t i = lane; i < dg; i += 32) u_list[i] = adj[(size_t)row * stride + i];
    __syncwarp();
    const float na = (metric == ISL_METRIC_COSINE) ? smem_sqnorm_fold(q_smem, d) : 0.0f;
    auto on_group = [&](uint32_t base, uint32_t cnt, float acc) {
      if (lane < cnt) {
        const float nb = (metric == ISL_METRIC_COSINE) ? __ldg(sqnorms + u_list[base + lane]) : 0.0f;
        out[(size_t)row * stride + base + lane] = finalize_distance(metric, acc, na, nb);
      }
    };
    stream_rows_fold<ACC, 1>(ring, vectors, ld, d, u_list, dg, q_smem, on_group);
  }
}

template <int ACC>
isl_status edge_dist_one(const isl_hnsw* h, const uint32_t* owner, uint32_t rows, const uint32_t* adj, const uint32_t* deg,
                         uint32_t stride, float* out, cudaStream_t st) {
  if (rows == 0) return ISL_OK;
  const uint32_t u_cap = std::max<uint32_t>(32, round_up(stride, 32));
  const size_t smem = ((size_t)StageGeom<64>::STAGE_FLOATS + h->ld + u_cap) * 4 + 16;
  ISL_CUDA_TRY(cudaFuncSetAttribute(hnsw_edge_dist_kernel<ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 /* the maximum, always: never lowered under a concurrent launch */));
  hnsw_edge_dist_kernel<ACC><<<std::min<uint32_t>(rows, (uint32_t)h->sms * 16), 32, smem, st>>>(
      h->vectors.p, h->sqnorms.p, h->ld, h->dim, h->cfg.metric, owner, rows, adj, deg, stride, out, u_cap);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

isl_status launch_edge_distances(isl_hnsw* h, cudaStream_t st) {
  // upper-pool rows need their owning node: row_map[node] .. + level
  std::vector<uint32_t> owner(h->n_upper);
  for (uint64_t i = 0; i < h->n; ++i)
    for (uint32_t L = 1; L <= h->h_levels[i]; ++L) owner[h->h_row_map[i] + L - 1] = (uint32_t)i;
  DevBuf<uint32_t> d_owner;
  if (h->n_upper) {
    ISL_CUDA_TRY(d_owner.alloc(h->n_upper));
    ISL_CUDA_TRY(cudaMemcpyAsync(d_owner.p, owner.data(), h->n_upper * 4, cudaMemcpyHostToDevice, st));
  }
  const uint32_t m0 = (uint32_t)h->cfg.m0, m = (uint32_t)h->cfg.m;
  isl_status s1, s2;
  switch (acc_kind_of_metric(h->cfg.metric)) {
    case ACC_DOT:
      s1 = edge_dist_one<ACC_DOT>(h, nullptr, (uint32_t)h->n, h->adj0.p, h->deg0.p, m0, h->dist0.p, st);
      s2 = edge_dist_one<ACC_DOT>(h, d_owner.p, (uint32_t)h->n_upper, h->adjU.p, h->degU.p, m, h->distU.p, st);
      break;
    case ACC_L2:
      s1 = edge_dist_one<ACC_L2>(h, nullptr, (uint32_t)h->n, h->adj0.p, h->deg0.p, m0, h->dist0.p, st);
      s2 = edge_dist_one<ACC_L2>(h, d_owner.p, (uint32_t)h->n_upper, h->adjU.p, h->degU.p, m, h->distU.p, st);
      break;
    default:
      s1 = edge_dist_one<ACC_L1>(h, nullptr, (uint32_t)h->n, h->adj0.p, h->deg0.p, m0, h->dist0.p, st);
      s2 = edge_dist_one<ACC_L1>(h, d_owner.p, (uint32_t)h->n_upper, h->adjU.p, h->degU.p, m, h->distU.p, st);
  }
  ISL_TRY(s1);
  ISL_TRY(s2);
  ISL_CUDA_TRY(cudaStreamSynchronize(st));  // d_owner goes out of scope
  return ISL_OK;
}

isl_status hnsw_insert_impl(isl_hnsw* h, const float* vectors, bool on_device, uint64_t count, uint32_t dim,
                            const uint64_t* levels_or_null, uint64_t seed, uint32_t batch, uint64_t* first_id) {
  if (first_id) *first_id = h->n;
  if (count == 0) return ISL_OK;
  if (!vectors) return fail(ISL_INVALID_ARGUMENT, "vectors is null");
  if (dim == 0) return fail(ISL_INVALID_ARGUMENT, "dim must be > 0");
  if (h->dim != 0 && dim != h->dim)  // hnsw.rs:216-223
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(h->dim) + ", got " + std::to_string(dim));
  if (h->cfg.max_layers == 0) return fail(ISL_INVALID_CONFIG, "max_layers must be > 0");
  const uint32_t m0 = (uint32_t)h->cfg.m0, m = (uint32_t)h->cfg.m, efc = (uint32_t)h->cfg.ef_construction;
  if (m0 > 1024 || m > 1024) return fail(ISL_INVALID_CONFIG, "m / m0 > 1024 is not supported by the GPU build");
  const uint64_t first = h->n, total = first + count;
  if (total >= (1ull << 31)) return fail(ISL_INVALID_ARGUMENT, "n must be < 2^31 per graph");
  if (batch == 0) batch = 1;
  cudaStream_t st = h->stream;
  if (h->dim == 0) {  // hnsw.rs:224-226
    h->dim = dim;
    h->ld = std::max<uint32_t>(4, round_up(dim, 4));
  }
  const uint32_t ld = h->ld;

  // levels, upper-pool rows
  h->h_levels.resize(total);
  h->h_row_map.resize(total);
  uint64_t n_upper = h->n_upper;
  for (uint64_t i = first; i < total; ++i) {
    const uint64_t lv = levels_or_null ? levels_or_null[i - first] : draw_level(seed, i, h->cfg.ml, h->cfg.max_layers);
    if (lv > 255) return fail(ISL_INVALID_ARGUMENT, "node level > 255");
    h->h_levels[i] = (uint32_t)lv;
    h->h_row_map[i] = (uint32_t)n_upper;
    n_upper += lv;
  }
  if (total + n_upper >= (1ull << 32)) return fail(ISL_INVALID_ARGUMENT, "graph too large");

  // capacity
  if (total > h->cap) {
    const uint64_t cap = std::max<uint64_t>(total, h->cap * 2);
    ISL_TRY(grow(h->vectors, first * ld, cap * ld, st));
    ISL_TRY(grow(h->sqnorms, first, cap, st));
    ISL_TRY(grow(h->adj0, first * m0, cap * m0, st));
    ISL_TRY(grow(h->dist0, first * m0, cap * m0, st));
    ISL_TRY(grow(h->deg0, first, cap, st));
    ISL_TRY(grow(h->node_levels, first, cap, st));
    ISL_TRY(grow(h->row_map, first, cap, st));
    ISL_TRY(grow(h->cur_by_node, first, cap, st));
    h->cap = cap;
  }
  if (n_upper > h->cap_upper) {
    const uint64_t cap = std::max<uint64_t>(n_upper, h->cap_upper * 2);
    ISL_TRY(grow(h->adjU, h->n_upper * m, cap * m, st));
    ISL_TRY(grow(h->distU, h->n_upper * m, cap * m, st));
    ISL_TRY(grow(h->degU, h->n_upper, cap, st));
    h->cap_upper = cap;
  }
  if (on_device) {
    ISL_TRY(launch_pad_rows(vectors, dim, h->vectors.p + first * ld, ld, count, st));
  } else {
    if (ld != dim) ISL_CUDA_TRY(cudaMemsetAsync(h->vectors.p + first * ld, 0, count * ld * 4, st));
    ISL_CUDA_TRY(cudaMemcpy2DAsync(h->vectors.p + first * ld, (size_t)ld * 4, vectors, (size_t)dim * 4, (size_t)dim * 4,
                                   count, cudaMemcpyHostToDevice, st));
  }
  ISL_TRY(launch_row_sqnorms(h->vectors.p + first * ld, count, dim, ld, h->sqnorms.p + first, h->sms, st));
  ISL_CUDA_TRY(cudaMemsetAsync(h->deg0.p + first, 0, count * 4, st));
  if (n_upper > h->n_upper) ISL_CUDA_TRY(cudaMemsetAsync(h->degU.p + h->n_upper, 0, (n_upper - h->n_upper) * 4, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->node_levels.p + first, h->h_levels.data() + first, count * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->row_map.p + first, h->h_row_map.data() + first, count * 4, cudaMemcpyHostToDevice, st));
  h->n_upper = n_upper;
  h->n = total;  // ids are assigned up front (hnsw.rs:227-228); rows of later rounds are unreachable until applied
  const uint32_t upper_base = (uint32_t)total;

  // rounds: per round and layer, the ascending list of round nodes that own the layer
  struct Round {
    uint64_t s, e;
    std::vector<uint32_t> layer_off;  // offsets into `lists` per layer 0..=lmax, plus the end
  };
  std::vector<Round> rounds;
  std::vector<uint32_t> lists;
  uint64_t max_round = 0, max_edges = 0;
  {
    uint64_t s = first;
    if (s == 0) s = 1;  // the first node only becomes the entry point (hnsw.rs:240-245)
    while (s < total) {
      const uint64_t round = std::min<uint64_t>(batch, std::max<uint64_t>(1, s / 2));
      Round r;
      r.s = s;
      r.e = std::min<uint64_t>(total, s + round);
      uint32_t lmax = 0;
      for (uint64_t i = r.s; i < r.e; ++i) lmax = std::max(lmax, h->h_levels[i]);
      uint64_t edges = 0;
      for (uint32_t L = 0; L <= lmax; ++L) {
        r.layer_off.push_back((uint32_t)lists.size());
        for (uint64_t i = r.s; i < r.e; ++i)
          if (h->h_levels[i] >= L) lists.push_back((uint32_t)i);
        edges += (uint64_t)(lists.size() - r.layer_off.back()) * (L == 0 ? m0 : m);
      }
      r.layer_off.push_back((uint32_t)lists.size());
      max_round = std::max<uint64_t>(max_round, r.e - r.s);
      max_edges = std::max(max_edges, edges);
      s = r.e;
      rounds.push_back(std::move(r));
    }
  }
  if (max_edges >= (1ull << 31)) return fail(ISL_INVALID_ARGUMENT, "hnsw: batch too large");

  if (first == 0) {
    h->entry = 0;
    h->max_level = h->h_levels[0];
  }
  if (rounds.empty()) {
    ISL_CUDA_TRY(cudaStreamSynchronize(st));
    return ISL_OK;
  }

  DevBuf<uint32_t> d_lists, cand_ids, cand_cnt, heads, n_heads;
  DevBuf<float> cand_dist, edge_vals, edge_vals2;
  DevBuf<uint64_t> edge_keys, edge_keys2;
  DevBuf<uint8_t> cub_tmp;
  ISL_CUDA_TRY(d_lists.alloc(lists.size()));
  ISL_CUDA_TRY(cudaMemcpyAsync(d_lists.p, lists.data(), lists.size() * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cand_ids.alloc(max_round * efc));
  ISL_CUDA_TRY(cand_dist.alloc(max_round * efc));
  ISL_CUDA_TRY(cand_cnt.alloc(max_round));
  ISL_CUDA_TRY(edge_keys.alloc(max_edges));
  ISL_CUDA_TRY(edge_keys2.alloc(max_edges));
  ISL_CUDA_TRY(edge_vals.alloc(max_edges));
  ISL_CUDA_TRY(edge_vals2.alloc(max_edges));
  ISL_CUDA_TRY(heads.alloc(max_edges));
  ISL_CUDA_TRY(n_heads.alloc(1));
  size_t cub_bytes = 0;
  ISL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, edge_keys.p, edge_keys2.p, edge_vals.p, edge_vals2.p,
                                               (int)max_edges, 0, 64, st));
  ISL_CUDA_TRY(cub_tmp.alloc(cub_bytes + 16));

  SearchPlan plan;
  const uint32_t u_cap = std::max<uint32_t>(32, round_up(std::max(m0, m), 32));
  ISL_TRY(plan_search(h->cfg.metric, ld, efc, u_cap, h->sms, &plan));
  const uint32_t vis_words = round_up((uint32_t)((total + 31) / 32), 4);
  const uint32_t slots = (uint32_t)std::min<uint64_t>(plan.grid, max_round);
  ISL_TRY(ensure(h->visited, (size_t)slots * vis_words));
  if (!plan.r_in_smem) ISL_TRY(ensure(h->r_global, (size_t)slots * efc));
  ISL_TRY(ensure(h->ties_global, (size_t)slots * efc));
  ISL_CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 4 * sizeof(unsigned int), st));

  const uint32_t sel_warps = 4;
  for (const Round& r : rounds) {
    const uint32_t cnt = (uint32_t)(r.e - r.s);
    // descent of every round node from the snapshot's top layer to its own level + 1
    GreedyArgs g{};
    g.queries = h->vectors.p + r.s * ld;
    g.q_ld = ld;
    g.nq = cnt;
    g.entry = (uint32_t)h->entry;
    g.top_layer = (uint32_t)h->max_level;
    g.stop_levels = h->node_levels.p + r.s;
    g.out_cur = h->cur_by_node.p + r.s;
    ISL_TRY(launch_greedy(h, g, st));

    uint64_t n_edges = 0;
    const uint32_t lmax = (uint32_t)r.layer_off.size() - 2;
    for (uint32_t L = lmax + 1; L-- > 0;) {
      const uint32_t items = r.layer_off[L + 1] - r.layer_off[L];
      const uint32_t* item_nodes = d_lists.p + r.layer_off[L];
      const uint32_t cap_conn = L == 0 ? m0 : m;
      ISL_CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned int), st));
      SearchArgs a{};
      fill_layer_args(h, L, &a);
      a.queries = nullptr;
      a.query_ids = item_nodes;
      a.entries = h->cur_by_node.p;
      a.q_ld = ld;
      a.nq = items;
      a.entry = 0;
      a.k = efc;
      a.ef = efc;
      a.visited = h->visited.p;
      a.vis_words = vis_words;
      a.r_global = h->r_global.p;
      a.ties_global = h->ties_global.p;
      a.u_cap = u_cap;
      a.out_ids = nullptr;
      a.out_ids32 = cand_ids.p;
      a.out_dist = cand_dist.p;
      a.out_count = cand_cnt.p;
      a.stats = nullptr;
      a.work_counter = h->counters.p;
      a.error_flag = h->counters.p + 1;
      ISL_TRY(launch_search(plan, a, st));
      hnsw_select_kernel<<<(items + sel_warps - 1) / sel_warps, sel_warps * 32, 0, st>>>(
          cand_ids.p, cand_dist.p, cand_cnt.p, efc, items, item_nodes, L, cap_conn, L == 0 ? h->adj0.p : h->adjU.p,
          L == 0 ? h->dist0.p : h->distU.p, L == 0 ? h->deg0.p : h->degU.p, h->row_map.p, h->node_levels.p, upper_base,
          edge_keys.p + n_edges, edge_vals.p + n_edges, h->cur_by_node.p);
      count_launch();
      n_edges += (uint64_t)items * cap_conn;
    }
    size_t tmp_bytes = cub_bytes;
    ISL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, edge_keys.p, edge_keys2.p, edge_vals.p,
                                                 edge_vals2.p, (int)n_edges, 0, 64, st));
    count_launch(3);
    ISL_CUDA_TRY(cudaMemsetAsync(n_heads.p, 0, 4, st));
    hnsw_segment_heads_kernel<<<(uint32_t)std::min<uint64_t>((n_edges + 255) / 256, 1184), 256, 0, st>>>(
        edge_keys2.p, (uint32_t)n_edges, heads.p, n_heads.p);
    count_launch();
    const uint32_t rev_blocks = (uint32_t)std::min<uint64_t>((n_edges + sel_warps - 1) / sel_warps, (uint64_t)h->sms * 8);
    hnsw_apply_reverse_kernel<<<rev_blocks, sel_warps * 32, (size_t)sel_warps * 4 * std::max(m0, m) * 4, st>>>(
        edge_keys2.p, edge_vals2.p, (uint32_t)n_edges, heads.p, n_heads.p, upper_base, m0, m, h->adj0.p, h->dist0.p,
        h->deg0.p, h->adjU.p, h->distU.p, h->degU.p);
    count_launch();
    ISL_CUDA_TRY(cudaGetLastError());
    for (uint64_t id = r.s; id < r.e; ++id) {  // hnsw.rs:322-325
      if (h->h_levels[id] > h->max_level) {
        h->max_level = h->h_levels[id];
        h->entry = (int64_t)id;
      }
    }
  }
  unsigned int hflags[4] = {0, 0, 0, 0};
  ISL_CUDA_TRY(cudaMemcpyAsync(hflags, h->counters.p, sizeof(hflags), cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  if (hflags[1]) return fail(ISL_CUDA_ERROR, "hnsw insert: internal invariant violated (tie list overflow)");
  return ISL_OK;
}

// queries on the device as [nq][q_ld]; outputs on the device.
isl_status hnsw_search_device(const isl_hnsw* h, const float* d_queries, uint32_t q_ld, uint64_t nq, uint32_t k,
                              uint32_t ef, uint64_t* d_ids, float* d_dist, uint32_t* d_count) {
  cudaStream_t st = h->stream;
  SearchPlan plan;
  const uint32_t u_cap = std::max<uint32_t>(32, round_up((uint32_t)std::max(h->cfg.m0, h->cfg.m), 32));
  ISL_TRY(plan_search(h->cfg.metric, h->ld, ef, u_cap, h->sms, &plan));
  const uint32_t vis_words = round_up((uint32_t)((h->n + 31) / 32), 4);
  const uint32_t slots = (uint32_t)std::min<uint64_t>(plan.grid, nq);
  ISL_TRY(ensure(h->visited, (size_t)slots * vis_words));
  if (!plan.r_in_smem) ISL_TRY(ensure(h->r_global, (size_t)slots * ef));
  ISL_TRY(ensure(h->ties_global, (size_t)slots * ef));
  ISL_TRY(ensure(h->q_cur, nq));
  ISL_CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 4 * sizeof(unsigned int), st));
  ISL_CUDA_TRY(cudaEventRecord(h->ev0, st));
  GreedyArgs g{};  // hnsw.rs:473-497
  g.queries = d_queries;
  g.q_ld = q_ld;
  g.nq = (uint32_t)nq;
  g.entry = (uint32_t)h->entry;
  g.top_layer = (uint32_t)h->max_level;
  g.stop_levels = nullptr;
  g.out_cur = h->q_cur.p;
  ISL_TRY(launch_greedy(h, g, st));
  SearchArgs a{};  // hnsw.rs:500-503
  fill_layer_args(h, 0, &a);
  a.queries = d_queries;
  a.query_ids = nullptr;
  a.entries = h->q_cur.p;
  a.q_ld = q_ld;
  a.nq = (uint32_t)nq;
  a.k = k;
  a.ef = ef;
  a.visited = h->visited.p;
  a.vis_words = vis_words;
  a.r_global = h->r_global.p;
  a.ties_global = h->ties_global.p;
  a.u_cap = u_cap;
  a.out_ids = d_ids;
  a.out_dist = d_dist;
  a.out_count = d_count;
  a.work_counter = h->counters.p;
  a.error_flag = h->counters.p + 1;
  ISL_TRY(launch_search(plan, a, st));
  ISL_CUDA_TRY(cudaEventRecord(h->ev1, st));
  return ISL_OK;
}

isl_status hnsw_search_finish(const isl_hnsw* h) {
  unsigned int hf[4] = {0, 0, 0, 0};
  ISL_CUDA_TRY(cudaMemcpyAsync(hf, h->counters.p, sizeof(hf), cudaMemcpyDeviceToHost, h->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.0f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_kernel_ms = ms;
  if (hf[1])
    return fail(ISL_INVALID_ARGUMENT, "search: more than 64 unexpanded candidates tie exactly with the worst result distance");
  return ISL_OK;
}

isl_status hnsw_search_checks(const isl_hnsw* h, const void* queries, uint64_t nq, uint32_t query_dim, uint32_t k,
                              uint32_t* ef, bool* trivial) {
  *trivial = false;
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  if (nq > 0 && !queries) return fail(ISL_INVALID_ARGUMENT, "queries is null");
  if (nq > 0xffffffffull) return fail(ISL_INVALID_ARGUMENT, "too many queries in one batch");
  if (h->n == 0 || nq == 0 || k == 0) {  // hnsw.rs:459-461
    *trivial = true;
    return ISL_OK;
  }
  if (query_dim != h->dim)  // hnsw.rs:463-470
    return fail(ISL_DIM_MISMATCH, "dimension mismatch: expected " + std::to_string(h->dim) + ", got " + std::to_string(query_dim));
  if (h->entry < 0) return fail(ISL_INDEX_NOT_BUILT, "index not built");  // hnsw.rs:472
  *ef = std::max(*ef, k);                                                 // hnsw.rs:500
  if (*ef > (1u << 24)) return fail(ISL_INVALID_ARGUMENT, "ef too large");
  return ISL_OK;
}

}  // namespace
}  // namespace isl

using namespace isl;

extern "C" {

isl_status isl_hnsw_new(const isl_hnsw_config* cfg, isl_hnsw** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  ISL_TRY(isl_hnsw_config_validate(cfg));  // HnswGraph::new (hnsw.rs:167-178)
  if (cfg->metric < 0 || cfg->metric > 3) return fail(ISL_INVALID_CONFIG, "unknown metric");
  std::unique_ptr<isl_hnsw> h(new isl_hnsw());
  h->cfg = *cfg;
  ISL_TRY(current_device(&h->device, &h->sms));
  ISL_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  ISL_CUDA_TRY(cudaEventCreate(&h->ev0));
  ISL_CUDA_TRY(cudaEventCreate(&h->ev1));
  ISL_TRY(ensure(h->counters, 4));
  *out = h.release();
  return ISL_OK;
} ISL_ABI_GUARD

void isl_hnsw_free(isl_hnsw* h) {
  if (!h) return;
  DeviceGuard g(h->device);
  delete h;
}

uint64_t isl_hnsw_len(const isl_hnsw* h) { return h ? h->n : 0; }
uint32_t isl_hnsw_dimension(const isl_hnsw* h) { return h ? h->dim : 0; }
int64_t isl_hnsw_entry_point(const isl_hnsw* h) { return h ? h->entry : ISL_NO_ENTRY; }
uint64_t isl_hnsw_max_level(const isl_hnsw* h) { return h ? h->max_level : 0; }

isl_status isl_hnsw_insert_batch(isl_hnsw* h, const float* vectors, uint64_t count, uint32_t dim,
                                 const uint64_t* levels_or_null, uint64_t seed, uint32_t batch, uint64_t* out_first_id) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  return hnsw_insert_impl(h, vectors, false, count, dim, levels_or_null, seed, batch, out_first_id);
} ISL_ABI_GUARD

isl_status isl_hnsw_insert_batch_dev(isl_hnsw* h, const float* d_vectors, uint64_t count, uint32_t dim,
                                     const uint64_t* levels_or_null, uint64_t seed, uint32_t batch,
                                     uint64_t* out_first_id) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  DeviceGuard g(h->device);
  cudaError_t e = cudaDeviceSynchronize();  // d_vectors may have been written on another stream
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceSynchronize");
  std::lock_guard<std::mutex> lock(h->mu);
  return hnsw_insert_impl(h, d_vectors, true, count, dim, levels_or_null, seed, batch, out_first_id);
} ISL_ABI_GUARD

isl_status isl_hnsw_node_level(const isl_hnsw* h, uint64_t node_id, uint64_t* out_level) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  if (node_id >= h->n) return fail(ISL_NODE_NOT_FOUND, "node " + std::to_string(node_id) + " not found");
  if (out_level) *out_level = h->h_levels[node_id];
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_hnsw_get_neighbors(const isl_hnsw* h, uint64_t node_id, uint64_t layer, uint64_t* out, uint64_t cap,
                                  uint64_t* out_count) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  if (node_id >= h->n) return fail(ISL_NODE_NOT_FOUND, "node " + std::to_string(node_id) + " not found");
  if (layer > h->h_levels[node_id])  // neighbors_at(layer) == None (hnsw.rs:108-110)
    return fail(ISL_INVALID_ARGUMENT, "node " + std::to_string(node_id) + " has no layer " + std::to_string(layer));
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  const uint32_t cc = layer == 0 ? (uint32_t)h->cfg.m0 : (uint32_t)h->cfg.m;
  const uint64_t row = layer == 0 ? node_id : h->h_row_map[node_id] + layer - 1;
  const uint32_t* adj = (layer == 0 ? h->adj0.p : h->adjU.p) + row * cc;
  const uint32_t* deg = (layer == 0 ? h->deg0.p : h->degU.p) + row;
  uint32_t d = 0;
  std::vector<uint32_t> ids(cc);
  ISL_CUDA_TRY(cudaMemcpyAsync(&d, deg, 4, cudaMemcpyDeviceToHost, h->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(ids.data(), adj, (size_t)cc * 4, cudaMemcpyDeviceToHost, h->stream));
  ISL_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (out_count) *out_count = d;
  for (uint64_t i = 0; i < d && i < cap && out; ++i) out[i] = ids[i];
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_hnsw_export_layer(const isl_hnsw* h, uint64_t layer, int64_t* out_degrees, uint64_t* out_neighbors) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  if (h->n == 0) return ISL_OK;
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  const uint32_t cc = layer == 0 ? (uint32_t)h->cfg.m0 : (uint32_t)h->cfg.m;
  const uint64_t rows = layer == 0 ? h->n : h->n_upper;
  std::vector<uint32_t> deg(rows), adj(rows * cc);
  if (rows) {
    ISL_CUDA_TRY(cudaMemcpyAsync(deg.data(), layer == 0 ? h->deg0.p : h->degU.p, rows * 4, cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA_TRY(cudaMemcpyAsync(adj.data(), layer == 0 ? h->adj0.p : h->adjU.p, rows * cc * 4, cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  for (uint64_t i = 0; i < h->n; ++i) {
    const bool has = h->h_levels[i] >= layer;
    const uint64_t row = layer == 0 ? i : h->h_row_map[i] + layer - 1;
    const uint32_t d = has ? deg[row] : 0;
    if (out_degrees) out_degrees[i] = has ? (int64_t)d : -1;
    if (out_neighbors)
      for (uint32_t j = 0; j < cc; ++j) out_neighbors[i * cc + j] = j < d ? adj[row * cc + j] : ISL_INVALID_ID;
  }
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_hnsw_search(const isl_hnsw* h, const float* queries, uint64_t nq, uint32_t query_dim, uint32_t k,
                           uint32_t ef, uint64_t* out_ids, float* out_dist, uint32_t* out_count) try {
  bool trivial;
  ISL_TRY(hnsw_search_checks(h, queries, nq, query_dim, k, &ef, &trivial));
  if (trivial) {
    fill_empty(nq, k, out_ids, out_dist, out_count, nullptr);
    return ISL_OK;
  }
  if (!out_ids || !out_dist) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  ISL_TRY(ensure(h->q_stage, nq * h->ld));
  ISL_TRY(ensure(h->out_ids, nq * k));
  ISL_TRY(ensure(h->out_dist, nq * k));
  ISL_TRY(ensure(h->out_count, nq));
  if (h->ld != h->dim) ISL_CUDA_TRY(cudaMemsetAsync(h->q_stage.p, 0, nq * h->ld * 4, h->stream));
  ISL_CUDA_TRY(cudaMemcpy2DAsync(h->q_stage.p, (size_t)h->ld * 4, queries, (size_t)h->dim * 4, (size_t)h->dim * 4, nq,
                                 cudaMemcpyHostToDevice, h->stream));
  ISL_TRY(hnsw_search_device(h, h->q_stage.p, h->ld, nq, k, ef, h->out_ids.p, h->out_dist.p, h->out_count.p));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, h->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, h->stream));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, h->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out_count)
    ISL_CUDA_TRY(cudaMemcpyAsync(out_count, h->out_count.p, nq * 4, cudaMemcpyDeviceToHost, h->stream));
  return hnsw_search_finish(h);
} ISL_ABI_GUARD

isl_status isl_hnsw_search_dev(const isl_hnsw* h, const float* d_queries, uint64_t nq, uint32_t query_dim, uint32_t k,
                               uint32_t ef, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_count) try {
  bool trivial;
  ISL_TRY(hnsw_search_checks(h, d_queries, nq, query_dim, k, &ef, &trivial));
  if (!d_out_ids || !d_out_dist) return fail(ISL_INVALID_ARGUMENT, "output pointer is null");
  if (trivial) {
    if (nq && k) {
      ISL_CUDA_TRY(cudaMemset(d_out_ids, 0xff, nq * k * 8));
      std::vector<float> inf(nq * k, std::numeric_limits<float>::infinity());
      ISL_CUDA_TRY(cudaMemcpy(d_out_dist, inf.data(), nq * k * 4, cudaMemcpyHostToDevice));
    }
    if (d_out_count && nq) ISL_CUDA_TRY(cudaMemset(d_out_count, 0, nq * 4));
    return ISL_OK;
  }
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  const float* q = d_queries;
  uint32_t q_ld = query_dim;
  if (query_dim % 4 != 0 || (reinterpret_cast<uintptr_t>(d_queries) & 15)) {
    ISL_TRY(ensure(h->q_stage, nq * h->ld));
    ISL_TRY(launch_pad_rows(d_queries, query_dim, h->q_stage.p, h->ld, nq, h->stream));
    q = h->q_stage.p;
    q_ld = h->ld;
  }
  ISL_TRY(hnsw_search_device(h, q, q_ld, nq, k, ef, d_out_ids, d_out_dist, d_out_count));
  return hnsw_search_finish(h);
} ISL_ABI_GUARD

// HnswGraph { config, nodes: HashMap<u64, HnswNode{id, vector, connections, level}>, entry_point,
// max_level, dimension, next_id } (hnsw.rs:151-164, :90-99) in the bincode layout of hnsw.rs:507-514.
// The map is written in ascending id order (any order is a valid encoding of a HashMap).
isl_status isl_hnsw_to_bytes(const isl_hnsw* h, uint8_t* out, uint64_t cap, uint64_t* out_len) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  const uint64_t n = h->n;
  const uint32_t m0 = (uint32_t)h->cfg.m0, m = (uint32_t)h->cfg.m;
  std::vector<float> vec(n * h->dim);
  std::vector<uint32_t> deg0(n), adj0(n * m0), degU(h->n_upper), adjU(h->n_upper * m);
  if (n) {
    ISL_CUDA_TRY(cudaMemcpy2DAsync(vec.data(), (size_t)h->dim * 4, h->vectors.p, (size_t)h->ld * 4, (size_t)h->dim * 4, n,
                                   cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA_TRY(cudaMemcpyAsync(deg0.data(), h->deg0.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA_TRY(cudaMemcpyAsync(adj0.data(), h->adj0.p, n * m0 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (h->n_upper) {
      ISL_CUDA_TRY(cudaMemcpyAsync(degU.data(), h->degU.p, h->n_upper * 4, cudaMemcpyDeviceToHost, h->stream));
      ISL_CUDA_TRY(cudaMemcpyAsync(adjU.data(), h->adjU.p, h->n_upper * m * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    ISL_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  ByteWriter w;
  w.u64(h->cfg.m);  // HnswConfig (hnsw.rs:15-28)
  w.u64(h->cfg.m0);
  w.u64(h->cfg.ef_construction);
  w.f64(h->cfg.ml);
  w.u32((uint32_t)h->cfg.metric);
  w.u64(h->cfg.max_layers);
  w.u64(n);
  for (uint64_t i = 0; i < n; ++i) {
    w.u64(i);  // key
    w.u64(i);  // HnswNode::id
    w.vec_f32(vec.data() + i * h->dim, h->dim);
    const uint32_t lv = h->h_levels[i];
    w.u64((uint64_t)lv + 1);  // connections: one list per layer 0..=level (hnsw.rs:104-106)
    for (uint32_t L = 0; L <= lv; ++L) {
      const uint64_t row = L == 0 ? i : h->h_row_map[i] + L - 1;
      const uint32_t d = L == 0 ? deg0[row] : degU[row];
      const uint32_t* a = L == 0 ? adj0.data() + row * m0 : adjU.data() + row * m;
      w.u64(d);
      for (uint32_t t = 0; t < d; ++t) w.u64(a[t]);
    }
    w.u64(lv);
  }
  w.opt_u64(h->entry >= 0, (uint64_t)h->entry);
  w.u64(h->max_level);
  w.opt_u64(h->dim != 0, h->dim);
  w.u64(n);  // next_id
  if (out_len) *out_len = w.buf.size();
  if (!out) return ISL_OK;
  if (cap < w.buf.size()) return fail(ISL_INVALID_ARGUMENT, "output buffer too small: need " + std::to_string(w.buf.size()) + " bytes");
  std::memcpy(out, w.buf.data(), w.buf.size());
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_hnsw_from_bytes(const uint8_t* bytes, uint64_t len, isl_hnsw** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (!bytes) return fail(ISL_INVALID_ARGUMENT, "bytes is null");
  ByteReader r(bytes, len);
  isl_hnsw_config c{};
  c.m = r.u64();
  c.m0 = r.u64();
  c.ef_construction = r.u64();
  c.ml = r.f64();
  c.metric = (int32_t)r.u32();
  c.max_layers = r.u64();
  const uint64_t n = r.seq_len(40);
  struct Node {
    std::vector<float> v;
    std::vector<std::vector<uint64_t>> conn;
    uint64_t level = 0;
    bool seen = false;
  };
  std::vector<Node> nodes(r.ok ? n : 0);
  for (uint64_t i = 0; i < n && r.ok; ++i) {
    const uint64_t key = r.u64(), id = r.u64();
    if (!r.ok || key != id || id >= n || nodes[id].seen) {  // ids are 0..n-1 (hnsw.rs:227-228)
      r.ok = false;
      break;
    }
    Node& nd = nodes[id];
    nd.seen = true;
    r.vec_f32(&nd.v);
    const uint64_t layers = r.seq_len(8);
    nd.conn.resize(r.ok ? layers : 0);
    for (uint64_t L = 0; L < layers && r.ok; ++L) r.vec_u64(&nd.conn[L]);
    nd.level = r.u64();
    if (nd.conn.size() != nd.level + 1 || nd.level > 255) r.ok = false;
  }
  uint64_t entry = 0, sdim = 0;
  const bool has_entry = r.opt_u64(&entry);
  const uint64_t max_level = r.u64();
  const bool has_dim = r.opt_u64(&sdim);
  const uint64_t next_id = r.u64();
  if (!r.done() || c.metric < 0 || c.metric > 3 || next_id != n || (n > 0 && (!has_entry || !has_dim || entry >= n)))
    return fail(ISL_SERIALIZATION, "deserialization failed: malformed HnswGraph bytes");
  isl_hnsw* h = nullptr;
  ISL_TRY(isl_hnsw_new(&c, &h));
  std::unique_ptr<isl_hnsw, void (*)(isl_hnsw*)> guard(h, isl_hnsw_free);
  if (n == 0) {
    *out = guard.release();
    return ISL_OK;
  }
  if (sdim == 0 || sdim > 0xffffffffull) return fail(ISL_SERIALIZATION, "deserialization failed: bad dimension");
  const uint32_t dim = (uint32_t)sdim, m0 = (uint32_t)c.m0, m = (uint32_t)c.m;
  DeviceGuard g(h->device);
  h->dim = dim;
  h->ld = std::max<uint32_t>(4, round_up(dim, 4));
  h->n = n;
  h->cap = n;
  h->h_levels.resize(n);
  h->h_row_map.resize(n);
  uint64_t n_upper = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (nodes[i].v.size() != dim) return fail(ISL_SERIALIZATION, "deserialization failed: vector length differs from dimension");
    h->h_levels[i] = (uint32_t)nodes[i].level;
    h->h_row_map[i] = (uint32_t)n_upper;
    n_upper += nodes[i].level;
  }
  h->n_upper = h->cap_upper = n_upper;
  std::vector<float> vec((size_t)n * h->ld, 0.0f);
  std::vector<uint32_t> deg0(n), adj0((size_t)n * m0, 0), degU(n_upper), adjU((size_t)n_upper * m, 0);
  for (uint64_t i = 0; i < n; ++i) {
    std::memcpy(vec.data() + i * h->ld, nodes[i].v.data(), (size_t)dim * 4);
    for (uint64_t L = 0; L <= nodes[i].level; ++L) {
      const auto& cl = nodes[i].conn[L];
      const uint32_t cc = L == 0 ? m0 : m;
      if (cl.size() > cc) return fail(ISL_SERIALIZATION, "deserialization failed: connection list longer than its capacity");
      const uint64_t row = L == 0 ? i : h->h_row_map[i] + L - 1;
      uint32_t* a = L == 0 ? adj0.data() + row * m0 : adjU.data() + row * m;
      for (size_t t = 0; t < cl.size(); ++t) {
        if (cl[t] >= n) return fail(ISL_NODE_NOT_FOUND, "node " + std::to_string(cl[t]) + " not found");
        a[t] = (uint32_t)cl[t];
      }
      (L == 0 ? deg0[row] : degU[row]) = (uint32_t)cl.size();
    }
  }
  cudaStream_t st = h->stream;
  ISL_CUDA_TRY(h->vectors.alloc(n * h->ld));
  ISL_CUDA_TRY(h->sqnorms.alloc(n));
  ISL_CUDA_TRY(h->adj0.alloc(n * m0));
  ISL_CUDA_TRY(h->dist0.alloc(n * m0));
  ISL_CUDA_TRY(h->deg0.alloc(n));
  ISL_CUDA_TRY(h->node_levels.alloc(n));
  ISL_CUDA_TRY(h->row_map.alloc(n));
  ISL_CUDA_TRY(h->cur_by_node.alloc(n));
  ISL_CUDA_TRY(h->adjU.alloc(std::max<uint64_t>(n_upper * m, 1)));
  ISL_CUDA_TRY(h->distU.alloc(std::max<uint64_t>(n_upper * m, 1)));
  ISL_CUDA_TRY(h->degU.alloc(std::max<uint64_t>(n_upper, 1)));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->vectors.p, vec.data(), vec.size() * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->adj0.p, adj0.data(), adj0.size() * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->deg0.p, deg0.data(), n * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->node_levels.p, h->h_levels.data(), n * 4, cudaMemcpyHostToDevice, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(h->row_map.p, h->h_row_map.data(), n * 4, cudaMemcpyHostToDevice, st));
  if (n_upper) {
    ISL_CUDA_TRY(cudaMemcpyAsync(h->adjU.p, adjU.data(), adjU.size() * 4, cudaMemcpyHostToDevice, st));
    ISL_CUDA_TRY(cudaMemcpyAsync(h->degU.p, degU.data(), n_upper * 4, cudaMemcpyHostToDevice, st));
  }
  ISL_TRY(launch_row_sqnorms(h->vectors.p, n, dim, h->ld, h->sqnorms.p, h->sms, st));
  // cached edge distances (prune_connections needs them on the next insert): recomputed row by row
  ISL_TRY(launch_edge_distances(h, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  h->entry = (int64_t)entry;
  h->max_level = max_level;
  *out = guard.release();
  return ISL_OK;
} ISL_ABI_GUARD

// HnswNode::vector of get_node(node_id) (hnsw.rs:93-95, :201-203).
isl_status isl_hnsw_get_vector(const isl_hnsw* h, uint64_t node_id, float* out) try {
  if (!h || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (node_id >= h->n) return fail(ISL_NODE_NOT_FOUND, "node " + std::to_string(node_id) + " not found");
  DeviceGuard g(h->device);
  ISL_CUDA_TRY(cudaMemcpy(out, h->vectors.p + node_id * h->ld, (size_t)h->dim * 4, cudaMemcpyDeviceToHost));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_hnsw_get_config(const isl_hnsw* h, isl_hnsw_config* out) try {
  if (!h || !out) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  *out = h->cfg;
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_hnsw_last_search_timing(const isl_hnsw* h, float* kernel_ms) try {
  if (!h) return fail(ISL_INVALID_ARGUMENT, "graph is null");
  if (kernel_ms) *kernel_ms = h->last_kernel_ms;
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
