// api_shard.cu — sharded search behind the C ABI: every rank searches its own shard (node range or island),
// the per-shard top-k lists are exchanged as packed 16-byte (dist, global id) records with ONE all-gather,
// and a per-query merge by (dist, id) produces the same top-k on every rank.
//
// Reference semantics: IndexerService::search loops over its per-repository graphs and merges by score
// (src/indexer/service.rs:777-801); MultiIndexSearcher::search does the same over named indexes and sorts
// ascending by distance (src/core/search.rs:211-237).  Here the loop is one rank per shard, and the merge rule
// is the fixed (distance, id) order used everywhere else in this library.
//
// Everything of one call runs on ONE stream (the call's leased stream): search kernel -> all-gather -> merge
// kernel, no host synchronisation in between.  Two exchange engines:
//   * NCCL (ncclAllGather over NVLink / NVSwitch), resolved at run time from libnccl.so.2 — the library has no
//     link-time dependency on it, and a process that already loaded NCCL (torch) shares that copy;
//   * peer stores (isl_shard_enable_peer_exchange): the gather buffers of all ranks are mapped into each other
//     with CUDA IPC and the SEARCH KERNEL ITSELF stores each finished query's records into every rank's buffer
//     (16-byte stores over NVLink, search_core.cuh epilogue), so the transfer overlaps the search query by query;
//     a flag handshake over the same mapped memory replaces the collective.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>

#include "api_common.h"

namespace isl {
namespace {

// ---- NCCL, resolved at run time ------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nme : names) {
      api.handle = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown error");
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(api.handle, s);
      if (!p && api.error.empty()) api.error = std::string("libnccl lacks symbol ") + s;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  });
  return &api;
}

isl_status nccl_fail(ncclResult_t r, const char* what) {
  NcclApi* api = nccl_api();
  return fail(ISL_CUDA_ERROR, std::string("NCCL error: ") + (api->GetErrorString ? api->GetErrorString(r) : "?") + " in " + what);
}
#define ISL_NCCL_TRY(expr)                              \
  do {                                                  \
    ncclResult_t _r = (expr);                           \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr); \
  } while (0)

// ---- K8 on packed records: per query, the k best of parts * k records by (dist, id) -----------------------
// records [parts][nq][k] (consecutive parts part_stride records apart) = {dist bits, 0, id lo, id hi}; all-ones
// ids are padding.  One warp per query.
__global__ void __launch_bounds__(128)
merge_packed_kernel(const uint4* __restrict__ rec, uint32_t parts, uint64_t part_stride, uint64_t nq, uint32_t k,
                    uint64_t* __restrict__ out_ids, float* __restrict__ out_dist, uint32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t total = parts * k;
  uint4* s = reinterpret_cast<uint4*>(smem_raw) + (size_t)warp * total;
  const uint64_t qi = blockIdx.x * (uint64_t)(blockDim.x >> 5) + warp;
  if (qi >= nq) return;
  for (uint32_t i = lane; i < total; i += 32) {
    const uint32_t p = i / k, j = i % k;
    s[i] = __ldcg(rec + (uint64_t)p * part_stride + qi * k + j);
  }
  __syncwarp();
  uint32_t produced = 0;
  for (uint32_t r = 0; r < k; ++r) {
    float bd = 0.0f;
    uint64_t bid = ISL_INVALID_ID;
    uint32_t bpos = 0xffffffffu;
    for (uint32_t i = lane; i < total; i += 32) {
      const uint4 e = s[i];
      const uint64_t id = ((uint64_t)e.w << 32) | e.z;
      if (id == ISL_INVALID_ID) continue;
      const float dd = __uint_as_float(e.x);
      if (bpos == 0xffffffffu || key_lt64(dd, id, bd, bid) || (!key_lt64(bd, bid, dd, id) && i < bpos)) {
        bd = dd;
        bid = id;
        bpos = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, off);
      const uint64_t oid = __shfl_xor_sync(0xffffffffu, bid, off);
      const uint32_t opos = __shfl_xor_sync(0xffffffffu, bpos, off);
      if (opos != 0xffffffffu &&
          (bpos == 0xffffffffu || key_lt64(od, oid, bd, bid) || (!key_lt64(bd, bid, od, oid) && opos < bpos))) {
        bd = od;
        bid = oid;
        bpos = opos;
      }
    }
    if (bpos == 0xffffffffu) break;
    if (lane == 0) {
      out_ids[qi * k + r] = bid;
      out_dist[qi * k + r] = bd;
      s[bpos].z = 0xffffffffu;
      s[bpos].w = 0xffffffffu;
    }
    produced++;
    __syncwarp();
  }
  for (uint32_t r = produced + lane; r < k; r += 32) {
    out_ids[qi * k + r] = ISL_INVALID_ID;
    out_dist[qi * k + r] = __int_as_float(0x7f800000);
  }
  if (lane == 0 && out_count) out_count[qi] = produced;
}

isl_status launch_merge_packed(const uint4* d_rec, uint32_t parts, uint64_t part_stride, uint64_t nq, uint32_t k,
                               uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_count, cudaStream_t st) {
  if (nq == 0 || k == 0) return ISL_OK;
  const uint32_t warps = 4;
  const size_t smem = (size_t)warps * parts * k * 16;
  if (smem > 200 * 1024) return fail(ISL_INVALID_ARGUMENT, "merge: parts*k too large for shared memory");
  ISL_CUDA_TRY(cudaFuncSetAttribute(merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const uint32_t grid = (uint32_t)((nq + warps - 1) / warps);
  merge_packed_kernel<<<grid, warps * 32, smem, st>>>(d_rec, parts, part_stride, nq, k, d_out_ids, d_out_dist, d_out_count);
  count_launch();
  ISL_CUDA_TRY(cudaGetLastError());
  return ISL_OK;
}

// ---- peer-store exchange --------------------------------------------------------------------------------
// The search kernel itself stores every finished query's records into slot `rank` of every rank's gather
// buffer (ShardOut::peer, 16-byte stores over NVLink), so the transfer rides under the search.  What is left of
// the collective is this handshake: publish the step number in every peer's flag word (system-scope release of
// the stores of the search kernel that ran before on this stream), then wait until every peer's flag shows the
// step.  Gather buffers are double-buffered by step parity: a rank can only start pushing step s+2 after every
// peer has signalled s+1, which a peer does after its merge of step s has left the buffer.
struct PeerTable {
  uint4* gather[kMaxPeers];         // peer r's gather buffer [2][world][cap] (own entry: the local buffer)
  unsigned int* flags[kMaxPeers];   // peer r's flag array [world]
};

__global__ void peer_signal_wait_kernel(PeerTable t, int rank, int world, unsigned int step, unsigned int* error_flag) {
  const int r = threadIdx.x;  // one thread per peer
  if (r < world) {
    __threadfence_system();
    volatile unsigned int* theirs = t.flags[r] + rank;
    *theirs = step;
    __threadfence_system();
    volatile unsigned int* mine = t.flags[rank] + r;
    const long long t0 = clock64();
    while (*mine < step) {
      // a peer that never arrives (its process died) must not hang this GPU: give up after ~10 s of SM clocks
      if (clock64() - t0 > 20000000000ll) {
        atomicExch(error_flag, 2u);
        break;
      }
    }
    __threadfence_system();
  }
}

}  // namespace
}  // namespace isl

using namespace isl;

struct isl_shard {
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  std::mutex mu;  // one exchange at a time per communicator
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // search start / search end / exchange end / merge end
  float search_ms = 0.0f, exchange_ms = 0.0f, merge_ms = 0.0f;
  // peer-store exchange
  bool peer = false;
  uint64_t peer_cap = 0;          // records per rank slot
  DevBuf<uint4> peer_gather;      // [2][world][peer_cap], mapped into every peer
  DevBuf<unsigned int> peer_flags;  // [world]
  DevBuf<uint4> nccl_stage;       // scratch for the handle all-gather
  PeerTable table{};
  std::vector<void*> opened;      // cudaIpcOpenMemHandle mappings to close
  unsigned int step = 0;
  ~isl_shard() {
    for (void* p : opened) cudaIpcCloseMemHandle(p);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    if (comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(comm);
  }
};

namespace {

// search (records) -> exchange -> merge on sc->stream.  d_queries already staged ([nq][q_ld]).
// mode: ISL_SHARD_EXACT (q = staged device queries), ISL_SHARD_ADC_RERANK / ISL_SHARD_ADC_RECOMPUTE (h_queries = the
// caller's host queries: those pipelines stage them themselves).
isl_status sharded_core(const isl_index* idx, isl_shard* sh, SearchScratch* sc, uint64_t id_base, int mode, const float* q,
                        uint32_t q_ld, const float* h_queries, uint64_t nq, uint32_t k, uint32_t ef, bool trivial,
                        uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_count) {
  NcclApi* api = nccl_api();
  cudaStream_t st = sc->stream;
  const uint64_t cnt = nq * k;
  ISL_TRY(ensure(sc->packed, cnt));
  const bool use_peer = sh->peer && cnt <= sh->peer_cap;
  if (!use_peer) ISL_TRY(ensure(sc->gathered, cnt * sh->world));
  ShardOut so;
  so.id_base = id_base;
  const uint4* gathered = nullptr;
  uint64_t part_stride = cnt;
  if (use_peer) {
    sh->step++;
    const uint64_t buf = (uint64_t)(sh->step & 1u) * sh->world * sh->peer_cap;
    for (int r = 0; r < sh->world; ++r) so.peer[r] = sh->table.gather[r] + buf + (uint64_t)sh->rank * sh->peer_cap;
    so.n_peer = (uint32_t)sh->world;
    gathered = sh->peer_gather.p + buf;
    part_stride = sh->peer_cap;
  } else {
    so.packed = sc->packed.p;
    gathered = sc->gathered.p;
  }
  ISL_CUDA_TRY(cudaEventRecord(sh->ev[0], st));
  if (trivial) {  // an empty shard contributes only padding records
    ISL_CUDA_TRY(cudaMemsetAsync(sc->counters.p, 0, 4 * sizeof(unsigned int), st));
    if (use_peer) {
      ISL_CUDA_TRY(cudaMemsetAsync(sc->packed.p, 0xff, cnt * 16, st));
      for (int r = 0; r < sh->world; ++r)
        ISL_CUDA_TRY(cudaMemcpyAsync(so.peer[r], sc->packed.p, cnt * 16, cudaMemcpyDefault, st));
    } else {
      ISL_CUDA_TRY(cudaMemsetAsync(sc->packed.p, 0xff, cnt * 16, st));
    }
  } else if (mode == ISL_SHARD_ADC_RERANK) {
    ISL_TRY(pq_search_on_scratch(2, idx, sc, h_queries, nq, k, ef, 0.0f, nullptr, nullptr, nullptr, nullptr, &so));
  } else if (mode == ISL_SHARD_ADC_RECOMPUTE) {
    ISL_TRY(adc_recompute_on_scratch(idx, sc, h_queries, nq, k, ef, nullptr, nullptr, nullptr, nullptr, &so));
  } else {
    ISL_TRY(search_device(idx, sc, q, q_ld, nq, k, ef, nullptr, nullptr, nullptr, nullptr, &so));
  }
  ISL_CUDA_TRY(cudaEventRecord(sh->ev[1], st));
  if (use_peer) {
    peer_signal_wait_kernel<<<1, 32, 0, st>>>(sh->table, sh->rank, sh->world, sh->step, sc->counters.p + 3);
    count_launch();
    ISL_CUDA_TRY(cudaGetLastError());
  } else {
    ISL_NCCL_TRY(api->AllGather(sc->packed.p, sc->gathered.p, cnt * 16, ncclChar, sh->comm, st));
  }
  ISL_CUDA_TRY(cudaEventRecord(sh->ev[2], st));
  ISL_TRY(launch_merge_packed(gathered, (uint32_t)sh->world, part_stride, nq, k, d_out_ids, d_out_dist, d_out_count, st));
  ISL_CUDA_TRY(cudaEventRecord(sh->ev[3], st));
  return ISL_OK;
}

isl_status sharded_finish(const isl_index* idx, isl_shard* sh, SearchScratch* sc, bool trivial) {
  if (trivial) {
    ISL_CUDA_TRY(cudaStreamSynchronize(sc->stream));
  } else {
    ISL_TRY(search_finish(idx, sc, 3));
  }
  unsigned int peer_err = 0;
  ISL_CUDA_TRY(cudaMemcpy(&peer_err, sc->counters.p + 3, sizeof(peer_err), cudaMemcpyDeviceToHost));
  if (peer_err) return fail(ISL_CUDA_ERROR, "sharded search: a peer rank did not reach the exchange (peer-store handshake timed out)");
  cudaEventElapsedTime(&sh->search_ms, sh->ev[0], sh->ev[1]);
  cudaEventElapsedTime(&sh->exchange_ms, sh->ev[1], sh->ev[2]);
  cudaEventElapsedTime(&sh->merge_ms, sh->ev[2], sh->ev[3]);
  return ISL_OK;
}

// Every rank must take part in the exchange, also one whose shard is empty: the argument checks that can
// differ between ranks (empty index) must not return early.
isl_status sharded_checks(const isl_index* idx, isl_shard* sh, const void* queries, uint64_t nq, uint32_t query_dim,
                          uint32_t k, uint32_t* ef, bool* trivial) {
  if (!sh) return fail(ISL_INVALID_ARGUMENT, "shard handle is null");
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (nq == 0 || k == 0) return fail(ISL_INVALID_ARGUMENT, "sharded search needs nq > 0 and k > 0 on every rank");
  if (idx->n == 0) {
    *trivial = true;
    return ISL_OK;
  }
  ISL_TRY(search_checks(idx, queries, nq, query_dim, k, ef, trivial));
  return ISL_OK;
}

}  // namespace

extern "C" {

isl_status isl_shard_unique_id(void* out, uint64_t cap) try {
  if (!out || cap < sizeof(ncclUniqueId)) return fail(ISL_INVALID_ARGUMENT, "isl_shard_unique_id needs a 128-byte buffer");
  NcclApi* api = nccl_api();
  if (!api->error.empty()) return fail(ISL_CUDA_ERROR, api->error);
  ncclUniqueId id;
  ISL_NCCL_TRY(api->GetUniqueId(&id));
  std::memcpy(out, &id, sizeof(id));
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_shard_init(int rank, int world, const void* nccl_uid, isl_shard** out) try {
  if (!out) return fail(ISL_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(ISL_INVALID_ARGUMENT, "rank / world out of range");
  if (!nccl_uid) return fail(ISL_INVALID_ARGUMENT, "nccl_uid is null");
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  NcclApi* api = nccl_api();
  if (!api->error.empty()) return fail(ISL_CUDA_ERROR, api->error);
  std::unique_ptr<isl_shard> sh(new isl_shard());
  sh->rank = rank;
  sh->world = world;
  sh->device = device;
  for (auto& e : sh->ev) ISL_CUDA_TRY(cudaEventCreate(&e));
  ncclUniqueId id;
  std::memcpy(&id, nccl_uid, sizeof(id));
  ISL_NCCL_TRY(api->CommInitRank(&sh->comm, world, id, rank));
  *out = sh.release();
  return ISL_OK;
} ISL_ABI_GUARD

void isl_shard_free(isl_shard* sh) {
  if (!sh) return;
  DeviceGuard g(sh->device);
  delete sh;
}
int isl_shard_rank(const isl_shard* sh) { return sh ? sh->rank : -1; }
int isl_shard_world(const isl_shard* sh) { return sh ? sh->world : 0; }

// Maps every rank's gather buffer and flag array into every other rank (CUDA IPC over NVLink peer access) so
// that the exchange becomes direct stores + a flag handshake.  Collective: every rank calls it with the same
// max_records (= the largest nq * k a later search will use).  All ranks must be on one node.
isl_status isl_shard_enable_peer_exchange(isl_shard* sh, uint64_t max_records) try {
  if (!sh) return fail(ISL_INVALID_ARGUMENT, "shard handle is null");
  if (sh->world > 16) return fail(ISL_INVALID_ARGUMENT, "peer exchange supports at most 16 ranks");
  if (max_records == 0) return fail(ISL_INVALID_ARGUMENT, "max_records must be > 0");
  DeviceGuard g(sh->device);
  std::lock_guard<std::mutex> lock(sh->mu);
  NcclApi* api = nccl_api();
  const int W = sh->world;
  ISL_CUDA_TRY(sh->peer_gather.alloc(2 * max_records * W));
  ISL_CUDA_TRY(sh->peer_flags.alloc(W));
  ISL_CUDA_TRY(cudaMemset(sh->peer_flags.p, 0, W * sizeof(unsigned int)));
  // exchange the IPC handles (2 x 64 bytes per rank) with the communicator we already have
  struct Handles {
    cudaIpcMemHandle_t gather, flags;
    int device;
    int pad[3];
  };
  static_assert(sizeof(Handles) % 16 == 0, "handle record must be a multiple of 16 bytes");
  Handles mine{};
  mine.device = sh->device;
  if (W > 1) {
    ISL_CUDA_TRY(cudaIpcGetMemHandle(&mine.gather, sh->peer_gather.p));
    ISL_CUDA_TRY(cudaIpcGetMemHandle(&mine.flags, sh->peer_flags.p));
  }
  const size_t hq = sizeof(Handles) / 16;
  ISL_CUDA_TRY(sh->nccl_stage.alloc(hq * (W + 1)));
  ISL_CUDA_TRY(cudaMemcpy(sh->nccl_stage.p, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  ISL_NCCL_TRY(api->AllGather(sh->nccl_stage.p, sh->nccl_stage.p + hq, sizeof(Handles), ncclChar, sh->comm, 0));
  ISL_CUDA_TRY(cudaStreamSynchronize(0));
  std::vector<Handles> all(W);
  ISL_CUDA_TRY(cudaMemcpy(all.data(), sh->nccl_stage.p + hq, sizeof(Handles) * W, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W; ++r) {
    if (r == sh->rank) {
      sh->table.gather[r] = sh->peer_gather.p;
      sh->table.flags[r] = sh->peer_flags.p;
      continue;
    }
    void *pg = nullptr, *pf = nullptr;
    ISL_CUDA_TRY(cudaIpcOpenMemHandle(&pg, all[r].gather, cudaIpcMemLazyEnablePeerAccess));
    sh->opened.push_back(pg);
    ISL_CUDA_TRY(cudaIpcOpenMemHandle(&pf, all[r].flags, cudaIpcMemLazyEnablePeerAccess));
    sh->opened.push_back(pf);
    sh->table.gather[r] = reinterpret_cast<uint4*>(pg);
    sh->table.flags[r] = reinterpret_cast<unsigned int*>(pf);
  }
  sh->peer_cap = max_records;
  sh->step = 0;
  sh->peer = true;
  return ISL_OK;
} ISL_ABI_GUARD

isl_status isl_index_search_sharded(const isl_index* idx, isl_shard* sh, uint64_t id_base, const float* queries,
                                    uint64_t nq, uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                    float* out_dist, uint32_t* out_count) try {
  bool trivial = false;
  ISL_TRY(sharded_checks(idx, sh, queries, nq, query_dim, k, &ef, &trivial));
  if (!queries || !out_ids || !out_dist) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  DeviceGuard g(sh->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  std::lock_guard<std::mutex> xl(sh->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  cudaStream_t st = sc->stream;
  const uint32_t ld = std::max<uint32_t>(4, round_up(query_dim, 4));
  ISL_TRY(ensure(sc->q_stage, nq * ld));
  ISL_TRY(ensure(sc->out_ids, nq * k));
  ISL_TRY(ensure(sc->out_dist, nq * k));
  ISL_TRY(ensure(sc->out_count, nq));
  if (ld != query_dim) ISL_CUDA_TRY(cudaMemsetAsync(sc->q_stage.p, 0, nq * ld * 4, st));
  ISL_CUDA_TRY(cudaMemcpy2DAsync(sc->q_stage.p, (size_t)ld * 4, queries, (size_t)query_dim * 4, (size_t)query_dim * 4, nq,
                                 cudaMemcpyHostToDevice, st));
  ISL_TRY(sharded_core(idx, sh, sc.get(), id_base, ISL_SHARD_EXACT, sc->q_stage.p, ld, nullptr, nq, k, ef, trivial,
                       sc->out_ids.p, sc->out_dist.p, sc->out_count.p));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, sc->out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, sc->out_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, sc->out_count.p, nq * 4, cudaMemcpyDeviceToHost, st));
  return sharded_finish(idx, sh, sc.get(), trivial);
} ISL_ABI_GUARD

isl_status isl_index_search_sharded_dev(const isl_index* idx, isl_shard* sh, uint64_t id_base, const float* d_queries,
                                        uint64_t nq, uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                                        float* d_out_dist, uint32_t* d_out_count) try {
  bool trivial = false;
  ISL_TRY(sharded_checks(idx, sh, d_queries, nq, query_dim, k, &ef, &trivial));
  if (!d_queries || !d_out_ids || !d_out_dist) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  DeviceGuard g(sh->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  std::lock_guard<std::mutex> xl(sh->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  ISL_TRY(order_after_caller(sc->stream, sc->ev_in));
  const float* q = d_queries;
  uint32_t q_ld = query_dim;
  if (!trivial) ISL_TRY(stage_device_queries(idx, sc.get(), d_queries, nq, query_dim, &q, &q_ld));
  ISL_TRY(sharded_core(idx, sh, sc.get(), id_base, ISL_SHARD_EXACT, q, q_ld, nullptr, nq, k, ef, trivial, d_out_ids, d_out_dist,
                       d_out_count));
  return sharded_finish(idx, sh, sc.get(), trivial);
} ISL_ABI_GUARD

// The sharded form of isl_index_search_adc_rerank / isl_index_search_adc_recompute: every rank runs the ADC traversal
// (+ encoder) + exact rerank on its shard, the rerank launch writes the exchange records, then exchange + merge.
isl_status isl_index_search_sharded_adc(const isl_index* idx, isl_shard* sh, uint64_t id_base, int32_t mode, const float* queries,
                                        uint64_t nq, uint32_t query_dim, uint32_t k, uint32_t ef, uint64_t* out_ids,
                                        float* out_dist, uint32_t* out_count) try {
  if (mode != ISL_SHARD_ADC_RERANK && mode != ISL_SHARD_ADC_RECOMPUTE)
    return fail(ISL_INVALID_ARGUMENT, "mode must be ISL_SHARD_ADC_RERANK or ISL_SHARD_ADC_RECOMPUTE");
  if (!sh) return fail(ISL_INVALID_ARGUMENT, "shard handle is null");
  if (!idx) return fail(ISL_INVALID_ARGUMENT, "index is null");
  if (nq == 0 || k == 0) return fail(ISL_INVALID_ARGUMENT, "sharded search needs nq > 0 and k > 0 on every rank");
  if (!queries || !out_ids || !out_dist) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  bool trivial = idx->n == 0;
  if (!trivial) ISL_TRY(search_checks(idx, queries, nq, query_dim, k, &ef, &trivial, mode == ISL_SHARD_ADC_RERANK));
  DeviceGuard g(sh->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  std::lock_guard<std::mutex> xl(sh->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  cudaStream_t st = sc->stream;
  ISL_TRY(ensure(sc->out_ids, nq * k));
  ISL_TRY(ensure(sc->out_dist, nq * k));
  ISL_TRY(ensure(sc->out_count, nq));
  DevBuf<uint64_t> m_ids;  // the pipelines use sc->out_* for their own (unmerged) results
  DevBuf<float> m_dist;
  DevBuf<uint32_t> m_cnt;
  ISL_CUDA_TRY(m_ids.alloc(nq * k));
  ISL_CUDA_TRY(m_dist.alloc(nq * k));
  ISL_CUDA_TRY(m_cnt.alloc(nq));
  ISL_TRY(sharded_core(idx, sh, sc.get(), id_base, mode, nullptr, 0, queries, nq, k, ef, trivial, m_ids.p, m_dist.p, m_cnt.p));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_ids, m_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ISL_CUDA_TRY(cudaMemcpyAsync(out_dist, m_dist.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (out_count) ISL_CUDA_TRY(cudaMemcpyAsync(out_count, m_cnt.p, nq * 4, cudaMemcpyDeviceToHost, st));
  return sharded_finish(idx, sh, sc.get(), trivial);
} ISL_ABI_GUARD

isl_status isl_shard_last_timing(const isl_shard* sh, float* search_ms, float* exchange_ms, float* merge_ms) try {
  if (!sh) return fail(ISL_INVALID_ARGUMENT, "shard handle is null");
  if (search_ms) *search_ms = sh->search_ms;
  if (exchange_ms) *exchange_ms = sh->exchange_ms;
  if (merge_ms) *merge_ms = sh->merge_ms;
  return ISL_OK;
} ISL_ABI_GUARD

// The two halves on their own (what a rank does before and after the exchange): G shards emulated on one GPU
// write their records into slot g of a [G][nq][k] buffer and merge them — the parity test of the sharded path.
isl_status isl_index_search_packed_dev(const isl_index* idx, uint64_t id_base, const float* d_queries, uint64_t nq,
                                       uint32_t query_dim, uint32_t k, uint32_t ef, isl_shard_record* d_records) try {
  bool trivial;
  ISL_TRY(search_checks(idx, d_queries, nq, query_dim, k, &ef, &trivial));
  if (!d_records && nq && k) return fail(ISL_INVALID_ARGUMENT, "d_records is null");
  if (trivial) {
    if (nq && k) ISL_CUDA_TRY(cudaMemset(d_records, 0xff, nq * k * 16));
    return ISL_OK;
  }
  DeviceGuard g(idx->device);
  std::shared_lock<std::shared_mutex> lock(idx->mu);
  ScratchLease sc(idx);
  ISL_TRY(sc.status);
  ISL_TRY(order_after_caller(sc->stream, sc->ev_in));
  const float* q;
  uint32_t q_ld;
  ISL_TRY(stage_device_queries(idx, sc.get(), d_queries, nq, query_dim, &q, &q_ld));
  ShardOut so;
  so.packed = reinterpret_cast<uint4*>(d_records);
  so.id_base = id_base;
  ISL_TRY(search_device(idx, sc.get(), q, q_ld, nq, k, ef, nullptr, nullptr, nullptr, nullptr, &so));
  return search_finish(idx, sc.get(), 1);
} ISL_ABI_GUARD

isl_status isl_merge_packed_dev(const isl_shard_record* d_records, uint32_t parts, uint64_t nq, uint32_t k,
                                uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_count) try {
  if (nq == 0 || k == 0) return ISL_OK;
  if (!d_records || !d_out_ids || !d_out_dist) return fail(ISL_INVALID_ARGUMENT, "null pointer");
  if (parts == 0) return fail(ISL_INVALID_ARGUMENT, "parts must be > 0");
  int device, sms;
  ISL_TRY(current_device(&device, &sms));
  cudaStream_t st = caller_stream();
  ISL_TRY(launch_merge_packed(reinterpret_cast<const uint4*>(d_records), parts, nq * k, nq, k, d_out_ids, d_out_dist, d_out_count, st));
  ISL_CUDA_TRY(cudaStreamSynchronize(st));
  return ISL_OK;
} ISL_ABI_GUARD

}  // extern "C"
