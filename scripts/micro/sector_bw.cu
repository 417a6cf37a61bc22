// Micro-benchmark (not product): what the B200 memory system sustains for the ADC traversal's access
// shape — independent random 32-byte sector reads, and random 4-byte atomicOr-with-return (visited
// bits) over regions smaller and larger than L2.  The byte roofline (copy bandwidth) is the wrong
// denominator for that kernel: every access moves one 32-byte sector, and an atomic on a line that
// is not L2 resident costs a DRAM read plus a DRAM write-back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 sector_bw.cu -o sector_bw
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// every lane reads UNROLL independent random 32-byte rows (two 16-byte loads each) per iteration
template <int UNROLL>
__global__ void __launch_bounds__(256) read_sectors(const uint4* __restrict__ t, uint32_t rows, uint32_t iters, uint32_t* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < iters; ++i) {
    uint4 a[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t r = hash32(tid * 2654435761u + i * UNROLL + u) % rows;
      a[u] = __ldg(t + (size_t)r * 2);
      b[u] = __ldg(t + (size_t)r * 2 + 1);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += a[u].x ^ b[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// every lane issues UNROLL random atomicOr (result used) per iteration; each warp owns a private
// region of `words_per_warp` words (the per-query visited bitset) inside `region`
template <int UNROLL>
__global__ void __launch_bounds__(32) atomics(uint32_t* __restrict__ region, uint32_t words_per_warp, uint32_t iters, uint32_t* out) {
  uint32_t* mine = region + (size_t)blockIdx.x * words_per_warp;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < iters; ++i) {
    uint32_t o[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t h = hash32((blockIdx.x * 32 + threadIdx.x) * 2654435761u + i * UNROLL + u);
      o[u] = atomicOr(mine + (h >> 5) % words_per_warp, 1u << (h & 31));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += o[u];
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// the hop shape: 2 atomics + 2 code rows (32 B) per lane, results needed before the next hop
__global__ void __launch_bounds__(32) hop_shape(uint32_t* __restrict__ region, uint32_t words_per_warp, const uint4* __restrict__ codes,
                                                uint32_t n, uint32_t hops, uint32_t* out) {
  uint32_t* mine = region + (size_t)blockIdx.x * words_per_warp;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < hops; ++i) {
    uint32_t o[2];
    uint4 c[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t id = hash32((blockIdx.x * 32 + threadIdx.x) * 2654435761u + i * 2 + u + (acc & 1)) % n;  // depends on the previous hop
      o[u] = atomicOr(mine + (id >> 5), 1u << (id & 31));
      c[u][0] = __ldg(codes + (size_t)id * 2);
      c[u][1] = __ldg(codes + (size_t)id * 2 + 1);
    }
    acc += (o[0] ^ o[1] ^ c[0][0].x ^ c[0][1].y ^ c[1][0].z ^ c[1][1].w) | 1u;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <class F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; CK(cudaMalloc(&out, 4));
  {  // random 32-byte reads: 4 GB table (DRAM) and 32 MB table (L2 resident, the PQ code table at 1M x m=32)
    for (size_t mb : {4096, 32}) {
      const uint32_t rows = (uint32_t)(mb * 1024 * 1024 / 32);
      uint4* t; CK(cudaMalloc(&t, (size_t)rows * 32)); CK(cudaMemset(t, 1, (size_t)rows * 32));
      for (int wps : {16, 32, 64}) {
        const uint32_t blocks = sms * wps / 8, iters = 64;
        auto r4 = [&]() { read_sectors<4><<<blocks, 256>>>(t, rows, iters, out); };
        auto r8 = [&]() { read_sectors<8><<<blocks, 256>>>(t, rows, iters, out); };
        const double ops = (double)blocks * 256 * iters;
        printf("read 32B  table %4zu MB  warps/SM=%2d  4 in flight/lane: %6.1f G sectors/s (%5.0f GB/s)   8 in flight: %6.1f G sectors/s (%5.0f GB/s)\n",
               mb, wps, ops * 4 / time_ms(r4) / 1e6, ops * 4 * 32 / time_ms(r4) / 1e6, ops * 8 / time_ms(r8) / 1e6, ops * 8 * 32 / time_ms(r8) / 1e6);
      }
      CK(cudaFree(t));
    }
  }
  {  // atomicOr with return: per-warp bitsets of 125 KB (1M nodes), total footprint by resident warps
    const uint32_t words = 31250;
    for (int wps : {2, 4, 6, 12, 16, 32}) {
      const uint32_t blocks = sms * wps, iters = 256;
      uint32_t* region; CK(cudaMalloc(&region, (size_t)blocks * words * 4)); CK(cudaMemset(region, 0, (size_t)blocks * words * 4));
      auto a2 = [&]() { atomics<2><<<blocks, 32>>>(region, words, iters, out); };
      auto a8 = [&]() { atomics<8><<<blocks, 32>>>(region, words, iters, out); };
      const double ops = (double)blocks * 32 * iters;
      printf("atomicOr  warps/SM=%2d  bitsets %6.1f MB  2 in flight/lane: %6.2f G atomics/s   8 in flight: %6.2f G atomics/s\n", wps,
             (double)blocks * words * 4 / 1e6, ops * 2 / time_ms(a2) / 1e6, ops * 8 / time_ms(a8) / 1e6);
      CK(cudaFree(region));
    }
  }
  {  // the dependent hop shape at the traversal kernel's residency
    const uint32_t n = 1000000, words = 31250;
    uint4* codes; CK(cudaMalloc(&codes, (size_t)n * 32)); CK(cudaMemset(codes, 1, (size_t)n * 32));
    for (int wps : {4, 6, 12, 16, 24, 32}) {
      const uint32_t blocks = sms * wps, hops = 512;
      uint32_t* region; CK(cudaMalloc(&region, (size_t)blocks * words * 4)); CK(cudaMemset(region, 0, (size_t)blocks * words * 4));
      auto h = [&]() { hop_shape<<<blocks, 32>>>(region, words, codes, n, hops, out); };
      const float ms = time_ms(h);
      printf("hop shape warps/SM=%2d  bitsets %6.1f MB + codes 32 MB: %7.2f us per hop per warp, %6.2f G node visits/s\n", wps,
             (double)blocks * words * 4 / 1e6, ms * 1e3 / hops, (double)blocks * hops * 64 / ms / 1e6);
      CK(cudaFree(region));
    }
    CK(cudaFree(codes));
  }
  return 0;
}
