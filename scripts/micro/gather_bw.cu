// Micro-benchmark (not product): achievable HBM bandwidth of random 3 KB row gathers on B200
// for different request shapes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// mode A: each warp reads whole random rows (ld floats) with coalesced 16B loads, UNROLL rows in flight
template <int UNROLL>
__global__ void gather_rows(const float4* __restrict__ v, uint32_t n, uint32_t ld4, uint32_t rows_per_warp, float* out) {
  const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  for (uint32_t r = 0; r < rows_per_warp; r += UNROLL) {
    float4 x[UNROLL][6];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t row = hash32(warp * 7919u + r + u) % n;
      const float4* p = v + (size_t)row * ld4;
#pragma unroll
      for (int j = 0; j < 6; ++j) x[u][j] = __ldg(p + j * 32 + lane);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc += x[u][j].x + x[u][j].y + x[u][j].z + x[u][j].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

// mode B: like the search kernel: 32 random rows per pass, CHB bytes of each row per stage via cp.async, STAGES deep
template <int CHF, int STAGES>
__global__ void __launch_bounds__(32) gather_chunks(const float* __restrict__ v, uint32_t n, uint32_t ld, uint32_t passes, float* out) {
  extern __shared__ __align__(16) float smem[];
  const uint32_t lane = threadIdx.x;
  constexpr int RS = CHF + 4;
  float acc = 0.f;
  __shared__ uint32_t ids[32];
  for (uint32_t p = 0; p < passes; ++p) {
    ids[lane] = hash32(blockIdx.x * 104729u + p * 32 + lane) % n;
    __syncwarp();
    const uint32_t nch = ld / CHF;
    auto issue = [&](uint32_t c) {
      if (c < nch) {
        float* buf = smem + (c % STAGES) * 32 * RS;
        for (uint32_t idx = lane; idx < 32 * (CHF / 4); idx += 32) {
          const uint32_t row = idx / (CHF / 4), w = idx % (CHF / 4);
          const float* src = v + (size_t)ids[row] * ld + c * CHF + w * 4;
          uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + row * RS + w * 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (uint32_t c = 0; c < nch; ++c) {
      issue(c + STAGES - 1);
      asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 1));
      __syncwarp();
      const float4* row = reinterpret_cast<const float4*>(smem + (c % STAGES) * 32 * RS + lane * RS);
#pragma unroll
      for (int w = 0; w < CHF / 4; ++w) { float4 y = row[w]; acc += y.x + y.y + y.z + y.w; }
      __syncwarp();
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <class F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  const uint32_t n = 1000000, ld = 768;
  float* v; float* out;
  CK(cudaMalloc(&v, (size_t)n * ld * 4)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(v, 0, (size_t)n * ld * 4));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  {  // A: whole rows
    const uint32_t rows_per_warp = 256;
    for (int wps : {8, 16, 32, 64}) {
      const uint32_t blocks = sms * wps / 8;  // 256 threads/block = 8 warps
      auto run1 = [&]() { gather_rows<1><<<blocks, 256>>>((const float4*)v, n, ld / 4, rows_per_warp, out); };
      auto run2 = [&]() { gather_rows<2><<<blocks, 256>>>((const float4*)v, n, ld / 4, rows_per_warp, out); };
      auto run4 = [&]() { gather_rows<4><<<blocks, 256>>>((const float4*)v, n, ld / 4, rows_per_warp, out); };
      const double bytes = (double)blocks * 8 * rows_per_warp * ld * 4;
      printf("rows  warps/SM=%2d  unroll1 %.0f GB/s  unroll2 %.0f GB/s  unroll4 %.0f GB/s\n", wps, bytes / time_ms(run1) / 1e6,
             bytes / time_ms(run2) / 1e6, bytes / time_ms(run4) / 1e6);
    }
  }
  auto runB = [&](auto kern, int chf, int stages, const char* name) {
    const size_t smem = (size_t)stages * 32 * (chf + 4) * 4;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    for (int cap : {7, 12, 16, 24, 32}) {
      const int w = per_sm < cap ? per_sm : cap;
      const uint32_t blocks = sms * w, passes = 64;
      auto run = [&]() { kern<<<blocks, 32, smem>>>(v, n, ld, passes, out); };
      const double bytes = (double)blocks * passes * 32 * ld * 4;
      printf("%s warps/SM=%2d (max %d, smem %zu)  %.0f GB/s\n", name, w, per_sm, smem, bytes / time_ms(run) / 1e6);
      if (w == per_sm) break;
    }
  };
  runB(gather_chunks<64, 3>, 64, 3, "chunk 256B x3 stages");
  runB(gather_chunks<64, 4>, 64, 4, "chunk 256B x4 stages");
  runB(gather_chunks<32, 4>, 32, 4, "chunk 128B x4 stages");
  runB(gather_chunks<32, 8>, 32, 8, "chunk 128B x8 stages");
  runB(gather_chunks<128, 2>, 128, 2, "chunk 512B x2 stages");
  runB(gather_chunks<128, 3>, 128, 3, "chunk 512B x3 stages");
  runB(gather_chunks<256, 2>, 256, 2, "chunk 1KB  x2 stages");
  return 0;
}
