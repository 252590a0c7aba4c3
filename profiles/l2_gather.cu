// Ceiling of the SpMM's access pattern on this GPU: random gathers of whole fp32 rows (64 B .. 512 B) from a table of a
// given footprint, nothing else -- no index stream, no plan, no FMA chain, 8 independent 128-bit loads in flight per lane.
// usage: l2_gather [row_bytes=256] ; prints GB/s per table size.  Build: nvcc -arch=sm_100a -O3 -o profiles/bin/l2_gather profiles/l2_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int G>  // lanes per row (row = G * 16 bytes)
__global__ void __launch_bounds__(256, 4) gather_kernel(const float4* __restrict__ table, uint32_t n_rows, int iters, float4* out) {
  const int lane = threadIdx.x & 31, gl = lane & (G - 1);
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  float4 acc = make_float4(0, 0, 0, 0);
  uint32_t s = group * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s = mix(s + u + 1);
      const uint32_t r = (uint32_t)(((uint64_t)s * n_rows) >> 32);
      const float4* p = table + (size_t)r * G + gl;
      asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 1234.5f) out[0] = acc;
}

template <int G>
static void run(int row_bytes) {
  const double sizes_mb[] = {4, 16, 32, 48, 64, 80, 96, 128, 192, 512, 1536};
  float4* out;
  cudaMalloc(&out, 64);
  for (double mb : sizes_mb) {
    const size_t bytes = (size_t)(mb * (1 << 20));
    const uint32_t n_rows = (uint32_t)(bytes / row_bytes);
    float4* t;
    cudaMalloc(&t, bytes);
    cudaMemset(t, 0, bytes);
    const int blocks = 148 * 4 * 8, iters = 64;
    gather_kernel<G><<<blocks, 256>>>(t, n_rows, 8, out);  // warm
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      gather_kernel<G><<<blocks, 256>>>(t, n_rows, iters, out);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    const double moved = (double)blocks * 256 / G * iters * 8 * row_bytes;
    printf("row %3d B  table %6.0f MB : %7.1f GB/s  (%.3f ms)\n", row_bytes, mb, moved / best / 1e6, best);
    fflush(stdout);
    cudaFree(t);
  }
}

int main(int argc, char** argv) {
  const int rb = argc > 1 ? atoi(argv[1]) : 256;
  if (rb == 64) run<4>(rb); else if (rb == 128) run<8>(rb); else if (rb == 256) run<16>(rb); else run<32>(512);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
