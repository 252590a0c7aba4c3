// Shared helpers for libb200rec.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/b200rec.h"

namespace b200rec {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define B2_REQUIRE(cond, msg)                                                        \
  do {                                                                               \
    if (!(cond)) return b200rec::fail(B200REC_ERR_ARG, "%s: %s", __func__, msg);     \
  } while (0)

#define B2_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return b200rec::fail(B200REC_ERR_CUDA, "%s: %s", __func__, cudaGetErrorString(e__));   \
  } while (0)

// call after every <<<>>> launch
#define B2_LAUNCHED()                                \
  do {                                               \
    b200rec::g_launches.fetch_add(1);                \
    B2_CUDA(cudaGetLastError());                     \
  } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ---------------------------------------------------------------------
// Kernels of one training step run back to back on one stream (and as consecutive nodes of the step graph).  Launched
// with the programmatic-serialization attribute, a kernel's blocks may become resident while the previous kernel's
// last wave drains: they read their STATIC inputs (work plan, CSR indices), then pdl_wait() until the previous grid has
// completed and its writes are visible.  Rules kept by every kernel launched through launch_pdl():
//   * pdl_wait() is executed by every block before the first access to anything another kernel of the step writes,
//     and before its own first global write (so completion stays transitive along the chain);
//   * pdl_trigger() only lets the NEXT grid start its prologue early; it promises nothing about this grid's data.
// B200REC_NO_PDL=1 launches everything fully serialised.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define B2_LAUNCH_PDL(...)                           \
  do {                                               \
    B2_CUDA(b200rec::launch_pdl(__VA_ARGS__));       \
    B2_LAUNCHED();                                   \
  } while (0)

// ---- device helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// coherent 128-bit load (never promoted to ld.global.nc): for tables the previous kernel of the step wrote, which a
// programmatically launched grid may only read after pdl_wait()
__device__ __forceinline__ float4 ldc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint8_t ldc_u8(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return (uint8_t)v;
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (evict-first) loads for data that is read exactly once (CSR arrays)
__device__ __forceinline__ int ld_stream_i32(const int32_t* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream_f32(const float* p) { return __ldcs(p); }

// streaming (evict-first) 128-bit accesses for rows that are touched once per launch (carried sums and outputs of a
// column-blocked pass): they must not push the pass's source block out of L2
__device__ __forceinline__ float4 ld_cs_f4(const float* ptr) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void st_cs_f4(float* ptr, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 128-bit fire-and-forget reduction (sm_90+): 4 fp32 adds in one L2 atomic transaction
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {  // butterfly inside aligned groups of G lanes
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (Salmon et al. 2011), the stated RNG of the device sampler; restated in oracle/oracle_c.c.
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ static inline void gen(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += W0;
      k1 += W1;
    }
  }
};

}  // namespace b200rec
