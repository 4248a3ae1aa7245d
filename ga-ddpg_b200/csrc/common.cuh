// common.cuh — shared helpers for the gaddpg_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define GADDPG_OK 0
#define GADDPG_ERR_ARG (-1)
#define GADDPG_ERR_CUDA (-2)
#define GADDPG_ERR_UNSUPPORTED (-3)

void gaddpg_set_error(const char* fmt, ...);
extern long long g_gaddpg_launches;  // kernels launched by this library in this process (bench.py "gpu_launches")

#define GADDPG_CHECK_ARG(cond, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      gaddpg_set_error(__VA_ARGS__);     \
      return GADDPG_ERR_ARG;             \
    }                                    \
  } while (0)

// call after every launch: launch-configuration errors surface here (async faults surface at the
// caller's next synchronisation, as with any stream-ordered API)
#define GADDPG_CHECK_LAUNCH(name)                                                         \
  do {                                                                                    \
    ++g_gaddpg_launches;                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      gaddpg_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));     \
      return GADDPG_ERR_CUDA;                                                             \
    }                                                                                     \
  } while (0)

#define GADDPG_CUDA(call)                                                                 \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      gaddpg_set_error("%s failed: %s", #call, cudaGetErrorString(e__));                 \
      return GADDPG_ERR_CUDA;                                                             \
    }                                                                                     \
  } while (0)

static inline int gaddpg_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// Fixed-order sum of n values: s = ((v0 + v1) + v2) + ... exactly like the plain loop, but U loads are issued before the
// first add so the loop is not one exposed memory latency per term (the partial-sum reductions read L2-resident slots).
template <int U, typename F>
__device__ __forceinline__ float ordered_sum(int n, F load) {
  float s = 0.f;
  int i = 0;
  for (; i + U <= n; i += U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = load(i + u);
#pragma unroll
    for (int u = 0; u < U; ++u) s += v[u];
  }
  for (; i < n; ++i) s += load(i);
  return s;
}
