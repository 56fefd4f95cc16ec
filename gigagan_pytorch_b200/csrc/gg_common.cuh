// Shared helpers for libgigagan_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

#define GG_F32 0
#define GG_BF16 1

int gg_fail(const char* fmt, ...);          // records thread-local message, returns -1
int gg_check_launch(const char* what);      // cudaGetLastError -> 0 / -2

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

#define GG_DISPATCH(dtype, ...)                                         \
  switch (dtype) {                                                      \
    case GG_F32: { using T = float; __VA_ARGS__; break; }               \
    case GG_BF16: { using T = bf16; __VA_ARGS__; break; }               \
    default: return gg_fail("unsupported dtype %d", (int)(dtype));      \
  }

static inline int gg_cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline int gg_blocks(long n, int threads, int cap = 148 * 16) {
  long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
