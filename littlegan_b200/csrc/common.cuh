// Shared helpers for the littlegan_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "littlegan_b200.h"

void lg_set_error(const char* fmt, ...);

#define LG_REQUIRE(cond, msg)                          \
  do {                                                 \
    if (!(cond)) {                                     \
      lg_set_error("%s: %s", __func__, msg);           \
      return LG_ERR_INVALID;                           \
    }                                                  \
  } while (0)

#define LG_LAUNCH_CHECK()                                                   \
  do {                                                                      \
    cudaError_t e__ = cudaPeekAtLastError();                                \
    if (e__ != cudaSuccess) {                                               \
      lg_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
      (void)cudaGetLastError();                                             \
      return LG_ERR_CUDA;                                                   \
    }                                                                       \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float leaky_f(float x, float alpha) { return x > 0.f ? x : alpha * x; }
__device__ __forceinline__ float leaky_d(float x, float alpha) { return x > 0.f ? 1.f : alpha; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of two doubles; result valid in thread 0.  `sh` needs 2*32 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh) {
  a = warp_sum(a);
  b = warp_sum(b);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[w] = a; sh[32 + w] = b; }
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? sh[lane] : 0.0;
    b = lane < nw ? sh[32 + lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
  __syncthreads();
}

// Grid of a persistent kernel that walks `items` work items round-robin over at most `max_ctas` CTAs: the
// smallest grid with the same number of rounds.  512 tiles on 148 SMs are 4 rounds either way; 128 CTAs do them
// with no idle tail and leave 20 SMs to the kernels of the step's other streams.
static inline int lg_even_grid(int items, int max_ctas) {
  if (items <= max_ctas) return items < 1 ? 1 : items;
  const int rounds = (items + max_ctas - 1) / max_ctas;
  return (items + rounds - 1) / rounds;
}

static inline int lg_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
