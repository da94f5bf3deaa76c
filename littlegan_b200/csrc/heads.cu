// The discriminator's two sigmoid Dense heads (model.py:62-63,70-72) as ONE pass over the flattened
// encoder features:  out_h = sigmoid(f @ W_h + b_h),  h = pr (1 unit), cond (cond_dim units).
//
// Shapes: f [N, F = 24576] (bf16 or fp32), W_pr [F,1], W_c [F,cond] fp32; U = 1 + cond <= 48 columns.
// 2*N*F*U FLOPs (0.26 GFLOP at N = 128, U = 41) against 10 MB of operands: the two heads used to be six
// generic GEMM launches (split-K atomics, the 6 MB feature map read twice, fp32 d(feat) + a cast).  Here
// the feature axis is sliced over the grid (KS = 64 features per CTA); a CTA stages its W slice, the
// feature slice and (backward) the logit gradients in shared memory and produces with fp32 FMAs
//
//   forward : partial f[:, slice] @ W[slice, :] -> atomics into a workspace; the LAST CTA to finish adds
//             the bias, applies the sigmoid, writes both outputs and re-zeroes the workspace.
//   backward: d(feat)[:, slice] = dl @ W[slice, :]^T  (written once, in the activation dtype) and
//             dW[slice, :] += f[:, slice]^T @ dl  (owned by the CTA: no atomics), bias gradient by CTA 0.
#include "common.cuh"

namespace {

constexpr int HT = 256;        // threads
constexpr int KS = 64;         // feature slice per CTA
constexpr int NC = 128;        // samples per pass
constexpr int UP = 48;         // padded head columns
// feature tile pitch (elements): an odd number of 32-bit words per row => conflict-free column walks
template <typename TF> struct Pitch { static constexpr int v = KS + (sizeof(TF) == 2 ? 2 : 1); };

template <typename TF>
__device__ __forceinline__ void stage_feat(const TF* __restrict__ feat, TF* fs, int n0, int N, int F, int k0) {
  // NC x KS tile, 16-byte global loads, zero beyond N
  constexpr int VE = 16 / sizeof(TF);
  for (int e = threadIdx.x; e < NC * (KS / VE); e += HT) {
    const int r = e / (KS / VE), c = (e - r * (KS / VE)) * VE;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n0 + r < N) v = *reinterpret_cast<const uint4*>(feat + (int64_t)(n0 + r) * F + k0 + c);
    constexpr int FP = Pitch<TF>::v;
    if constexpr (sizeof(TF) == 2) {
      uint32_t* d = reinterpret_cast<uint32_t*>(fs + r * FP + c);          // (r*66 + c) is even
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    } else {
      const TF* pv = reinterpret_cast<const TF*>(&v);
#pragma unroll
      for (int i = 0; i < VE; ++i) fs[r * FP + c + i] = pv[i];
    }
  }
}

__device__ __forceinline__ void stage_w(const float* __restrict__ W0, int U0, const float* __restrict__ W1, int U1,
                                        float* ws, int k0) {
  // ws[k][u]: columns [0,U0) from W0, [U0,U0+U1) from W1, zero padding up to UP
  for (int e = threadIdx.x; e < KS * UP; e += HT) {
    const int k = e / UP, u = e - k * UP;
    float v = 0.f;
    if (u < U0) v = W0[(int64_t)(k0 + k) * U0 + u];
    else if (u < U0 + U1) v = W1[(int64_t)(k0 + k) * U1 + (u - U0)];
    ws[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename TF>
__global__ void __launch_bounds__(HT) heads_fwd_kernel(const TF* __restrict__ feat, const float* __restrict__ W0,
                                                       const float* __restrict__ b0, int U0,
                                                       const float* __restrict__ W1, const float* __restrict__ b1,
                                                       int U1, float* __restrict__ out0, float* __restrict__ out1,
                                                       int N, int F, int act, float* ws_acc, unsigned* counter) {
  extern __shared__ __align__(16) uint8_t hsm[];
  float* ws = reinterpret_cast<float*>(hsm);                 // [KS][UP]
  TF* fs = reinterpret_cast<TF*>(ws + KS * UP);              // [NC][FP]
  __shared__ bool last;
  constexpr int FP = Pitch<TF>::v;
  const int k0 = blockIdx.x * KS;
  const int U = U0 + U1;
  stage_w(W0, U0, W1, U1, ws, k0);
  // thread tile: 4 samples (ns, ns+32, ns+64, ns+96) x 6 columns (warp-uniform column group)
  const int ns = threadIdx.x & 31, ug = threadIdx.x >> 5;
  for (int n0 = 0; n0 < N; n0 += NC) {
    __syncthreads();
    stage_feat<TF>(feat, fs, n0, N, F, k0);
    __syncthreads();
    float acc[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < KS; ++k) {
      float w[6];
      const float2 w01 = *reinterpret_cast<const float2*>(ws + k * UP + ug * 6);
      const float2 w23 = *reinterpret_cast<const float2*>(ws + k * UP + ug * 6 + 2);
      const float2 w45 = *reinterpret_cast<const float2*>(ws + k * UP + ug * 6 + 4);
      w[0] = w01.x; w[1] = w01.y; w[2] = w23.x; w[3] = w23.y; w[4] = w45.x; w[5] = w45.y;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f = to_f(fs[(ns + 32 * i) * FP + k]);
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(f, w[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ns + 32 * i;
      if (n >= N) continue;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int u = ug * 6 + j;
        if (u < U) atomicAdd(&ws_acc[(int64_t)n * UP + u], acc[i][j]);
      }
    }
  }
  // last CTA: bias + activation, write the two outputs, leave the workspace zeroed for the next call
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int e = threadIdx.x; e < N * UP; e += HT) {
    const int n = e / UP, u = e - n * UP;
    if (u >= U) continue;
    float v = __ldcg(&ws_acc[e]);
    ws_acc[e] = 0.f;
    v += (u < U0) ? (b0 ? b0[u] : 0.f) : (b1 ? b1[u - U0] : 0.f);
    if (act == LG_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    else if (act == LG_ACT_TANH) v = tanhf(v);
    if (u < U0) out0[(int64_t)n * U0 + u] = v;
    else out1[(int64_t)n * U1 + (u - U0)] = v;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// ------------------------------------------------------------------------------------------------
// backward: dl_h = gradient w.r.t. the pre-activation logits of head h (either may be NULL = zero)
// ------------------------------------------------------------------------------------------------
template <typename TF>
__global__ void __launch_bounds__(HT) heads_bwd_kernel(const TF* __restrict__ feat, const float* __restrict__ dl0,
                                                       const float* __restrict__ dl1, const float* __restrict__ W0,
                                                       int U0, const float* __restrict__ W1, int U1,
                                                       TF* __restrict__ dfeat, float* dW0, float* dW1, float* db0,
                                                       float* db1, int N, int F) {
  extern __shared__ __align__(16) uint8_t hsm[];
  float* ws = reinterpret_cast<float*>(hsm);                 // [KS][UP]
  float* dls = ws + KS * UP;                                 // [NC][UP+1]
  TF* fs = reinterpret_cast<TF*>(dls + NC * (UP + 1));       // [NC][FP]   (weight gradient only)
  constexpr int DP = UP + 1;
  constexpr int FP = Pitch<TF>::v;
  const int k0 = blockIdx.x * KS;
  const int U = U0 + U1;
  const bool wgrad = dW0 != nullptr || dW1 != nullptr;
  stage_w(W0, U0, W1, U1, ws, k0);

  // d(feat) tile: 8 samples (ns + 16 i) x 4 features (kg*4 .. +3) per thread
  const int ns = threadIdx.x & 15, kg = threadIdx.x >> 4;
  // dW tile: feature kw, columns uh*24 .. +23
  const int kw = threadIdx.x & 63, uh = threadIdx.x >> 6;     // 4 column groups of 12
  float wacc[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) wacc[j] = 0.f;
  float bacc = 0.f;                                          // CTA 0: bias gradient, thread u < U

  for (int n0 = 0; n0 < N; n0 += NC) {
    __syncthreads();
    for (int e = threadIdx.x; e < NC * UP; e += HT) {
      const int r = e / UP, u = e - r * UP;
      float v = 0.f;
      if (n0 + r < N) {
        if (u < U0) v = dl0 ? dl0[(int64_t)(n0 + r) * U0 + u] : 0.f;
        else if (u < U) v = dl1 ? dl1[(int64_t)(n0 + r) * U1 + (u - U0)] : 0.f;
      }
      dls[r * DP + u] = v;
    }
    if (wgrad) stage_feat<TF>(feat, fs, n0, N, F, k0);
    __syncthreads();

    if (dfeat != nullptr) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int u = 0; u < U; ++u) {
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = ws[(kg * 4 + j) * UP + u];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = dls[(ns + 16 * i) * DP + u];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d, w[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n0 + ns + 16 * i;
        if (n >= N) continue;
        TF* dst = dfeat + (int64_t)n * F + k0 + kg * 4;
        if constexpr (sizeof(TF) == 2) {
          uint2 v;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
          h[0] = __floats2bfloat162_rn(acc[i][0], acc[i][1]);
          h[1] = __floats2bfloat162_rn(acc[i][2], acc[i][3]);
          *reinterpret_cast<uint2*>(dst) = v;
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
      }
    }
    if (wgrad) {
      const int rows = min(NC, N - n0);
      for (int r = 0; r < rows; ++r) {
        const float f = to_f(fs[r * FP + kw]);
#pragma unroll
        for (int j = 0; j < 12; ++j) wacc[j] = fmaf(f, dls[r * DP + uh * 12 + j], wacc[j]);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < U) {
      const int rows = min(NC, N - n0);
      for (int r = 0; r < rows; ++r) bacc += dls[r * DP + threadIdx.x];
    }
  }
  if (wgrad) {
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int u = uh * 12 + j;
      if (u < U0) { if (dW0) dW0[(int64_t)(k0 + kw) * U0 + u] += wacc[j]; }
      else if (u < U) { if (dW1) dW1[(int64_t)(k0 + kw) * U1 + (u - U0)] += wacc[j]; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < U) {
    const int u = threadIdx.x;
    if (u < U0) { if (db0) db0[u] += bacc; }
    else if (db1) db1[u - U0] += bacc;
  }
}

size_t fwd_smem(int esz) { return (size_t)KS * UP * 4 + (size_t)NC * (KS + 2) * esz; }
size_t bwd_smem(int esz) { return (size_t)KS * UP * 4 + (size_t)NC * (UP + 1) * 4 + (size_t)NC * (KS + 2) * esz; }

}  // namespace

extern "C" int64_t lg_dense_heads_workspace_bytes(int N) {
  if (N <= 0) return LG_ERR_INVALID;
  return (int64_t)N * UP * 4 + 256;
}

extern "C" int lg_dense_heads_fwd(const void* feat, const float* W0, const float* b0, int U0, const float* W1,
                                  const float* b1, int U1, float* out0, float* out1, int N, int F, int act,
                                  int dtype, void* workspace, void* stream) {
  LG_REQUIRE(feat && W0 && W1 && out0 && out1 && workspace, "NULL argument");
  LG_REQUIRE(N > 0 && U0 > 0 && U1 > 0, "bad shape");
  if (U0 + U1 > UP || F % KS != 0) { lg_set_error("lg_dense_heads_fwd: needs U0+U1 <= 48 and F %% 64 == 0"); return LG_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* counter = reinterpret_cast<unsigned*>(workspace);
  float* acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  if (dtype == LG_BF16) {
    heads_fwd_kernel<bf16><<<F / KS, HT, fwd_smem(2), st>>>((const bf16*)feat, W0, b0, U0, W1, b1, U1, out0, out1, N, F,
                                                           act, acc, counter);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(heads_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr = true; }
    heads_fwd_kernel<float><<<F / KS, HT, fwd_smem(4), st>>>((const float*)feat, W0, b0, U0, W1, b1, U1, out0, out1, N,
                                                            F, act, acc, counter);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_dense_heads_bwd(const void* feat, const float* dl0, const float* dl1, const float* W0, int U0,
                                  const float* W1, int U1, void* dfeat, float* dW0, float* dW1, float* db0, float* db1,
                                  int N, int F, int dtype, void* stream) {
  LG_REQUIRE(W0 && W1 && (dl0 || dl1), "NULL argument");
  LG_REQUIRE(N > 0 && U0 > 0 && U1 > 0, "bad shape");
  LG_REQUIRE(feat || (!dW0 && !dW1), "the weight gradient needs the features");
  if (U0 + U1 > UP || F % KS != 0) { lg_set_error("lg_dense_heads_bwd: needs U0+U1 <= 48 and F %% 64 == 0"); return LG_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(heads_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(heads_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  if (dtype == LG_BF16)
    heads_bwd_kernel<bf16><<<F / KS, HT, bwd_smem(2), st>>>((const bf16*)feat, dl0, dl1, W0, U0, W1, U1, (bf16*)dfeat, dW0,
                                                           dW1, db0, db1, N, F);
  else
    heads_bwd_kernel<float><<<F / KS, HT, bwd_smem(4), st>>>((const float*)feat, dl0, dl1, W0, U0, W1, U1, (float*)dfeat,
                                                            dW0, dW1, db0, db1, N, F);
  LG_LAUNCH_CHECK();
  return LG_OK;
}
