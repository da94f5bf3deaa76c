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
  // NC x KS tile, 16-byte global loads (all issued before the first store), zero beyond N
  constexpr int VE = 16 / sizeof(TF);
  constexpr int FP = Pitch<TF>::v;
  constexpr int PER = NC * (KS / VE) / HT;
  static_assert(NC * (KS / VE) % HT == 0, "tile must divide over the CTA");
  uint4 v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int e = threadIdx.x + i * HT;
    const int r = e / (KS / VE), c = (e - r * (KS / VE)) * VE;
    v[i] = make_uint4(0, 0, 0, 0);
    if (n0 + r < N) v[i] = __ldg(reinterpret_cast<const uint4*>(feat + (int64_t)(n0 + r) * F + k0 + c));
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int e = threadIdx.x + i * HT;
    const int r = e / (KS / VE), c = (e - r * (KS / VE)) * VE;
    if constexpr (sizeof(TF) == 2) {
      uint32_t* d = reinterpret_cast<uint32_t*>(fs + r * FP + c);          // (r*66 + c) is even
      d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w;
    } else {
      const TF* pv = reinterpret_cast<const TF*>(&v[i]);
#pragma unroll
      for (int j = 0; j < VE; ++j) fs[r * FP + c + j] = pv[j];
    }
  }
}

__device__ __forceinline__ void stage_w(const float* __restrict__ W0, int U0, const float* __restrict__ W1, int U1,
                                        float* ws, int k0) {
  // ws[k][u]: columns [0,U0) from W0, [U0,U0+U1) from W1, zero padding up to UP.  All loads are issued
  // before the first store (a load-store-load-store chain pays one memory round trip per element).
  constexpr int PER = KS * UP / HT;
  static_assert(KS * UP % HT == 0, "tile must divide over the CTA");
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int e = threadIdx.x + i * HT;
    const int k = e / UP, u = e - k * UP;
    v[i] = 0.f;
    if (u < U0) v[i] = __ldg(W0 + (int64_t)(k0 + k) * U0 + u);
    else if (u < U0 + U1) v[i] = __ldg(W1 + (int64_t)(k0 + k) * U1 + (u - U0));
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) ws[threadIdx.x + i * HT] = v[i];
}

// dls[r][u] (pitch DP): logit gradients of samples n0 .. n0+NC-1, columns as in stage_w, zero padded
template <int DP>
__device__ __forceinline__ void stage_dl(const float* __restrict__ dl0, int U0, const float* __restrict__ dl1, int U1,
                                         float* dls, int n0, int N) {
  constexpr int PER = NC * UP / HT, HALF = PER / 2;
  static_assert(NC * UP % (2 * HT) == 0, "tile must divide over the CTA");
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[HALF];
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      const int e = threadIdx.x + (h * HALF + i) * HT;
      const int r = e / UP, u = e - r * UP;
      v[i] = 0.f;
      if (n0 + r < N) {
        if (u < U0) { if (dl0) v[i] = __ldg(dl0 + (int64_t)(n0 + r) * U0 + u); }
        else if (u < U0 + U1) { if (dl1) v[i] = __ldg(dl1 + (int64_t)(n0 + r) * U1 + (u - U0)); }
      }
    }
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      const int e = threadIdx.x + (h * HALF + i) * HT;
      const int r = e / UP, u = e - r * UP;
      dls[r * DP + u] = v[i];
    }
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
  const int U = U0 + U1;
  const int nslices = F / KS;
  // thread tile: 4 samples (ns, ns+32, ns+64, ns+96) x 6 columns (warp-uniform column group); a CTA walks
  // its feature slices with the accumulators in registers and adds them to the workspace once per pass
  const int ns = threadIdx.x & 31, ug = threadIdx.x >> 5;
  for (int n0 = 0; n0 < N; n0 += NC) {
    float acc[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
      const int k0 = sl * KS;
      __syncthreads();
      stage_w(W0, U0, W1, U1, ws, k0);
      stage_feat<TF>(feat, fs, n0, N, F, k0);
      __syncthreads();
      constexpr int KSTEP = sizeof(TF) == 2 ? 2 : 1;
#pragma unroll 2
      for (int k = 0; k < KS; k += KSTEP) {
        float w[KSTEP][6];
#pragma unroll
        for (int kk = 0; kk < KSTEP; ++kk) {
          const float2 w01 = *reinterpret_cast<const float2*>(ws + (k + kk) * UP + ug * 6);
          const float2 w23 = *reinterpret_cast<const float2*>(ws + (k + kk) * UP + ug * 6 + 2);
          const float2 w45 = *reinterpret_cast<const float2*>(ws + (k + kk) * UP + ug * 6 + 4);
          w[kk][0] = w01.x; w[kk][1] = w01.y; w[kk][2] = w23.x; w[kk][3] = w23.y; w[kk][4] = w45.x; w[kk][5] = w45.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[KSTEP];
          if constexpr (sizeof(TF) == 2) {
            const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(fs + (ns + 32 * i) * FP + k));
            f[0] = f2.x; f[1] = f2.y;
          } else {
            f[0] = to_f(fs[(ns + 32 * i) * FP + k]);
          }
#pragma unroll
          for (int kk = 0; kk < KSTEP; ++kk)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(f[kk], w[kk][j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ns + 32 * i;
      if (n >= N) continue;
      float* dst = ws_acc + (int64_t)n * UP + ug * 6;          // padded columns stay zero: sums of zeros
#pragma unroll
      for (int j = 0; j < 6; j += 2)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + j), "f"(acc[i][j]), "f"(acc[i][j + 1]) : "memory");
    }
  }
  // last CTA: bias + activation, write the two outputs, leave the workspace zeroed for the next call
  asm volatile("fence.acq_rel.gpu;" ::: "memory");     // release of this thread's reductions (lighter than fence.sc)
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int e0 = threadIdx.x; e0 < N * UP; e0 += 4 * HT) {
    float v4[4];                                 // four independent L2 reads in flight per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * HT;
      v4[i] = e < N * UP ? __ldcg(&ws_acc[e]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * HT;
      if (e >= N * UP) continue;
      ws_acc[e] = 0.f;
      const int n = e / UP, u = e - n * UP;
      if (u >= U) continue;
      float v = v4[i] + ((u < U0) ? (b0 ? b0[u] : 0.f) : (b1 ? b1[u - U0] : 0.f));
      if (act == LG_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      else if (act == LG_ACT_TANH) v = tanhf(v);
      if (u < U0) out0[(int64_t)n * U0 + u] = v;
      else out1[(int64_t)n * U1 + (u - U0)] = v;
    }
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
  constexpr int DP = UP + 4;                                 // 52 floats: 16-byte rows, conflict-free LDS.128
  constexpr int FP = Pitch<TF>::v;
  float* ws = reinterpret_cast<float*>(hsm);                 // [KS][UP]
  float* dls = ws + KS * UP;                                 // [NC][DP]
  TF* fs = reinterpret_cast<TF*>(dls + NC * DP);             // [NC][FP]   (weight gradient only)
  const int k0 = blockIdx.x * KS;
  const int U = U0 + U1;
  const bool wgrad = dW0 != nullptr || dW1 != nullptr;
  stage_w(W0, U0, W1, U1, ws, k0);

  // d(feat) tile: 8 samples (ns + 16 i) x 4 features (kg*4 .. +3) per thread, columns four at a time
  const int ns = threadIdx.x & 15, kg = threadIdx.x >> 4;
  // dW tile: feature kw, columns uh*12 .. +11
  const int kw = threadIdx.x & 63, uh = threadIdx.x >> 6;
  float wacc[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) wacc[j] = 0.f;
  float bacc = 0.f;                                          // CTA 0: bias gradient, thread u < U

  for (int n0 = 0; n0 < N; n0 += NC) {
    __syncthreads();
    stage_dl<DP>(dl0, U0, dl1, U1, dls, n0, N);
    if (wgrad) stage_feat<TF>(feat, fs, n0, N, F, k0);
    __syncthreads();

    if (dfeat != nullptr) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int u = 0; u < U; u += 4) {                       // columns >= U are zero in both operands
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(ws + (kg * 4 + j) * UP + u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 d = *reinterpret_cast<const float4*>(dls + (ns + 16 * i) * DP + u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[i][j] = fmaf(d.x, w[j].x, fmaf(d.y, w[j].y, fmaf(d.z, w[j].z, fmaf(d.w, w[j].w, acc[i][j]))));
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
        const float4 d0 = *reinterpret_cast<const float4*>(dls + r * DP + uh * 12);
        const float4 d1 = *reinterpret_cast<const float4*>(dls + r * DP + uh * 12 + 4);
        const float4 d2 = *reinterpret_cast<const float4*>(dls + r * DP + uh * 12 + 8);
        wacc[0] = fmaf(f, d0.x, wacc[0]); wacc[1] = fmaf(f, d0.y, wacc[1]);
        wacc[2] = fmaf(f, d0.z, wacc[2]); wacc[3] = fmaf(f, d0.w, wacc[3]);
        wacc[4] = fmaf(f, d1.x, wacc[4]); wacc[5] = fmaf(f, d1.y, wacc[5]);
        wacc[6] = fmaf(f, d1.z, wacc[6]); wacc[7] = fmaf(f, d1.w, wacc[7]);
        wacc[8] = fmaf(f, d2.x, wacc[8]); wacc[9] = fmaf(f, d2.y, wacc[9]);
        wacc[10] = fmaf(f, d2.z, wacc[10]); wacc[11] = fmaf(f, d2.w, wacc[11]);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < U) {
      const int rows = min(NC, N - n0);
      for (int r = 0; r < rows; ++r) bacc += dls[r * DP + threadIdx.x];
    }
  }
  if (wgrad) {
    // dW += wacc: all twelve loads in flight before the first add (a load-add-store chain per element
    // pays one global round trip per element: it was half of this kernel's time)
    float* q[12];
    float old[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int u = uh * 12 + j;
      q[j] = nullptr;
      if (u < U0) { if (dW0) q[j] = dW0 + (int64_t)(k0 + kw) * U0 + u; }
      else if (u < U) { if (dW1) q[j] = dW1 + (int64_t)(k0 + kw) * U1 + (u - U0); }
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) old[j] = q[j] ? *q[j] : 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) if (q[j]) *q[j] = old[j] + wacc[j];
  }
  if (blockIdx.x == 0 && threadIdx.x < U) {
    const int u = threadIdx.x;
    if (u < U0) { if (db0) db0[u] += bacc; }
    else if (db1) db1[u - U0] += bacc;
  }
}

size_t fwd_smem(int esz) { return (size_t)KS * UP * 4 + (size_t)NC * (KS + 2) * esz; }
size_t bwd_smem(int esz) { return (size_t)KS * UP * 4 + (size_t)NC * (UP + 4) * 4 + (size_t)NC * (KS + 2) * esz; }

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
  const int grid = F / KS < 2 * lg_num_sms() ? F / KS : 2 * lg_num_sms();   // two CTAs per SM: loads overlap FMAs
  if (dtype == LG_BF16) {
    heads_fwd_kernel<bf16><<<grid, HT, fwd_smem(2), st>>>((const bf16*)feat, W0, b0, U0, W1, b1, U1, out0, out1, N, F,
                                                           act, acc, counter);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(heads_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr = true; }
    heads_fwd_kernel<float><<<grid, HT, fwd_smem(4), st>>>((const float*)feat, W0, b0, U0, W1, b1, U1, out0, out1, N,
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
