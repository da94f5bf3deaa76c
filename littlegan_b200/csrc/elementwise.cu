// Bandwidth-bound kernels of the LittleGAN hot path: InstanceNormalization(axis=None) statistics /
// apply / backward, bias gradient, losses, TF-style Adam, casts and weight packing.
// All are coalesced, 16-byte vectorised, warp-shuffle reduced; HBM is the roofline.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int NT = 256;

// Vector load/store of V elements as float[V]: float x4 (16 B), bf16 x8 (16 B), bf16 x4 (8 B), or scalar.
template <typename T, int V>
__device__ __forceinline__ void ldv(const T* p, float* o) {
  if constexpr (V == 1) { o[0] = to_f(p[0]); }
  else if constexpr (sizeof(T) == 4) {
    static_assert(V == 4, "float vectors are 4 wide");
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else if constexpr (V == 8) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  } else {
    static_assert(V == 4, "bf16 vectors are 4 or 8 wide");
    uint2 v = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 2; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
}
template <typename T, int V>
__device__ __forceinline__ void stv(T* p, const float* o) {
  if constexpr (V == 1) { p[0] = from_f<T>(o[0]); }
  else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else if constexpr (V == 8) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  } else {
    uint2 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint2*>(p) = v;
  }
}

struct NormParams { float mu, inv, s_over_sigma; };
__device__ __forceinline__ NormParams norm_params(const double* stats, int n, int64_t M, float eps) {
  double mu = stats[2 * n] / (double)M;
  double var = stats[2 * n + 1] / (double)M - mu * mu;
  double sigma = var > 0.0 ? sqrt(var) : 0.0;
  double s = sigma + (double)eps;
  NormParams p;
  p.mu = (float)mu; p.inv = (float)(1.0 / s); p.s_over_sigma = sigma > 0.0 ? (float)(s / sigma) : 0.f;
  return p;
}

// ------------------------------------------------------------------------------------------------
// per-sample sum / sum of squares
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(NT) rowstats_kernel(const T* __restrict__ z, double* stats, int64_t M,
                                                      int64_t per_cta, float alpha_pre) {
  __shared__ double sh[64];
  const int n = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * per_cta, end = min(M, beg + per_cta);
  const T* zp = z + (int64_t)n * M;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = beg + (int64_t)threadIdx.x * V; i < end; i += (int64_t)NT * V) {
    float v[V];
    ldv<T, V>(zp + i, v);
#pragma unroll
    for (int k = 0; k < V; ++k) { float u = leaky_f(v[k], alpha_pre); s1 += u; s2 += u * u; }
  }
  double a = s1, b = s2;
  block_sum2(a, b, sh);
  if (threadIdx.x == 0) { atomicAdd(&stats[2 * n], a); atomicAdd(&stats[2 * n + 1], b); }
}

// ------------------------------------------------------------------------------------------------
// out = leaky_post(gamma * (pre(z) - mu) * inv + beta) [+ skip]
// ------------------------------------------------------------------------------------------------
template <typename TZ, typename T, int V>
__global__ void __launch_bounds__(NT) instnorm_fwd_kernel(const TZ* __restrict__ z, const double* __restrict__ stats,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const T* __restrict__ skip, T* __restrict__ out, int64_t M,
                                                          int64_t per_cta, float eps, float alpha_pre, float alpha_post) {
  const int n = blockIdx.y;
  const NormParams np = norm_params(stats, n, M, eps);
  const float g = gamma[0] * np.inv, b = beta[0] - gamma[0] * np.inv * np.mu;   // y = g*u + b
  const int64_t beg = (int64_t)blockIdx.x * per_cta, end = min(M, beg + per_cta);
  const int64_t base = (int64_t)n * M;
  for (int64_t i = beg + (int64_t)threadIdx.x * V; i < end; i += (int64_t)NT * V) {
    float v[V], o[V];
    ldv<TZ, V>(z + base + i, v);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = leaky_f(fmaf(g, leaky_f(v[k], alpha_pre), b), alpha_post);
    if (skip != nullptr) {
      float sk[V];
      ldv<T, V>(skip + base + i, sk);
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] += sk[k];
    }
    stv<T, V>(out + base + i, o);
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass 1: red[n] += (sum dy, sum dy*xhat)
// ------------------------------------------------------------------------------------------------
template <typename TZ, typename T, int V>
__global__ void __launch_bounds__(NT) instnorm_bwd_reduce_kernel(const T* __restrict__ gp, const TZ* __restrict__ z,
                                                                 const double* __restrict__ stats,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 double* red, int64_t M, int64_t per_cta, float eps,
                                                                 float alpha_pre, float alpha_post) {
  __shared__ double sh[64];
  const int n = blockIdx.y;
  const NormParams np = norm_params(stats, n, M, eps);
  const float ga = gamma[0], be = beta[0];
  const int64_t beg = (int64_t)blockIdx.x * per_cta, end = min(M, beg + per_cta);
  const int64_t base = (int64_t)n * M;
  float s1 = 0.f, s2 = 0.f;
  // xhat = fma(u, k1, k0), y = fma(u, y1, y0): 6 flops per element; two vectors in flight per thread
  const float k1 = np.inv, k0 = -np.mu * np.inv, y1 = ga * k1, y0 = fmaf(ga, k0, be);
  const bool pre = alpha_pre != 1.f;
  const int64_t step = (int64_t)NT * V;
  for (int64_t i = beg + (int64_t)threadIdx.x * V; i < end; i += 2 * step) {
    float v[2][V], g[2][V];
    const bool two = i + step < end;
    ldv<TZ, V>(z + base + i, v[0]);
    ldv<T, V>(gp + base + i, g[0]);
    if (two) { ldv<TZ, V>(z + base + i + step, v[1]); ldv<T, V>(gp + base + i + step, g[1]); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 0 || two) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float u = pre ? leaky_f(v[h][k], alpha_pre) : v[h][k];
          const float xh = fmaf(u, k1, k0);
          const float dy = g[h][k] * leaky_d(fmaf(u, y1, y0), alpha_post);
          s1 += dy; s2 = fmaf(dy, xh, s2);
        }
      }
    }
  }
  double a = s1, b = s2;
  block_sum2(a, b, sh);
  if (threadIdx.x == 0) { atomicAdd(&red[2 * n], a); atomicAdd(&red[2 * n + 1], b); }
}

// ------------------------------------------------------------------------------------------------
// backward pass 2
// ------------------------------------------------------------------------------------------------
// DB: also accumulate the per-channel sums of dz (the producing conv's bias gradient; channels innermost,
// C divides NT*V so a thread sees the same V channels in every iteration).  dy_ready: `gp` already holds
// dy = g * leaky_post'(y) (written by a conv epilogue, norm_bwd.cuh).
template <typename TZ, typename T, int V, bool DB>
__global__ void __launch_bounds__(NT) instnorm_bwd_apply_kernel(const T* __restrict__ gp, const TZ* __restrict__ z,
                                                                const double* __restrict__ stats, const double* __restrict__ red,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                TZ* __restrict__ dz, float* dgamma, float* dbeta, float* dbias,
                                                                int C, int dy_ready, int N, int64_t M, int64_t per_cta,
                                                                float eps, float alpha_pre, float alpha_post) {
  __shared__ float shc[DB ? NT * V : 1];
  const int n = blockIdx.y;
  if (blockIdx.x == 0 && n == 0 && threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < N; i += 32) { a += red[2 * i]; b += red[2 * i + 1]; }
    a = warp_sum(a); b = warp_sum(b);
    if (threadIdx.x == 0) {
      if (dbeta) atomicAdd(dbeta, (float)a);
      if (dgamma) atomicAdd(dgamma, (float)b);
    }
  }
  if constexpr (DB) {
    for (int c = threadIdx.x; c < C; c += NT) shc[c] = 0.f;
    __syncthreads();
  }
  const NormParams np = norm_params(stats, n, M, eps);
  const float ga = gamma[0], be = beta[0];
  const float mdy = (float)(red[2 * n] / (double)M);
  const float mdyx = (float)(red[2 * n + 1] / (double)M) * np.s_over_sigma;
  const float scale = ga * np.inv;
  const int64_t beg = (int64_t)blockIdx.x * per_cta, end = min(M, beg + per_cta);
  const int64_t base = (int64_t)n * M;
  // dz = pre'(z) * (c1*dy - c3 - c2*xhat) with xhat = fma(u,k1,k0), y = fma(u,y1,y0)
  const float k1 = np.inv, k0 = -np.mu * np.inv, y1 = ga * k1, y0 = fmaf(ga, k0, be);
  const float c1 = scale, c2 = mdyx * scale, c3 = mdy * scale;
  const bool pre = alpha_pre != 1.f;
  const float a_post = dy_ready ? 1.f : alpha_post;            // leaky_d(., 1) == 1: dy is taken as is
  const int64_t step = (int64_t)NT * V;
  float bsum[V];
#pragma unroll
  for (int k = 0; k < V; ++k) bsum[k] = 0.f;
  for (int64_t i = beg + (int64_t)threadIdx.x * V; i < end; i += 2 * step) {
    float v[2][V], g[2][V], o[V];
    const bool two = i + step < end;
    ldv<TZ, V>(z + base + i, v[0]);
    ldv<T, V>(gp + base + i, g[0]);
    if (two) { ldv<TZ, V>(z + base + i + step, v[1]); ldv<T, V>(gp + base + i + step, g[1]); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 0 || two) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float u = pre ? leaky_f(v[h][k], alpha_pre) : v[h][k];
          const float xh = fmaf(u, k1, k0);
          const float dy = g[h][k] * leaky_d(fmaf(u, y1, y0), a_post);
          const float r = fmaf(-xh, c2, fmaf(dy, c1, -c3));
          o[k] = pre ? r * leaky_d(v[h][k], alpha_pre) : r;
          if constexpr (DB) bsum[k] += o[k];
        }
        stv<TZ, V>(dz + base + i + h * step, o);
      }
    }
  }
  if constexpr (DB) {
    // lanes l and l + G (G = C/V channel groups per warp, a power of two) hold the same channels: fold them
    // with shuffles first, so a warp issues G*V shared atomics instead of 32*V
    const int G = C / V;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (o >= G) {
#pragma unroll
        for (int k = 0; k < V; ++k) bsum[k] += __shfl_xor_sync(0xffffffffu, bsum[k], o);
      }
    }
    const int c0 = (int)((beg + (int64_t)threadIdx.x * V) % C);
    if ((threadIdx.x & 31) < G || G >= 32) {
#pragma unroll
      for (int k = 0; k < V; ++k) atomicAdd(&shc[c0 + k], bsum[k]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += NT) atomicAdd(&dbias[c], shc[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// bias gradient: db[c] += sum_rows g[row][c]
// ------------------------------------------------------------------------------------------------
// PERIODIC: (NT*V) % C == 0 and C % V == 0, so a thread sees the same V columns every iteration.
template <typename T, int V, bool PERIODIC>
__global__ void __launch_bounds__(NT) bias_grad_kernel(const T* __restrict__ g, float* db, int64_t total, int C,
                                                       int64_t per_cta) {
  extern __shared__ float shc[];
  for (int c = threadIdx.x; c < C; c += NT) shc[c] = 0.f;
  __syncthreads();
  const int64_t beg = (int64_t)blockIdx.x * per_cta, end = min(total, beg + per_cta);
  if constexpr (PERIODIC) {
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const int64_t first = beg + (int64_t)threadIdx.x * V;
    for (int64_t i = first; i < end; i += (int64_t)NT * V) {
      float v[V];
      ldv<T, V>(g + i, v);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += v[k];
    }
    const int c0 = (int)(first % C);
#pragma unroll
    for (int k = 0; k < V; ++k) atomicAdd(&shc[c0 + k], acc[k]);
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += NT) atomicAdd(&shc[(int)(i % C)], to_f(g[i]));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) atomicAdd(&db[c], shc[c]);
}

// C % V == 0 but (NT*V) % C != 0 (C = 384, 192, ...): G = C/V threads span a row, R = NT/G rows per pass; every
// thread keeps its V columns in registers, the R row-lanes are folded through shared memory, one global atomic per
// column and CTA.  (The scalar fallback above does one SHARED atomic per element: 43 us for a 6 MB map.)
template <typename T, int V>
__global__ void __launch_bounds__(NT) bias_grad_group_kernel(const T* __restrict__ g, float* db, int64_t rows, int C,
                                                             int64_t rows_per) {
  extern __shared__ float shc[];                       // [R][C]
  const int G = C / V, R = NT / G;
  const int cg = threadIdx.x % G, rl = threadIdx.x / G;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(rows, r0 + rows_per);
  if (rl < R) {
    for (int64_t r = r0 + rl; r < r1; r += 2 * R) {
      float v0[V], v1[V];
      const bool two = r + R < r1;
      ldv<T, V>(g + r * C + cg * V, v0);
      if (two) ldv<T, V>(g + (r + R) * C + cg * V, v1);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += v0[k] + (two ? v1[k] : 0.f);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) shc[rl * C + cg * V + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) {
    float t = 0.f;
    for (int r = 0; r < R; ++r) t += shc[r * C + c];
    atomicAdd(&db[c], t);
  }
}

// C == 3 (the RGB bias gradients): a thread walks groups of 3 vectors = 3V elements, so element e of a
// group always belongs to channel e % 3 - register accumulation, one block reduction, 3 atomics per CTA.
template <typename T, int V>
__global__ void __launch_bounds__(NT) bias_grad_c3_kernel(const T* __restrict__ g, float* db, int64_t total) {
  __shared__ float sh[3][NT / 32];
  float acc[3] = {0.f, 0.f, 0.f};
  const int64_t ngroups = total / (3 * V);
  for (int64_t gi = (int64_t)blockIdx.x * NT + threadIdx.x; gi < ngroups; gi += (int64_t)gridDim.x * NT) {
    float v[3][V];
#pragma unroll
    for (int j = 0; j < 3; ++j) ldv<T, V>(g + gi * 3 * V + j * V, v[j]);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < V; ++k) acc[(j * V + k) % 3] += v[j][k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t e = ngroups * 3 * V; e < total; ++e) acc[e % 3] += to_f(g[e]);      // tail (< 3V elements)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = warp_sum(acc[c]);
    if ((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < NT / 32; ++w) s += sh[threadIdx.x][w];
    atomicAdd(&db[threadIdx.x], s);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) colsum_wide_kernel(const T* __restrict__ g, float* db, int64_t rows, int C,
                                                         int64_t rows_per) {
  const int c = blockIdx.x * NT + threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += to_f(g[r * C + c]);
  atomicAdd(&db[c], acc);
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
__global__ void bias_act_kernel(float* x, const float* bias, int rows, int cols, int act) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  float v = x[i] + (bias ? bias[i % cols] : 0.f);
  if (act == LG_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
  else if (act == LG_ACT_TANH) v = tanhf(v);
  x[i] = v;
}

// Keras binary_crossentropy (TF 1.15 eager branch) on probabilities + sigmoid backward.
__global__ void __launch_bounds__(NT) bce_kernel(const float* __restrict__ p, const float* __restrict__ target,
                                                 float target_const, int n, float weight, float* loss_accum,
                                                 float* dlogit) {
  __shared__ double sh[64];
  const float eps = 1e-7f, hi = 1.f - 1e-7f;
  float s = 0.f;
  const float scale = weight / (float)n;
  for (int i = threadIdx.x; i < n; i += NT) {
    float pr = p[i], t = target ? target[i] : target_const;
    float pc = fminf(fmaxf(pr, eps), hi);
    s += -(t * logf(pc + eps) + (1.f - t) * logf(1.f - pc + eps));
    if (dlogit) {
      float d = 0.f;
      if (pr >= eps && pr <= hi) d = -(t / (pc + eps) - (1.f - t) / (1.f - pc + eps));
      dlogit[i] = scale * d * pr * (1.f - pr);
    }
  }
  double a = s, b = 0.0;
  block_sum2(a, b, sh);
  if (threadIdx.x == 0 && loss_accum) atomicAdd(loss_accum, (float)(a * (double)scale));
}

template <typename T, int V>
__global__ void __launch_bounds__(NT) l1_tanh_bwd_kernel(const T* __restrict__ y, const T* __restrict__ t,
                                                         const T* __restrict__ g_in, T* __restrict__ dpre, int64_t n,
                                                         float scale, float* loss_accum) {
  __shared__ double sh[64];
  float s = 0.f;
  for (int64_t i = ((int64_t)blockIdx.x * NT + threadIdx.x) * V; i < n; i += (int64_t)gridDim.x * NT * V) {
    float yv[V], tv[V], gv[V], o[V];
    ldv<T, V>(y + i, yv);
    ldv<T, V>(t + i, tv);
    if (g_in) ldv<T, V>(g_in + i, gv);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float d = yv[k] - tv[k];
      s += fabsf(d);
      float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      o[k] = ((g_in ? gv[k] : 0.f) + scale * sg) * (1.f - yv[k] * yv[k]);
    }
    if (dpre) stv<T, V>(dpre + i, o);
  }
  double a = s, b = 0.0;
  block_sum2(a, b, sh);
  if (threadIdx.x == 0 && loss_accum) atomicAdd(loss_accum, (float)(a * (double)scale));
}

__global__ void adam_advance_kernel(double* st, double lr, double b1, double b2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = st[0] + 1.0;
    double p1 = (st[0] == 0.0 ? 1.0 : st[1]) * b1;
    double p2 = (st[0] == 0.0 ? 1.0 : st[2]) * b2;
    st[0] = t; st[1] = p1; st[2] = p2;
    st[3] = lr * sqrt(1.0 - p2) / (1.0 - p1);
  }
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_t, float b1, float b2,
                                         float eps, float clip) {
  if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p -= lr_t * m / (sqrtf(v) + eps);
}

// 28 bytes per parameter (read g, p, m, v; write p, m, v): 16-byte vectors, two vectors (8 loads) in flight per
// thread; VEC = false is the scalar form for ranges that are not 16-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(NT) adam_apply_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const double* __restrict__ state, float b1, float b2,
                                                        float eps, float clip) {
  const float lr_t = (float)state[3];
  if constexpr (VEC) {
    const int64_t nv = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * NT;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < nv; i += 2 * stride) {
      const int64_t j = i + stride;
      const bool two = j < nv;
      float4 gi = g4[i], pi = p4[i], mi = m4[i], vi = v4[i];
      float4 gj, pj, mj, vj;
      if (two) { gj = g4[j]; pj = p4[j]; mj = m4[j]; vj = v4[j]; }
      adam_one(pi.x, gi.x, mi.x, vi.x, lr_t, b1, b2, eps, clip);
      adam_one(pi.y, gi.y, mi.y, vi.y, lr_t, b1, b2, eps, clip);
      adam_one(pi.z, gi.z, mi.z, vi.z, lr_t, b1, b2, eps, clip);
      adam_one(pi.w, gi.w, mi.w, vi.w, lr_t, b1, b2, eps, clip);
      p4[i] = pi; m4[i] = mi; v4[i] = vi;
      if (two) {
        adam_one(pj.x, gj.x, mj.x, vj.x, lr_t, b1, b2, eps, clip);
        adam_one(pj.y, gj.y, mj.y, vj.y, lr_t, b1, b2, eps, clip);
        adam_one(pj.z, gj.z, mj.z, vj.z, lr_t, b1, b2, eps, clip);
        adam_one(pj.w, gj.w, mj.w, vj.w, lr_t, b1, b2, eps, clip);
        p4[j] = pj; m4[j] = mj; v4[j] = vj;
      }
    }
    // tail (< 4 elements)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
      const int64_t e = (nv << 2) + threadIdx.x;
      adam_one(p[e], g[e], m[e], v[e], lr_t, b1, b2, eps, clip);
    }
  } else {
    const int64_t stride = (int64_t)gridDim.x * NT;
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += stride)
      adam_one(p[i], g[i], m[i], v[i], lr_t, b1, b2, eps, clip);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = from_f<D>(to_f(s[i]));
}

// d = s * scale + shift (fp32 arithmetic): copies, casts and the adjuster's (cond + 1) / 2 (eager_trainer.py:155)
template <typename S, typename D>
__global__ void scale_shift_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n, float scale, float shift) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    d[i] = from_f<D>(fmaf(to_f(s[i]), scale, shift));
}

// data_rescale of utils.py:51-52 on decoded image bytes: y = x / 127.5 - 1 (fp32 division, then the subtraction,
// each rounded once as TF's two ops are), 16 pixels-bytes per thread
template <typename D>
__global__ void u8_rescale_kernel(const uint8_t* __restrict__ s, D* __restrict__ d, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nv = n >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(s) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float o[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) o[b] = __fsub_rn(__fdiv_rn((float)((w[q] >> (8 * b)) & 0xffu), 127.5f), 1.0f);
      stv<D, 4>(d + i * 16 + q * 4, o);
    }
  }
  for (int64_t i = (nv << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    d[i] = from_f<D>(__fsub_rn(__fdiv_rn((float)s[i], 127.5f), 1.0f));
}

// dst[row][0..Cp) = src[row][0..C) followed by zeros (3-channel images -> 16-channel TMA-able rows)
template <typename T>
__global__ void pad_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int C, int Cp) {
  const int64_t total = rows * Cp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / Cp;
    const int c = (int)(i - r * Cp);
    dst[i] = c < C ? src[r * C + c] : from_f<T>(0.f);
  }
}

// RGB bf16 image -> 8-channel pixels (16 B, what the row-streaming conv kernel fetches by TMA): one pixel per
// thread, three 2-byte loads (coalesced across the warp) and one 16-byte store.
__global__ void __launch_bounds__(256) pad3to8_kernel(const uint16_t* __restrict__ src, uint4* __restrict__ dst,
                                                      int64_t rows) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const uint16_t* s = src + 3 * r;
    const uint32_t c0 = __ldg(s), c1 = __ldg(s + 1), c2 = __ldg(s + 2);
    dst[r] = make_uint4(c0 | (c1 << 16), c2, 0u, 0u);
  }
}

// W[25][A][B] fp32 -> Wt[25][Ap][Bp] bf16 (b contiguous), Wf[25][Bp][Ap] bf16 (a contiguous).
// One 64x64 (a, b) tile of one tap per CTA: W is read and Wt written along b, the tile is transposed in
// shared memory and Wf written along a, so all three streams are coalesced (the weights are re-packed
// after every optimiser step, ~20 MB of traffic per step).
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W, bf16* __restrict__ Wt,
                                                           bf16* __restrict__ Wf, int A, int B, int Ap, int Bp) {
  __shared__ float tile[64][65];
  const int tap = blockIdx.z, a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* Wp = W + (int64_t)tap * A * B;
  for (int r = ty; r < 64; r += 8) {
    const int a = a0 + r, b = b0 + 2 * tx;
    const float v0 = (a < A && b < B) ? Wp[(int64_t)a * B + b] : 0.f;
    const float v1 = (a < A && b + 1 < B) ? Wp[(int64_t)a * B + b + 1] : 0.f;
    tile[r][2 * tx] = v0; tile[r][2 * tx + 1] = v1;
    if (a < Ap && b < Bp)
      *reinterpret_cast<__nv_bfloat162*>(Wt + ((int64_t)tap * Ap + a) * Bp + b) = __floats2bfloat162_rn(v0, v1);
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 8) {
    const int b = b0 + r, a = a0 + 2 * tx;
    if (a < Ap && b < Bp)
      *reinterpret_cast<__nv_bfloat162*>(Wf + ((int64_t)tap * Bp + b) * Ap + a) =
          __floats2bfloat162_rn(tile[2 * tx][r], tile[2 * tx + 1][r]);
  }
}

// The same for up to LG_PACK_MAX layers in ONE launch (the nine conv layers are re-packed every step: nine ~5 us
// launches otherwise); a CTA finds its layer by its position in the table of CTA prefix sums.
struct PackTable {
  const float* W[LG_PACK_MAX];
  bf16* Wt[LG_PACK_MAX];
  bf16* Wf[LG_PACK_MAX];
  int A[LG_PACK_MAX], B[LG_PACK_MAX], Ap[LG_PACK_MAX], Bp[LG_PACK_MAX];
  int cta_end[LG_PACK_MAX];          // exclusive prefix sums of (b tiles x a tiles x 25)
  int n;
};

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackTable t) {
  __shared__ float tile[64][65];
  int l = 0;
  while (l < t.n - 1 && (int)blockIdx.x >= t.cta_end[l]) ++l;
  const int local = blockIdx.x - (l ? t.cta_end[l - 1] : 0);
  const int A = t.A[l], B = t.B[l], Ap = t.Ap[l], Bp = t.Bp[l];
  const int bt = (Bp + 63) / 64, at = (Ap + 63) / 64;
  const int tap = local / (bt * at), rem = local - tap * (bt * at);
  const int a0 = (rem / bt) * 64, b0 = (rem % bt) * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* Wp = t.W[l] + (int64_t)tap * A * B;
  bf16* Wt = t.Wt[l];
  bf16* Wf = t.Wf[l];
  for (int r = ty; r < 64; r += 8) {
    const int a = a0 + r, b = b0 + 2 * tx;
    const float v0 = (a < A && b < B) ? Wp[(int64_t)a * B + b] : 0.f;
    const float v1 = (a < A && b + 1 < B) ? Wp[(int64_t)a * B + b + 1] : 0.f;
    tile[r][2 * tx] = v0; tile[r][2 * tx + 1] = v1;
    if (a < Ap && b < Bp)
      *reinterpret_cast<__nv_bfloat162*>(Wt + ((int64_t)tap * Ap + a) * Bp + b) = __floats2bfloat162_rn(v0, v1);
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 8) {
    const int b = b0 + r, a = a0 + 2 * tx;
    if (a < Ap && b < Bp)
      *reinterpret_cast<__nv_bfloat162*>(Wf + ((int64_t)tap * Bp + b) * Ap + a) =
          __floats2bfloat162_rn(tile[2 * tx][r], tile[2 * tx + 1][r]);
  }
}

// Several Keras-BCE terms in one launch (one CTA per term): the seven loss terms of a train step are 1-CTA kernels
// whose cost is launch latency.
struct BceTable {
  lg_bce_item_t it[LG_BCE_MAX];
  int first_cta[LG_BCE_MAX + 1];      // term i owns CTAs [first_cta[i], first_cta[i+1]): BCE_PER elements each
};
constexpr int BCE_PER = NT * 2;

__global__ void __launch_bounds__(NT) bce_multi_kernel(const BceTable t, int nterms) {
  __shared__ double sh[64];
  int term = 0;
  while (term + 1 < nterms && (int)blockIdx.x >= t.first_cta[term + 1]) ++term;
  const lg_bce_item_t& q = t.it[term];
  const int beg = ((int)blockIdx.x - t.first_cta[term]) * BCE_PER;
  const int end = min(q.n, beg + BCE_PER);
  const float eps = 1e-7f, hi = 1.f - 1e-7f;
  float s = 0.f;
  const float scale = q.weight / (float)q.n;
  for (int i = beg + threadIdx.x; i < end; i += NT) {
    float pr = q.p[i], tg = q.target ? q.target[i] : q.target_const;
    float pc = fminf(fmaxf(pr, eps), hi);
    s += -(tg * logf(pc + eps) + (1.f - tg) * logf(1.f - pc + eps));
    if (q.dlogit) {
      float d = 0.f;
      if (pr >= eps && pr <= hi) d = -(tg / (pc + eps) - (1.f - tg) / (1.f - pc + eps));
      q.dlogit[i] = scale * d * pr * (1.f - pr);
    }
  }
  double a = s, b = 0.0;
  block_sum2(a, b, sh);
  if (threadIdx.x == 0 && q.loss_accum) atomicAdd(q.loss_accum, (float)(a * (double)scale));
}

inline void chunking(int64_t M, int V, int N, int64_t* per_cta, int* chunks, int iters = 4) {
  // `iters` vector iterations per thread, rounded so that chunk boundaries stay vector aligned.  Measured on the
  // largest map of the step ([128,128,128,32] bf16, alone, L2 flushed): the forward pass is fastest at 4 (5.5 TB/s;
  // 8: 5.4, 16: 5.3), the two backward kernels - whose CTAs pay a longer per-sample set-up (statistics, reductions,
  // shared-memory bias sums) - at 8 (reduce + apply 4.52 -> 5.03 TB/s, apply alone 4.65 -> 5.21 TB/s).
  int64_t unit = (int64_t)NT * V;
  int64_t per = unit * iters;
  int64_t c = (M + per - 1) / per;
  if (c < 1) c = 1;
  if (c > 65535) { c = 65535; per = ((M + c - 1) / c + unit - 1) / unit * unit; c = (M + per - 1) / per; }
  *per_cta = per; *chunks = (int)c;
  (void)N;
}

}  // namespace

#define DISPATCH_TV(dtype, M, CALL)                                       \
  do {                                                                    \
    if ((dtype) == LG_BF16) {                                             \
      if ((M) % 8 == 0) { CALL(bf16, 8); } else { CALL(bf16, 1); }        \
    } else {                                                              \
      if ((M) % 4 == 0) { CALL(float, 4); } else { CALL(float, 1); }      \
    }                                                                     \
  } while (0)

extern "C" int lg_rowstats(const void* z, double* stats, int N, int64_t M, float alpha_pre, int dtype,
                           void* stream) {
  LG_REQUIRE(z && stats && N > 0 && M > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(T, V)                                                                              \
  {                                                                                             \
    int64_t per; int ch; chunking(M, V, N, &per, &ch);                                          \
    rowstats_kernel<T, V><<<dim3(ch, N), NT, 0, st>>>((const T*)z, stats, M, per, alpha_pre);   \
  }
  DISPATCH_TV(dtype, M, CALL);
#undef CALL
  LG_LAUNCH_CHECK();
  return LG_OK;
}

// (z_dtype, dtype) -> <TZ, T, V>; a bf16 z with an fp32 out is not a combination the path uses.
#define DISPATCH_ZT(zd, od, M, CALL)                                                 \
  do {                                                                               \
    if ((zd) == LG_BF16 && (od) == LG_BF16) {                                        \
      if ((M) % 8 == 0) { CALL(bf16, bf16, 8); } else { CALL(bf16, bf16, 1); }       \
    } else if ((zd) == LG_F32 && (od) == LG_BF16) {                                  \
      if ((M) % 4 == 0) { CALL(float, bf16, 4); } else { CALL(float, bf16, 1); }     \
    } else {                                                                         \
      if ((M) % 4 == 0) { CALL(float, float, 4); } else { CALL(float, float, 1); }   \
    }                                                                                \
  } while (0)

extern "C" int lg_instnorm_act_fwd(const void* z, const double* stats, const float* gamma, const float* beta,
                                   const void* skip, void* out, int N, int64_t M, float eps, float alpha_pre,
                                   float alpha_post, int z_dtype, int dtype, void* stream) {
  LG_REQUIRE(z && stats && gamma && beta && out && N > 0 && M > 0, "bad arguments");
  LG_REQUIRE(!(z_dtype == LG_BF16 && dtype == LG_F32), "bf16 z with fp32 out is not supported");
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(TZ, T, V)                                                                                         \
  {                                                                                                            \
    int64_t per; int ch; chunking(M, V, N, &per, &ch);                                                         \
    instnorm_fwd_kernel<TZ, T, V><<<dim3(ch, N), NT, 0, st>>>((const TZ*)z, stats, gamma, beta, (const T*)skip, \
                                                              (T*)out, M, per, eps, alpha_pre, alpha_post);    \
  }
  DISPATCH_ZT(z_dtype, dtype, M, CALL);
#undef CALL
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_instnorm_act_bwd_reduce(const void* g, const void* z, const double* stats, const float* gamma,
                                          const float* beta, double* red, int N, int64_t M, float eps,
                                          float alpha_pre, float alpha_post, int z_dtype, int dtype, void* stream) {
  LG_REQUIRE(g && z && stats && gamma && beta && red && N > 0 && M > 0, "bad arguments");
  LG_REQUIRE(!(z_dtype == LG_BF16 && dtype == LG_F32), "bf16 z with fp32 g is not supported");
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(TZ, T, V)                                                                                       \
  {                                                                                                          \
    int64_t per; int ch; chunking(M, V, N, &per, &ch, 8);                                                    \
    instnorm_bwd_reduce_kernel<TZ, T, V><<<dim3(ch, N), NT, 0, st>>>((const T*)g, (const TZ*)z, stats, gamma, \
                                                                     beta, red, M, per, eps, alpha_pre,      \
                                                                     alpha_post);                            \
  }
  DISPATCH_ZT(z_dtype, dtype, M, CALL);
#undef CALL
  LG_LAUNCH_CHECK();
  return LG_OK;
}

static int apply_vec(int64_t M, int z_dtype) {
  if (z_dtype == LG_BF16) return M % 8 == 0 ? 8 : 1;
  return M % 4 == 0 ? 4 : 1;
}

extern "C" int lg_instnorm_bias_grad_fusable(int64_t M, int C, int z_dtype) {
  const int V = apply_vec(M, z_dtype);
  return (V > 1 && C > 0 && C <= NT * V && (NT * V) % C == 0 && C % V == 0 && M % C == 0) ? 1 : 0;
}

extern "C" int lg_instnorm_act_bwd_apply(const void* g, const void* z, const double* stats, const double* red,
                                         const float* gamma, const float* beta, void* dz, float* dgamma,
                                         float* dbeta, float* dbias, int C, int dy_ready, int N, int64_t M, float eps,
                                         float alpha_pre, float alpha_post, int z_dtype, int dtype, void* stream) {
  LG_REQUIRE(g && z && stats && red && gamma && beta && dz && N > 0 && M > 0, "bad arguments");
  LG_REQUIRE(!(z_dtype == LG_BF16 && dtype == LG_F32), "bf16 z with fp32 g is not supported");
  if (dbias != nullptr && !lg_instnorm_bias_grad_fusable(M, C, z_dtype)) {
    lg_set_error("lg_instnorm_act_bwd_apply: bias gradient not fusable for M=%lld C=%d", (long long)M, C);
    return LG_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(TZ, T, V)                                                                                           \
  {                                                                                                              \
    int64_t per; int ch; chunking(M, V, N, &per, &ch, 8);                                                        \
    if (dbias != nullptr)                                                                                        \
      instnorm_bwd_apply_kernel<TZ, T, V, (V > 1)><<<dim3(ch, N), NT, 0, st>>>(                                  \
          (const T*)g, (const TZ*)z, stats, red, gamma, beta, (TZ*)dz, dgamma, dbeta, dbias, C, dy_ready, N, M,  \
          per, eps, alpha_pre, alpha_post);                                                                      \
    else                                                                                                         \
      instnorm_bwd_apply_kernel<TZ, T, V, false><<<dim3(ch, N), NT, 0, st>>>(                                    \
          (const T*)g, (const TZ*)z, stats, red, gamma, beta, (TZ*)dz, dgamma, dbeta, nullptr, C, dy_ready, N,   \
          M, per, eps, alpha_pre, alpha_post);                                                                   \
  }
  DISPATCH_ZT(z_dtype, dtype, M, CALL);
#undef CALL
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_bias_grad(const void* g, float* db, int64_t rows, int C, int dtype, void* stream) {
  LG_REQUIRE(g && db && rows > 0 && C > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = rows * C;
  if (C == 3 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int V = dtype == LG_BF16 ? 8 : 4;
    int64_t need = (total / (3 * V) + NT - 1) / NT;
    int64_t cap = (int64_t)lg_num_sms() * 8;
    int gsz = (int)(need < 1 ? 1 : (need < cap ? need : cap));
    if (dtype == LG_BF16) bias_grad_c3_kernel<bf16, 8><<<gsz, NT, 0, st>>>((const bf16*)g, db, total);
    else bias_grad_c3_kernel<float, 4><<<gsz, NT, 0, st>>>((const float*)g, db, total);
    LG_LAUNCH_CHECK();
    return LG_OK;
  }
  if (C > 4096) {   // wide, short matrices (the dense heads): one thread per column, coalesced over columns
    int64_t rows_per = (rows + 7) / 8;
    dim3 grid((C + NT - 1) / NT, (unsigned)((rows + rows_per - 1) / rows_per));
    if (dtype == LG_BF16) colsum_wide_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)g, db, rows, C, rows_per);
    else colsum_wide_kernel<float><<<grid, NT, 0, st>>>((const float*)g, db, rows, C, rows_per);
    LG_LAUNCH_CHECK();
    return LG_OK;
  }
  const int V = dtype == LG_BF16 ? 8 : 4;
  const bool periodic = (C % V == 0) && ((NT * V) % C == 0);
  if (!periodic && C % V == 0 && C / V <= NT && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int R = NT / (C / V);
    int64_t ctas = (int64_t)lg_num_sms() * 4;
    int64_t rows_per = (rows + ctas - 1) / ctas;
    rows_per = (rows_per + 2 * R - 1) / (2 * R) * (2 * R);
    ctas = (rows + rows_per - 1) / rows_per;
    const size_t shm = (size_t)R * C * sizeof(float);
    if (dtype == LG_BF16) bias_grad_group_kernel<bf16, 8><<<(int)ctas, NT, shm, st>>>((const bf16*)g, db, rows, C, rows_per);
    else bias_grad_group_kernel<float, 4><<<(int)ctas, NT, shm, st>>>((const float*)g, db, rows, C, rows_per);
    LG_LAUNCH_CHECK();
    return LG_OK;
  }
  int64_t unit = (int64_t)NT * V;
  int64_t per = unit * 8;
  int64_t ctas = (total + per - 1) / per;
  int64_t cap = (int64_t)lg_num_sms() * 8;
  if (ctas > cap) { per = ((total + cap - 1) / cap + unit - 1) / unit * unit; ctas = (total + per - 1) / per; }
  // per must keep chunk starts on row-pattern boundaries for the periodic path: unit % C == 0 holds.
  size_t shm = (size_t)C * sizeof(float);
  if (dtype == LG_BF16) {
    if (periodic) bias_grad_kernel<bf16, 8, true><<<(int)ctas, NT, shm, st>>>((const bf16*)g, db, total, C, per);
    else bias_grad_kernel<bf16, 1, false><<<(int)ctas, NT, shm, st>>>((const bf16*)g, db, total, C, per);
  } else {
    if (periodic) bias_grad_kernel<float, 4, true><<<(int)ctas, NT, shm, st>>>((const float*)g, db, total, C, per);
    else bias_grad_kernel<float, 1, false><<<(int)ctas, NT, shm, st>>>((const float*)g, db, total, C, per);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_bias_act(float* x, const float* bias, int rows, int cols, int act, void* stream) {
  LG_REQUIRE(x && rows > 0 && cols > 0, "bad arguments");
  int n = rows * cols;
  bias_act_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, bias, rows, cols, act);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_bce_sigmoid(const float* p, const float* target, float target_const, int rows, int cols,
                              float weight, float* loss_accum, float* dlogit, void* stream) {
  LG_REQUIRE(p && rows > 0 && cols > 0, "bad arguments");
  bce_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(p, target, target_const, rows * cols, weight, loss_accum, dlogit);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_bce_sigmoid_multi(const lg_bce_item_t* items, int n, void* stream) {
  LG_REQUIRE(items && n > 0 && n <= LG_BCE_MAX, "bad arguments");
  BceTable t;
  int ctas = 0;
  for (int i = 0; i < n; ++i) {
    LG_REQUIRE(items[i].p && items[i].n > 0, "bad item");
    t.it[i] = items[i];
    t.first_cta[i] = ctas;
    ctas += (items[i].n + BCE_PER - 1) / BCE_PER;
  }
  t.first_cta[n] = ctas;
  bce_multi_kernel<<<ctas, NT, 0, (cudaStream_t)stream>>>(t, n);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_l1_tanh_bwd(const void* y, const void* t, const void* g_in, void* dpre, int64_t n, float weight,
                              float* loss_accum, int dtype, void* stream) {
  LG_REQUIRE(y && t && n > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = weight / (float)n;
  int ctas = lg_num_sms() * 4;
#define CALL(T, V)                                                                                         \
  {                                                                                                        \
    int64_t need = (n + (int64_t)NT * V - 1) / ((int64_t)NT * V);                                          \
    int gsz = (int)(need < ctas ? need : ctas);                                                            \
    l1_tanh_bwd_kernel<T, V><<<gsz, NT, 0, st>>>((const T*)y, (const T*)t, (const T*)g_in, (T*)dpre, n,    \
                                                 scale, loss_accum);                                       \
  }
  DISPATCH_TV(dtype, n, CALL);
#undef CALL
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_adam_advance(double* state, double lr, double beta1, double beta2, void* stream) {
  LG_REQUIRE(state, "bad arguments");
  adam_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, lr, beta1, beta2);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_adam_apply(float* p, const float* g, float* m, float* v, int64_t n, const double* state,
                             float beta1, float beta2, float eps, float clip, void* stream) {
  LG_REQUIRE(p && g && m && v && state && n > 0, "bad arguments");
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && n >= 4;
  int64_t need = vec ? ((n >> 2) + 2 * NT - 1) / (2 * NT) : (n + NT - 1) / NT;
  int64_t cap = (int64_t)lg_num_sms() * 8;
  int gsz = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  if (vec) adam_apply_kernel<true><<<gsz, NT, 0, (cudaStream_t)stream>>>(p, g, m, v, n, state, beta1, beta2, eps, clip);
  else adam_apply_kernel<false><<<gsz, NT, 0, (cudaStream_t)stream>>>(p, g, m, v, n, state, beta1, beta2, eps, clip);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  LG_REQUIRE(src && dst && n > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t need = (n + 255) / 256;
  int64_t cap = (int64_t)lg_num_sms() * 16;
  int gsz = (int)(need < cap ? need : cap);
  if (src_dtype == LG_F32 && dst_dtype == LG_BF16) cast_kernel<float, bf16><<<gsz, 256, 0, st>>>((const float*)src, (bf16*)dst, n);
  else if (src_dtype == LG_BF16 && dst_dtype == LG_F32) cast_kernel<bf16, float><<<gsz, 256, 0, st>>>((const bf16*)src, (float*)dst, n);
  else if (src_dtype == LG_F32 && dst_dtype == LG_F32) cast_kernel<float, float><<<gsz, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else cast_kernel<bf16, bf16><<<gsz, 256, 0, st>>>((const bf16*)src, (bf16*)dst, n);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_scale_shift(const void* src, void* dst, int64_t n, float scale, float shift, int src_dtype,
                              int dst_dtype, void* stream) {
  LG_REQUIRE(src && dst && n > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t need = (n + 255) / 256;
  int64_t cap = (int64_t)lg_num_sms() * 16;
  int gsz = (int)(need < cap ? need : cap);
  if (src_dtype == LG_F32 && dst_dtype == LG_BF16) scale_shift_kernel<float, bf16><<<gsz, 256, 0, st>>>((const float*)src, (bf16*)dst, n, scale, shift);
  else if (src_dtype == LG_BF16 && dst_dtype == LG_F32) scale_shift_kernel<bf16, float><<<gsz, 256, 0, st>>>((const bf16*)src, (float*)dst, n, scale, shift);
  else if (src_dtype == LG_F32 && dst_dtype == LG_F32) scale_shift_kernel<float, float><<<gsz, 256, 0, st>>>((const float*)src, (float*)dst, n, scale, shift);
  else scale_shift_kernel<bf16, bf16><<<gsz, 256, 0, st>>>((const bf16*)src, (bf16*)dst, n, scale, shift);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_u8_rescale(const void* src, void* dst, int64_t n, int dst_dtype, void* stream) {
  LG_REQUIRE(src && dst && n > 0, "bad arguments");
  LG_REQUIRE(dst_dtype == LG_F32 || dst_dtype == LG_BF16, "dst_dtype must be LG_F32 or LG_BF16");
  LG_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "src and dst must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t need = ((n + 15) / 16 + 255) / 256;
  int64_t cap = (int64_t)lg_num_sms() * 8;
  int gsz = (int)(need < cap ? need : cap);
  if (dst_dtype == LG_F32) u8_rescale_kernel<float><<<gsz, 256, 0, st>>>((const uint8_t*)src, (float*)dst, n);
  else u8_rescale_kernel<bf16><<<gsz, 256, 0, st>>>((const uint8_t*)src, (bf16*)dst, n);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_pad_channels(const void* src, void* dst, int64_t rows, int C, int Cpad, int dtype, void* stream) {
  LG_REQUIRE(src && dst && rows > 0 && C > 0 && Cpad >= C, "bad arguments");
  const int64_t total = rows * Cpad;
  int64_t need = (total + 255) / 256, cap = (int64_t)lg_num_sms() * 16;
  int gsz = (int)(need < cap ? need : cap);
  if (dtype == LG_BF16 && C == 3 && Cpad == 8) {
    need = (rows + 255) / 256;
    gsz = (int)(need < cap ? need : cap);
    pad3to8_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, (uint4*)dst, rows);
  } else if (dtype == LG_BF16)
    pad_channels_kernel<bf16><<<gsz, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, rows, C, Cpad);
  else
    pad_channels_kernel<float><<<gsz, 256, 0, (cudaStream_t)stream>>>((const float*)src, (float*)dst, rows, C, Cpad);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_pack_conv_weights_multi(const float* const* W, void* const* wpack, const int* A, const int* B, int n,
                                          void* stream) {
  LG_REQUIRE(W && wpack && A && B && n > 0 && n <= LG_PACK_MAX, "bad arguments");
  PackTable t;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    LG_REQUIRE(W[i] && wpack[i] && A[i] > 0 && B[i] > 0, "bad layer");
    const int Ap = (A[i] + 15) / 16 * 16, Bp = (B[i] + 15) / 16 * 16;
    t.W[i] = W[i]; t.Wt[i] = (bf16*)wpack[i]; t.Wf[i] = (bf16*)wpack[i] + (int64_t)25 * Ap * Bp;
    t.A[i] = A[i]; t.B[i] = B[i]; t.Ap[i] = Ap; t.Bp[i] = Bp;
    total += ((Bp + 63) / 64) * ((Ap + 63) / 64) * 25;
    t.cta_end[i] = total;
  }
  t.n = n;
  pack_weights_multi_kernel<<<total, 256, 0, (cudaStream_t)stream>>>(t);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int64_t lg_pack_conv_weights(const float* W, void* wpack, int A, int B, void* stream) {
  if (A <= 0 || B <= 0) { lg_set_error("lg_pack_conv_weights: bad arguments"); return LG_ERR_INVALID; }
  const int Ap = (A + 15) / 16 * 16, Bp = (B + 15) / 16 * 16;
  const int64_t half = (int64_t)25 * Ap * Bp;
  if (wpack == nullptr) return 2 * half * (int64_t)sizeof(bf16);
  if (W == nullptr) { lg_set_error("lg_pack_conv_weights: W is NULL"); return LG_ERR_INVALID; }
  bf16* wt = (bf16*)wpack;
  bf16* wf = wt + half;
  pack_weights_kernel<<<dim3((Bp + 63) / 64, (Ap + 63) / 64, 25), 256, 0, (cudaStream_t)stream>>>(W, wt, wf, A, B, Ap,
                                                                                                  Bp);
  LG_LAUNCH_CHECK();
  return 2 * half * (int64_t)sizeof(bf16);
}
