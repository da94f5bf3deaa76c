// fp32-accumulate SIMT implicit-GEMM kernels: the exact-precision ("fp32 mode") path of the
// convolution family and the dense layers, and the path for the 3-channel edge layers.
//
// One 64x64x16 register-tiled skeleton; the three convolution ops and the dense GEMM differ only
// in how a GEMM coordinate maps to memory (the Prob functors below).  See include/littlegan_b200.h
// for the index relations (big/small map, W[5,5,A,B]).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

// ---------------------------------------------------------------------------------------------
// skeleton
// ---------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(NT) simt_gemm_kernel(P p) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  p.set_slice(blockIdx.z);
  const int kbeg = p.k_begin(), kend = p.k_end();

  // load coordinates of this thread (4 A elements, 4 B elements per k-chunk)
  typename P::MCtx mctx[4];
  typename P::NCtx nctx[4];
  int a_m[4], a_k[4], b_n[4], b_k[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (P::A_KFAST) { a_m[r] = (tid >> 4) + 16 * r; a_k[r] = tid & 15; }
    else            { a_m[r] = tid & 63;            a_k[r] = (tid >> 6) + 4 * r; }
    if (P::B_KFAST) { b_n[r] = (tid >> 4) + 16 * r; b_k[r] = tid & 15; }
    else            { b_n[r] = tid & 63;            b_k[r] = (tid >> 6) + 4 * r; }
    mctx[r] = p.prep_m(m0 + a_m[r]);
    nctx[r] = p.prep_n(n0 + b_n[r]);
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    float av[4], bv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int k = k0 + a_k[r];
      av[r] = (k < kend) ? p.load_a(mctx[r], p.prep_k(k)) : 0.f;
      k = k0 + b_k[r];
      bv[r] = (k < kend) ? p.load_b(p.prep_k(k), nctx[r]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      As[a_k[r]][a_m[r]] = av[r];
      Bs[b_k[r]][b_n[r]] = bv[r];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  // epilogue: store + optional per-sample statistics of the stored (pre-activation) values
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    float s1 = 0.f, s2 = 0.f;
    if (m < p.M) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < p.N) {
          float v = p.store(m, n, acc[i][j]);
          s1 += v;
          s2 += v * v;
        }
      }
    }
    if (P::HAS_STATS && p.stats != nullptr) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (tx == 0 && m < p.M) {
        int smp = p.sample_of(m);
        atomicAdd(&p.stats[2 * smp], (double)s1);
        atomicAdd(&p.stats[2 * smp + 1], (double)s2);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// conv fprop: small[m=(n,i,j)][b] = sum_k=(ky,kx,a) big[n, s*i+ky-p, s*j+kx-p, a] * W[k][b]
// ---------------------------------------------------------------------------------------------
template <typename T>
struct FpropProb {
  static constexpr bool A_KFAST = true, B_KFAST = false, HAS_STATS = true;
  const T* big; const float* W; const float* bias; T* out; double* stats;
  int M, N, K;                 // M = Nimg*Hs*Ws, N = B, K = 25*A
  int Hb, Wb, A, Hs, Ws, s, pad;
  struct MCtx { int n, y0, x0; bool ok; };
  struct NCtx { int n; };
  struct KCtx { int ky, kx, a, k; };
  __device__ void set_slice(int) {}
  __device__ int k_begin() const { return 0; }
  __device__ int k_end() const { return K; }
  __device__ MCtx prep_m(int m) const {
    MCtx c; c.ok = m < M;
    int j = m % Ws, t = m / Ws, i = t % Hs; c.n = t / Hs;
    c.y0 = s * i - pad; c.x0 = s * j - pad; return c;
  }
  __device__ NCtx prep_n(int n) const { return NCtx{n}; }
  __device__ KCtx prep_k(int k) const {
    KCtx c; c.k = k; int tap = k / A; c.a = k - tap * A; c.ky = tap / 5; c.kx = tap - 5 * c.ky; return c;
  }
  __device__ float load_a(const MCtx& m, const KCtx& k) const {
    int y = m.y0 + k.ky, x = m.x0 + k.kx;
    if (!m.ok || y < 0 || y >= Hb || x < 0 || x >= Wb) return 0.f;
    return to_f(big[(((int64_t)m.n * Hb + y) * Wb + x) * A + k.a]);
  }
  __device__ float load_b(const KCtx& k, const NCtx& n) const {
    return n.n < N ? W[(int64_t)k.k * N + n.n] : 0.f;
  }
  __device__ float store(int m, int n, float v) const {
    if (bias) v += bias[n];
    out[(int64_t)m * N + n] = from_f<T>(v);
    return v;
  }
  __device__ int sample_of(int m) const { return m / (Hs * Ws); }
};

// ---------------------------------------------------------------------------------------------
// conv dgrad / transposed-conv forward, one launch z-slice per output phase (py,px):
//   big[n, s*i+py, s*j+px, a] = sum_{taps of the phase, b} small[n, i+di, j+dj, b] * W[tap][a][b]
// with ky in the phase iff (py + pad - ky) % s == 0 and di = (py + pad - ky) / s.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DgradProb {
  static constexpr bool A_KFAST = true, B_KFAST = true, HAS_STATS = true;
  const T* small; const float* W; const float* bias; T* out; double* stats;
  int M, N, K;                 // M = Nimg*Hs*Ws, N = A, K = ntaps(phase)*B
  int Hb, Wb, A, B, Hs, Ws, s, pad, act;
  int py, px, nky, nkx;        // phase state
  int kys[5], dis[5], kxs[5], djs[5];
  struct MCtx { int n, i, j; bool ok; };
  struct NCtx { int n; };
  struct KCtx { int tap, di, dj, b; };
  __device__ void set_slice(int z) {
    py = z / s; px = z - py * s; nky = 0; nkx = 0;
    for (int k = 0; k < 5; ++k) {
      int d = py + pad - k;
      if (d % s == 0) { kys[nky] = k; dis[nky] = d / s; ++nky; }
      d = px + pad - k;
      if (d % s == 0) { kxs[nkx] = k; djs[nkx] = d / s; ++nkx; }
    }
    K = nky * nkx * B;
  }
  __device__ int k_begin() const { return 0; }
  __device__ int k_end() const { return K; }
  __device__ MCtx prep_m(int m) const {
    MCtx c; c.ok = m < M; c.j = m % Ws; int t = m / Ws; c.i = t % Hs; c.n = t / Hs; return c;
  }
  __device__ NCtx prep_n(int n) const { return NCtx{n}; }
  __device__ KCtx prep_k(int k) const {
    KCtx c; int t = k / B; c.b = k - t * B; int iy = t / nkx, ix = t - iy * nkx;
    c.tap = kys[iy] * 5 + kxs[ix]; c.di = dis[iy]; c.dj = djs[ix]; return c;
  }
  __device__ float load_a(const MCtx& m, const KCtx& k) const {
    int i = m.i + k.di, j = m.j + k.dj;
    if (!m.ok || i < 0 || i >= Hs || j < 0 || j >= Ws) return 0.f;
    return to_f(small[(((int64_t)m.n * Hs + i) * Ws + j) * B + k.b]);
  }
  __device__ float load_b(const KCtx& k, const NCtx& n) const {
    return n.n < N ? W[((int64_t)k.tap * A + n.n) * B + k.b] : 0.f;
  }
  __device__ float store(int m, int n, float v) const {
    int j = m % Ws, t = m / Ws, i = t % Hs, img = t / Hs;
    if (bias) v += bias[n];
    float o = act == LG_ACT_TANH ? tanhf(v) : v;
    out[(((int64_t)img * Hb + (s * i + py)) * Wb + (s * j + px)) * A + n] = from_f<T>(o);
    return v;
  }
  __device__ int sample_of(int m) const { return m / (Hs * Ws); }
};

// ---------------------------------------------------------------------------------------------
// conv wgrad (split-K over grid.z, fp32 atomics):
//   dW[mm=(tap,a)][b] += sum_k=(n,i,j) big[n, s*i+ky-p, s*j+kx-p, a] * small[k][b]
// ---------------------------------------------------------------------------------------------
template <typename T>
struct WgradProb {
  static constexpr bool A_KFAST = false, B_KFAST = false, HAS_STATS = false;
  const T* big; const T* small; float* dW; double* stats;
  int M, N, K;                 // M = 25*A, N = B, K = Nimg*Hs*Ws
  int Hb, Wb, A, Hs, Ws, s, pad, kper;
  int kb, ke;
  struct MCtx { int dy, dx, a; bool ok; };
  struct NCtx { int n; };
  struct KCtx { int n, i, j, k; };
  __device__ void set_slice(int z) { kb = z * kper; ke = min(K, kb + kper); }
  __device__ int k_begin() const { return kb; }
  __device__ int k_end() const { return ke; }
  __device__ MCtx prep_m(int m) const {
    MCtx c; c.ok = m < M; int tap = m / A; c.a = m - tap * A; int ky = tap / 5;
    c.dy = ky - pad; c.dx = (tap - 5 * ky) - pad; return c;
  }
  __device__ NCtx prep_n(int n) const { return NCtx{n}; }
  __device__ KCtx prep_k(int k) const {
    KCtx c; c.k = k; c.j = k % Ws; int t = k / Ws; c.i = t % Hs; c.n = t / Hs; return c;
  }
  __device__ float load_a(const MCtx& m, const KCtx& k) const {
    int y = s * k.i + m.dy, x = s * k.j + m.dx;
    if (!m.ok || y < 0 || y >= Hb || x < 0 || x >= Wb) return 0.f;
    return to_f(big[(((int64_t)k.n * Hb + y) * Wb + x) * A + m.a]);
  }
  __device__ float load_b(const KCtx& k, const NCtx& n) const {
    return n.n < N ? to_f(small[(int64_t)k.k * N + n.n]) : 0.f;
  }
  __device__ float store(int m, int n, float v) const {
    atomicAdd(&dW[(int64_t)m * N + n], v);
    return v;
  }
  __device__ int sample_of(int) const { return 0; }
};

// ---------------------------------------------------------------------------------------------
// dense: C[M,N] (+)= op(A)[M,K] * op(B)[K,N]; split-K over grid.z when accumulating
// ---------------------------------------------------------------------------------------------
template <typename TA, typename TC, bool TRANS_A, bool TRANS_B>
struct DenseProb {
  static constexpr bool A_KFAST = !TRANS_A, B_KFAST = TRANS_B, HAS_STATS = false;
  const TA* Ap; const float* Bp; const float* bias; TC* C; double* stats;
  int M, N, K, accumulate, kper;
  int kb, ke;
  struct MCtx { int m; };
  struct NCtx { int n; };
  struct KCtx { int k; };
  __device__ void set_slice(int z) { kb = z * kper; ke = min(K, kb + kper); }
  __device__ int k_begin() const { return kb; }
  __device__ int k_end() const { return ke; }
  __device__ MCtx prep_m(int m) const { return MCtx{m}; }
  __device__ NCtx prep_n(int n) const { return NCtx{n}; }
  __device__ KCtx prep_k(int k) const { return KCtx{k}; }
  __device__ float load_a(const MCtx& m, const KCtx& k) const {
    if (m.m >= M) return 0.f;
    return to_f(TRANS_A ? Ap[(int64_t)k.k * M + m.m] : Ap[(int64_t)m.m * K + k.k]);
  }
  __device__ float load_b(const KCtx& k, const NCtx& n) const {
    if (n.n >= N) return 0.f;
    return TRANS_B ? Bp[(int64_t)n.n * K + k.k] : Bp[(int64_t)k.k * N + n.n];
  }
  __device__ float store(int m, int n, float v) const {
    int64_t o = (int64_t)m * N + n;
    if (accumulate) atomicAdd(reinterpret_cast<float*>(C) + o, v);   // fp32 C only (checked on host)
    else C[o] = from_f<TC>(bias ? v + bias[n] : v);
    return v;
  }
  __device__ int sample_of(int) const { return 0; }
};

// ---------------------------------------------------------------------------------------------
// general conv + folded batch-norm + ReLU (Inception-2015 pool_3 forward, fid.py:36-106): any kh x kw, stride,
// zero padding; reads a channel slice of an NHWC tensor, writes a channel slice of the block's concat output.
//   y[m=(n,i,j)][co] = relu(scale[co] * sum_k=(ky,kx,c) x[n, s*i+ky-ph, s*j+kx-pw, c] * W[k][co] + shift[co])
// ---------------------------------------------------------------------------------------------
template <typename T>
struct ConvBnProb {
  static constexpr bool A_KFAST = true, B_KFAST = false, HAS_STATS = false;
  const T* x; const float* W; const float* scale; const float* shift; T* y; double* stats;
  int M, N, K;                 // M = Nimg*Ho*Wo, N = Cout, K = kh*kw*Cin
  int H, Wd, Cin, xs, Ho, Wo, kw, s, ph, pw, ys, relu;
  struct MCtx { int n, y0, x0; bool ok; };
  struct NCtx { int n; };
  struct KCtx { int ky, kx, c, k; };
  __device__ void set_slice(int) {}
  __device__ int k_begin() const { return 0; }
  __device__ int k_end() const { return K; }
  __device__ MCtx prep_m(int m) const {
    MCtx c; c.ok = m < M;
    int j = m % Wo, t = m / Wo, i = t % Ho; c.n = t / Ho;
    c.y0 = s * i - ph; c.x0 = s * j - pw; return c;
  }
  __device__ NCtx prep_n(int n) const { return NCtx{n}; }
  __device__ KCtx prep_k(int k) const {
    KCtx c; c.k = k; int tap = k / Cin; c.c = k - tap * Cin; c.ky = tap / kw; c.kx = tap - kw * c.ky; return c;
  }
  __device__ float load_a(const MCtx& m, const KCtx& k) const {
    int yy = m.y0 + k.ky, xx = m.x0 + k.kx;
    if (!m.ok || yy < 0 || yy >= H || xx < 0 || xx >= Wd) return 0.f;
    return to_f(x[(((int64_t)m.n * H + yy) * Wd + xx) * xs + k.c]);
  }
  __device__ float load_b(const KCtx& k, const NCtx& n) const {
    return n.n < N ? W[(int64_t)k.k * N + n.n] : 0.f;
  }
  __device__ float store(int m, int n, float v) const {
    v = fmaf(v, scale[n], shift[n]);
    if (relu) v = fmaxf(v, 0.f);
    y[(int64_t)m * ys + n] = from_f<T>(v);
    return v;
  }
  __device__ int sample_of(int) const { return 0; }
};

template <class P>
int launch(P p, int zslices, cudaStream_t st) {
  dim3 grid((p.M + TM - 1) / TM, (p.N + TN - 1) / TN, zslices);
  simt_gemm_kernel<P><<<grid, NT, 0, st>>>(p);
  return 0;
}

inline int choose_slices(int tiles, int K, int min_k) {
  int want = (2 * lg_num_sms() + tiles - 1) / tiles;
  int maxs = (K + min_k - 1) / min_k;
  int z = want < 1 ? 1 : want;
  if (z > maxs) z = maxs;
  if (z < 1) z = 1;
  return z;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// host entry points used by api.cu
// -------------------------------------------------------------------------------------------------
template <typename T>
static int simt_fprop_t(const void* big, const float* W, const float* bias, void* out, double* stats,
                        int N, int Hb, int Wb, int A, int B, int s, cudaStream_t st) {
  FpropProb<T> p;
  p.big = (const T*)big; p.W = W; p.bias = bias; p.out = (T*)out; p.stats = stats;
  p.Hb = Hb; p.Wb = Wb; p.A = A; p.Hs = Hb / s; p.Ws = Wb / s; p.s = s; p.pad = (s == 2) ? 1 : 2;
  p.M = N * p.Hs * p.Ws; p.N = B; p.K = 25 * A;
  return launch(p, 1, st);
}

template <typename T>
static int simt_dgrad_t(const void* small, const float* W, const float* bias, void* out, double* stats,
                        int N, int Hb, int Wb, int A, int B, int s, int act, cudaStream_t st) {
  DgradProb<T> p;
  p.small = (const T*)small; p.W = W; p.bias = bias; p.out = (T*)out; p.stats = stats;
  p.Hb = Hb; p.Wb = Wb; p.A = A; p.B = B; p.Hs = Hb / s; p.Ws = Wb / s; p.s = s;
  p.pad = (s == 2) ? 1 : 2; p.act = act;
  p.M = N * p.Hs * p.Ws; p.N = A; p.K = 0;
  p.py = p.px = p.nky = p.nkx = 0;
  return launch(p, s * s, st);
}

template <typename T>
static int simt_wgrad_t(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A,
                        int B, int s, cudaStream_t st) {
  WgradProb<T> p;
  p.big = (const T*)big; p.small = (const T*)small; p.dW = dW; p.stats = nullptr;
  p.Hb = Hb; p.Wb = Wb; p.A = A; p.Hs = Hb / s; p.Ws = Wb / s; p.s = s; p.pad = (s == 2) ? 1 : 2;
  p.M = 25 * A; p.N = B; p.K = N * p.Hs * p.Ws;
  int tiles = ((p.M + TM - 1) / TM) * ((p.N + TN - 1) / TN);
  int z = choose_slices(tiles, p.K, 256);
  p.kper = ((p.K + z - 1) / z + TK - 1) / TK * TK;
  z = (p.K + p.kper - 1) / p.kper;
  p.kb = p.ke = 0;
  return launch(p, z, st);
}

int lg_simt_fprop(const void* big, const float* W, const float* bias, void* out, double* stats, int N,
                  int Hb, int Wb, int A, int B, int s, int dtype, cudaStream_t st) {
  return dtype == LG_BF16 ? simt_fprop_t<bf16>(big, W, bias, out, stats, N, Hb, Wb, A, B, s, st)
                          : simt_fprop_t<float>(big, W, bias, out, stats, N, Hb, Wb, A, B, s, st);
}
int lg_simt_dgrad(const void* small, const float* W, const float* bias, void* out, double* stats, int N,
                  int Hb, int Wb, int A, int B, int s, int act, int dtype, cudaStream_t st) {
  return dtype == LG_BF16 ? simt_dgrad_t<bf16>(small, W, bias, out, stats, N, Hb, Wb, A, B, s, act, st)
                          : simt_dgrad_t<float>(small, W, bias, out, stats, N, Hb, Wb, A, B, s, act, st);
}
int lg_simt_wgrad(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A, int B,
                  int s, int dtype, cudaStream_t st) {
  return dtype == LG_BF16 ? simt_wgrad_t<bf16>(big, small, dW, N, Hb, Wb, A, B, s, st)
                          : simt_wgrad_t<float>(big, small, dW, N, Hb, Wb, A, B, s, st);
}

template <typename T>
static int simt_conv_bn_t(const void* x, const float* W, const float* scale, const float* shift, void* y, int N,
                          int H, int Wd, int Cin, int xs, int xo, int kh, int kw, int s, int ph, int pw, int Cout,
                          int ys, int yo, int relu, cudaStream_t st) {
  ConvBnProb<T> p;
  p.x = (const T*)x + xo; p.W = W; p.scale = scale; p.shift = shift; p.y = (T*)y + yo; p.stats = nullptr;
  p.H = H; p.Wd = Wd; p.Cin = Cin; p.xs = xs; p.kw = kw; p.s = s; p.ph = ph; p.pw = pw; p.ys = ys; p.relu = relu;
  p.Ho = (H + 2 * ph - kh) / s + 1; p.Wo = (Wd + 2 * pw - kw) / s + 1;
  p.M = N * p.Ho * p.Wo; p.N = Cout; p.K = kh * kw * Cin;
  return launch(p, 1, st);
}
int lg_simt_conv_bn(const void* x, const float* W, const float* scale, const float* shift, void* y, int N, int H,
                    int Wd, int Cin, int xs, int xo, int kh, int kw, int s, int ph, int pw, int Cout, int ys, int yo,
                    int relu, int dtype, cudaStream_t st) {
  return dtype == LG_BF16
             ? simt_conv_bn_t<bf16>(x, W, scale, shift, y, N, H, Wd, Cin, xs, xo, kh, kw, s, ph, pw, Cout, ys, yo, relu, st)
             : simt_conv_bn_t<float>(x, W, scale, shift, y, N, H, Wd, Cin, xs, xo, kh, kw, s, ph, pw, Cout, ys, yo, relu, st);
}

template <typename TA, typename TC, bool TRA, bool TRB>
static int dense_t(const void* A, const float* Bm, const float* bias, void* C, int M, int N, int K,
                   int accumulate, cudaStream_t st) {
  DenseProb<TA, TC, TRA, TRB> p;
  p.Ap = (const TA*)A; p.Bp = Bm; p.bias = bias; p.C = (TC*)C; p.stats = nullptr;
  p.M = M; p.N = N; p.K = K; p.accumulate = accumulate;
  int z = 1;
  if (accumulate) {
    int tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN);
    z = choose_slices(tiles, K, 256);
  }
  p.kper = ((K + z - 1) / z + TK - 1) / TK * TK;
  z = (K + p.kper - 1) / p.kper;
  p.kb = p.ke = 0;
  return launch(p, z, st);
}

template <typename TA, typename TC>
static int dense_tt(const void* A, const float* Bm, const float* bias, void* C, int M, int N, int K, int tA,
                    int tB, int acc, cudaStream_t st) {
  if (tA) return tB ? dense_t<TA, TC, true, true>(A, Bm, bias, C, M, N, K, acc, st)
                    : dense_t<TA, TC, true, false>(A, Bm, bias, C, M, N, K, acc, st);
  return tB ? dense_t<TA, TC, false, true>(A, Bm, bias, C, M, N, K, acc, st)
            : dense_t<TA, TC, false, false>(A, Bm, bias, C, M, N, K, acc, st);
}

int lg_simt_dense(const void* A, const float* Bm, const float* bias, void* C, int M, int N, int K, int tA,
                  int tB, int acc, int a_dtype, int c_dtype, cudaStream_t st) {
  if (a_dtype == LG_BF16)
    return c_dtype == LG_BF16 ? dense_tt<bf16, bf16>(A, Bm, bias, C, M, N, K, tA, tB, acc, st)
                              : dense_tt<bf16, float>(A, Bm, bias, C, M, N, K, tA, tB, acc, st);
  return c_dtype == LG_BF16 ? dense_tt<float, bf16>(A, Bm, bias, C, M, N, K, tA, tB, acc, st)
                            : dense_tt<float, float>(A, Bm, bias, C, M, N, K, tA, tB, acc, st);
}
