// FID feature statistics (fid.py:169-188): streaming fp64 accumulation of sum(x) and sum(x x^T)
// over fp32 feature rows, so that mean / unbiased covariance match np.mean / np.cov (fp64).
// S2 is a rank-n symmetric update: only its upper triangle (tiles with bx <= by) is accumulated;
// lg_fid_finalize mirrors it.
#include "common.cuh"

namespace {

constexpr int FT = 64, FK = 16, FNT = 256;

__global__ void __launch_bounds__(FNT) fid_accumulate_kernel(const float* __restrict__ X, const double* __restrict__ shift,
                                                             double* S1, double* S2, int64_t n, int d, int64_t rows_per) {
  if (blockIdx.x > blockIdx.y) return;
  __shared__ __align__(16) double As[FK][FT];
  __shared__ __align__(16) double Bs[FK][FT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * FT, j0 = blockIdx.y * FT;
  const bool diag = blockIdx.x == blockIdx.y;
  const int64_t rbeg = (int64_t)blockIdx.z * rows_per, rend = min(n, rbeg + rows_per);
  const int lc = tid & 63, lr = tid >> 6;
  const int ci = i0 + lc, cj = j0 + lc;
  const double shi = (shift && ci < d) ? shift[ci] : 0.0, shj = (shift && cj < d) ? shift[cj] : 0.0;

  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  double colsum = 0.0;

  for (int64_t r0 = rbeg; r0 < rend; r0 += FK) {
    double av[4], bv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int64_t r = r0 + lr + 4 * q;
      bool ok = r < rend;
      av[q] = (ok && ci < d) ? (double)X[r * d + ci] - shi : 0.0;
      bv[q] = diag ? av[q] : ((ok && cj < d) ? (double)X[r * d + cj] - shj : 0.0);
      colsum += av[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) { As[lr + 4 * q][lc] = av[q]; Bs[lr + 4 * q][lc] = bv[q]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      double a[4], b[4];
      double2 t0 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
      double2 t1 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
      a[0] = t0.x; a[1] = t0.y; a[2] = t1.x; a[3] = t1.y;
      t0 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
      t1 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
      b[0] = t0.x; b[1] = t0.y; b[2] = t1.x; b[3] = t1.y;
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
    }
  }

#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int i = i0 + ty * 4 + p;
    if (i >= d) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int j = j0 + tx * 4 + q;
      if (j >= d) continue;
      if (j >= i) atomicAdd(&S2[(int64_t)i * d + j], acc[p][q]);       // upper triangle only (finalize mirrors)
    }
  }
  if (diag && S1 != nullptr) {
    // colsum of thread (lc, lr) covers rows lr, lr+4, ...; fold the 4 row-lanes through smem
    __syncthreads();
    As[lr][lc] = colsum;
    __syncthreads();
    if (lr == 0 && ci < d) atomicAdd(&S1[ci], As[0][lc] + As[1][lc] + As[2][lc] + As[3][lc]);
  }
}

__global__ void fid_finalize_kernel(const double* __restrict__ S1, const double* __restrict__ S2,
                                    const double* __restrict__ shift, double* mu, double* sigma, int64_t n, int d) {
  const int64_t total = (int64_t)d * d;
  const double inv_n = 1.0 / (double)n, inv_nm1 = 1.0 / (double)(n - 1);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(e / d), j = (int)(e % d);
    // only the upper triangle of S2 is accumulated (atomics arrive in any order, so a mirrored copy would
    // differ in the last bit): reading (min, max) makes sigma exactly symmetric, like np.cov's
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    sigma[e] = (S2[(int64_t)lo * d + hi] - S1[lo] * S1[hi] * inv_n) * inv_nm1;
    if (e < d) mu[e] = (shift ? shift[e] : 0.0) + S1[e] * inv_n;
  }
}

}  // namespace

extern "C" int lg_fid_accumulate(const float* X, const double* shift, double* S1, double* S2, int64_t n, int d,
                                 void* stream) {
  LG_REQUIRE(X && S2 && n > 0 && d > 0, "bad arguments");
  int tiles = (d + FT - 1) / FT;
  int tri = tiles * (tiles + 1) / 2;
  int64_t want = (4LL * lg_num_sms() + tri - 1) / tri;
  int64_t rows_per = (n + want - 1) / want;
  if (rows_per < 256) rows_per = 256;
  rows_per = (rows_per + FK - 1) / FK * FK;
  int z = (int)((n + rows_per - 1) / rows_per);
  fid_accumulate_kernel<<<dim3(tiles, tiles, z), FNT, 0, (cudaStream_t)stream>>>(X, shift, S1, S2, n, d, rows_per);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_fid_finalize(const double* S1, const double* S2, const double* shift, double* mu, double* sigma,
                               int64_t n, int d, void* stream) {
  LG_REQUIRE(S1 && S2 && mu && sigma && n > 1 && d > 0, "bad arguments");
  int gsz = (int)(((int64_t)d * d + 255) / 256);
  if (gsz > lg_num_sms() * 8) gsz = lg_num_sms() * 8;
  fid_finalize_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>(S1, S2, shift, mu, sigma, n, d);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

// ------------------------------------------------------------------------------------------------
// Frechet distance (fid.py:112-163): Tr sqrtm(sigma1 sigma2) without a library eigen-solver.
// The spectrum of sigma1 sigma2 is that of the symmetric PSD matrix R sigma2 R with R = sigma1^(1/2), and
// both square roots come from the coupled Newton-Schulz iteration (Higham, "Functions of Matrices", 6.35)
//     Y0 = A / c,  Z0 = I;   T = (3 I - Z Y) / 2;   Y <- Y T,  Z <- T Z;   Y -> (A / c)^(1/2)
// which is nothing but fp64 matrix products: lg_dgemm below (C = alpha A B + diag I) and two small
// reductions.  The iteration is driven from the host side (littlegan_b200/fid.py) through these entry points.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int GT = 128, GK = 16, GNT = 256;

// C[n,n] = alpha * A[n,n] @ B[n,n] + diag * I, row-major fp64.  128x128 tile per CTA, 8x8 outputs per thread
// (two 4-wide groups 64 apart in each direction: a half-warp's 32-byte shared-memory reads are contiguous),
// register-staged double buffering of the global loads.
__global__ void __launch_bounds__(GNT) dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                    double* __restrict__ C, int n, double alpha, double diag) {
  __shared__ __align__(16) double As[GK][GT + 2];      // As[k][m] (transposed on the way in)
  __shared__ __align__(16) double Bs[GK][GT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  // loader mapping.  A tile [128 rows][16 k]: thread -> row tid/2, k half (tid&1)*8 .. +7 (64 contiguous bytes).
  // B tile [16 k][128 cols]: thread -> k row tid/16, cols (tid%16)*8 .. +7.
  const int ar = tid >> 1, ak = (tid & 1) * 8;
  const int bk = tid >> 4, bc = (tid & 15) * 8;
  double ra[8], rb[8];
  auto gload = [&](int k0) {
    const int gr = m0 + ar;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gk = k0 + ak + i;
      ra[i] = (gr < n && gk < n) ? A[(int64_t)gr * n + gk] : 0.0;
    }
    const int gkb = k0 + bk;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gc = n0 + bc + i;
      rb[i] = (gkb < n && gc < n) ? B[(int64_t)gkb * n + gc] : 0.0;
    }
  };
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  gload(0);
  for (int k0 = 0; k0 < n; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) As[ak + i][ar] = ra[i];
#pragma unroll
    for (int i = 0; i < 8; i += 2) *reinterpret_cast<double2*>(&Bs[bk][bc + i]) = make_double2(rb[i], rb[i + 1]);
    __syncthreads();
    if (k0 + GK < n) gload(k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      double a[8], b[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double2 a0 = *reinterpret_cast<const double2*>(&As[k][h * 64 + ty * 4]);
        const double2 a1 = *reinterpret_cast<const double2*>(&As[k][h * 64 + ty * 4 + 2]);
        a[4 * h] = a0.x; a[4 * h + 1] = a0.y; a[4 * h + 2] = a1.x; a[4 * h + 3] = a1.y;
        const double2 b0 = *reinterpret_cast<const double2*>(&Bs[k][h * 64 + tx * 4]);
        const double2 b1 = *reinterpret_cast<const double2*>(&Bs[k][h * 64 + tx * 4 + 2]);
        b[4 * h] = b0.x; b[4 * h + 1] = b0.y; b[4 * h + 2] = b1.x; b[4 * h + 3] = b1.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (c < n) C[(int64_t)r * n + c] = alpha * acc[i][j] + (r == c ? diag : 0.0);
    }
  }
}

// out[0] = trace(A), out[1] = ||A||_F^2, out[2] = max |A_ij - A_ji| (a symmetry check); out zeroed by the launch.
__global__ void __launch_bounds__(256) dmat_stats_kernel(const double* __restrict__ A, int n, double* out) {
  double tr = 0.0, fr = 0.0, asym = 0.0;
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e % n);
    const double v = A[e];
    fr = fma(v, v, fr);
    if (i == j) tr += v;
    if (j > i) asym = fmax(asym, fabs(v - A[(int64_t)j * n + i]));
  }
  __shared__ double s[3][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
    fr += __shfl_xor_sync(0xffffffffu, fr, o);
    asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
  }
  if (lane == 0) { s[0][w] = tr; s[1][w] = fr; s[2][w] = asym; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < 8; ++i) { a += s[0][i]; b += s[1][i]; c = fmax(c, s[2][i]); }
    atomicAdd(&out[0], a);
    atomicAdd(&out[1], b);
    // fmax through the bit pattern of a non-negative double
    atomicMax(reinterpret_cast<unsigned long long*>(&out[2]), (unsigned long long)__double_as_longlong(c));
  }
}

// dst = alpha * src + diag * I (also the symmetrisation dst = (src + src^T) / 2 when `sym`).
__global__ void __launch_bounds__(256) dmat_axpby_kernel(const double* __restrict__ src, double* __restrict__ dst, int n,
                                                         double alpha, double diag, int sym) {
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e % n);
    double v = src ? src[e] : 0.0;
    if (sym) v = 0.5 * (v + src[(int64_t)j * n + i]);
    dst[e] = alpha * v + (i == j ? diag : 0.0);
  }
}

}  // namespace

extern "C" int lg_dgemm(const double* A, const double* B, double* C, int n, double alpha, double diag, void* stream) {
  LG_REQUIRE(A && B && C && n > 0 && C != A && C != B, "bad arguments (C must not alias an operand)");
  const int t = (n + GT - 1) / GT;
  dgemm_kernel<<<dim3(t, t), GNT, 0, (cudaStream_t)stream>>>(A, B, C, n, alpha, diag);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_dmat_stats(const double* A, int n, double* out3, void* stream) {
  LG_REQUIRE(A && out3 && n > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out3, 0, 3 * sizeof(double), st) != cudaSuccess) {
    lg_set_error("%s: cudaMemsetAsync failed", __func__);
    return LG_ERR_CUDA;
  }
  int gsz = (int)(((int64_t)n * n + 255) / 256);
  if (gsz > lg_num_sms() * 8) gsz = lg_num_sms() * 8;
  dmat_stats_kernel<<<gsz, 256, 0, st>>>(A, n, out3);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_dmat_scale_shift(const double* src, double* dst, int n, double alpha, double diag, int symmetrise,
                                   void* stream) {
  LG_REQUIRE(dst && n > 0 && (src || !symmetrise) && !(symmetrise && src == dst), "bad arguments");
  int gsz = (int)(((int64_t)n * n + 255) / 256);
  if (gsz > lg_num_sms() * 8) gsz = lg_num_sms() * 8;
  dmat_axpby_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>(src, dst, n, alpha, diag, symmetrise);
  LG_LAUNCH_CHECK();
  return LG_OK;
}
