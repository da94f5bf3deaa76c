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
