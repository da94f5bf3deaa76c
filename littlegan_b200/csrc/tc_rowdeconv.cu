// Row-streaming transposed convolutions with THREE output channels.  Stride 1 (first kernel): the generator's final
// Conv2DTranspose(3, k5, s1, tanh) (model.py:86,104) on a 128-pixel-wide, 32-channel input.  The layer is
// HBM bound (1 MB in, 48 KB out per image); the GEMM + col2im kernel (tc_deconv_small.cu) spends its time in a
// shared-memory overlap-add of 75 partial sums per pixel.  Here the horizontal half of the 5x5 window is
// folded into the MMA and the vertical half into registers, with no shared-memory traffic at all:
//
//   T_y[X][(ky,a)] = sum_{kx,b} small[y][X+2-kx][b] * W[ky][kx][a][b]       per INPUT row y: 10 MMAs
//                    (M = 128 pixels, N = 16 >= 5*3, K = 16), the A operand of tap kx being the input row in
//                    its shared-memory ring slot read at a start address shifted by kx pixels: the slot is
//                    [pixel][64 B] exactly as TMA writes it with SWIZZLE_64B = a swizzled K-major operand whose
//                    descriptor may start at any 64-byte row (the swizzle is a function of the absolute address);
//   out[Y][X][a]   = bias[a] + sum_ky T_{Y+2-ky}[X][(ky,a)]                   the same TMEM lane (thread X) for
//                    every term: five rolling 3-channel accumulators per thread, one row retired per input row.
//
// Every input row is fetched once by one TMA box and freed as soon as its 10 MMAs retire.  Epilogue: tanh, bf16
// store of the RGB row (6-byte pixels, contiguous per warp) and optionally of an 8-channel zero-padded copy
// (16-byte pixels) that the row-streaming forward kernel of the next layer fetches by TMA.
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "internal.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int RD_THREADS = 192;          // warps 0-3 epilogue, 4 row producer, 5 MMA issuer
constexpr int RD_ACC = 4;                // TMEM accumulator stages (16 columns each)
constexpr int RD_W = 128;                // map width = MMA M
constexpr int RD_B = 32;                 // input channels
constexpr int RD_ROWB = (RD_W + 8) * RD_B * 2;   // one input row: 136 pixels x 64 B
constexpr int RD_KTOT = 5 * RD_B;        // K of the weight operand: (kx, b)
constexpr int RD_RING = 5;
constexpr int RD_CTAS_PER_SM = 4;        // 49 KB of shared memory and 64 TMEM columns each: the epilogue is latency bound

struct RdParams {
  int Nimg, H, R, strips_per_img, total_strips, act;
  const float* W;                        // [5][5][3][32] fp32
  const float* bias;
  bf16* out3;                            // [N,H,128,3]
  uint4* out8;                           // [N,H,128,8] or NULL
  double* stats;
};

__global__ void __launch_bounds__(RD_THREADS)
tc_rowdeconv_kernel(const __grid_constant__ CUtensorMap tmIn, const RdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sRing = smem;                                          // RING x [136 px][64 B], SWIZZLE_64B (slots 512-B aligned)
  uint8_t* sW = sRing + RD_RING * RD_ROWB;                        // 16 rows x 160 k, no-swizzle K-major (5 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 16 * RD_KTOT * 2);
  uint64_t* full = bars;                     // [RING] TMA -> MMA
  uint64_t* empty = full + RD_RING;          // [RING] MMA -> TMA
  uint64_t* tfull = empty + RD_RING;         // [ACC]  MMA -> epilogue
  uint64_t* tempty = tfull + RD_ACC;         // [ACC]  epilogue -> MMA (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + RD_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights: W[ky][kx][a][b] -> bf16 operand row n = ky*3 + a (row 15 = 0), k = kx*32 + b.  All loads are issued
  // before the first store: a load-store-load chain would pay one memory round trip per element.
  {
    constexpr int TOT = 16 * RD_KTOT, PER = (TOT + RD_THREADS - 1) / RD_THREADS;
    float wv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = threadIdx.x + i * RD_THREADS;
      const int n = e / RD_KTOT, k = e - n * RD_KTOT;
      const int ky = n / 3, a = n - ky * 3, kx = k >> 5, b = k & 31;
      wv[i] = (e < TOT && n < 15) ? __ldg(p.W + ((ky * 5 + kx) * 3 + a) * RD_B + b) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = threadIdx.x + i * RD_THREADS;
      const int n = e / RD_KTOT, k = e - n * RD_KTOT;
      const int off = (n >> 3) * ((RD_KTOT / 8) * 128) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
      if (e < TOT) *reinterpret_cast<bf16*>(sW + off) = __float2bfloat16_rn(wv[i]);
    }
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmIn);
    for (int i = 0; i < RD_RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < RD_ACC; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 64);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nin = p.R + 4;                   // input rows a strip walks (rows outside the image are virtual)

  if (warp == 4) {
    // ------------------------------------------------------------ input-row producer (real rows only)
    if (tc::elect_one()) {
      const uint32_t ring_addr = tc::smem_u32(sRing), full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      int slot = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, Y0 = (t - n * p.strips_per_img) * p.R;
        for (int k = 0; k < nin; ++k) {
          const int y = Y0 - 2 + k;
          if (y < 0 || y >= p.H) continue;
          tc::mbar_wait_addr(empty_addr + slot * 8, phase ^ 1);
          tc::mbar_expect_tx_addr(full_addr + slot * 8, (uint32_t)RD_ROWB);
          tc::tma_load_3d_addr(ring_addr + slot * RD_ROWB, &tmIn, full_addr + slot * 8, 0, -2, n * p.H + y);
          if (++slot == RD_RING) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer: 10 MMAs per real input row
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, 16, 0, 0);
      const uint32_t ring_lo = tc::smem_u32(sRing) >> 4;
      const uint32_t full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      const uint32_t tfull_addr = tc::smem_u32(tfull), tempty_addr = tc::smem_u32(tempty);
      const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);                   // SBO = 8 rows x 64 B, SWIZZLE_64B
      const uint32_t b_hi = (((uint32_t)(RD_KTOT / 8) * 128u) >> 4) | (1u << 14);   // SBO = 20 cores
      const uint32_t b_lo0 = (tc::smem_u32(sW) >> 4) | ((128u >> 4) << 16);         // LBO = 128 B
      constexpr uint32_t A_LBO = 1u << 16;                                          // unused for swizzled K-major
      int slot = 0; uint32_t phase = 0;
      int acc = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, Y0 = (t - n * p.strips_per_img) * p.R;
        (void)n;
        for (int k = 0; k < nin; ++k) {
          const int y = Y0 - 2 + k;
          if (y < 0 || y >= p.H) continue;
          tc::mbar_wait_addr(full_addr + slot * 8, phase);
          tc::mbar_wait_addr(tempty_addr + acc * 8, aphase ^ 1);
          tc::fence_after_sync();
          const uint32_t sa_lo = ring_lo + (uint32_t)slot * (uint32_t)(RD_ROWB >> 4);
          const uint32_t tacc = tmem_base + (uint32_t)(acc * 16);
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
            for (int cp = 0; cp < 2; ++cp) {
              const uint32_t a_off = (uint32_t)((4 - kx) * 64 + cp * 32) >> 4;       // pixel shift + K step inside the 64-B row
              tc::mma_bf16_lohi(tacc, sa_lo + (a_off | A_LBO), a_hi, b_lo0 + (uint32_t)(kx * 2 + cp) * 16u, b_hi, idesc,
                                (kx | cp) ? 1u : 0u);
            }
          }
          tc::mma_commit_addr(tfull_addr + acc * 8);
          tc::mma_commit_addr(empty_addr + slot * 8);
          if (++slot == RD_RING) { slot = 0; phase ^= 1; }
          if (++acc == RD_ACC) { acc = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread X owns output column X
    const int X = warp * 32 + lane;
    const uint32_t tlane = (uint32_t)(warp * 32) << 16;
    const float b0 = p.bias ? p.bias[0] : 0.f, b1 = p.bias ? p.bias[1] : 0.f, b2 = p.bias ? p.bias[2] : 0.f;
    int acc = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
      const int n = t / p.strips_per_img, Y0 = (t - n * p.strips_per_img) * p.R;
      // r[j] = partial sums of output row (y - 2 + j) while input row y is being added
      float r[5][3];
#pragma unroll
      for (int j = 0; j < 5; ++j) { r[j][0] = 0.f; r[j][1] = 0.f; r[j][2] = 0.f; }
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < nin; ++k) {
        const int y = Y0 - 2 + k;
        if (y >= 0 && y < p.H) {
          tc::mbar_wait(&tfull[acc], aphase);
          tc::fence_after_sync();
          float v[16];
          tc::tmem_ld16(tmem_base + tlane + (uint32_t)(acc * 16), v);
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tempty[acc]);
          if (++acc == RD_ACC) { acc = 0; aphase ^= 1; }
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            r[j][0] += v[3 * j]; r[j][1] += v[3 * j + 1]; r[j][2] += v[3 * j + 2];
          }
        }
        // output row Y = y - 2 is complete (its last contribution comes from input row Y + 2 = y)
        const int Y = y - 2;
        if (Y >= Y0) {                                             // always < Y0 + R here
          float o0 = r[0][0] + b0, o1 = r[0][1] + b1, o2 = r[0][2] + b2;
          s1 += o0 + o1 + o2;
          s2 = fmaf(o0, o0, fmaf(o1, o1, fmaf(o2, o2, s2)));
          if (p.act == LG_ACT_TANH) { o0 = tanhf(o0); o1 = tanhf(o1); o2 = tanhf(o2); }
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(o0, o1);
          const bf16 h2 = __float2bfloat16_rn(o2);
          const int64_t pix = ((int64_t)n * p.H + Y) * RD_W + X;
          bf16* d = p.out3 + pix * 3;
          d[0] = h01.x; d[1] = h01.y; d[2] = h2;
          if (p.out8 != nullptr)
            p.out8[pix] = make_uint4(*reinterpret_cast<const uint32_t*>(&h01), (uint32_t)__bfloat16_as_ushort(h2), 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { r[j][0] = r[j + 1][0]; r[j][1] = r[j + 1][1]; r[j][2] = r[j + 1][2]; }
        r[4][0] = 0.f; r[4][1] = 0.f; r[4][2] = 0.f;
      }
      if (p.stats != nullptr) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { atomicAdd(&p.stats[2 * n], (double)s1); atomicAdd(&p.stats[2 * n + 1], (double)s2); }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 64);
  }
}

// =====================================================================================================
// Stride-2 variant: the input gradient of the encoder's first Conv2D (model.py:15): small [N,64,64,64] ->
// RGB gradient [N,128,128,3].  Thread j owns the output columns 2j and 2j+1.  Per INPUT row i one accumulator
//   T_i[j][(ky,px,a)] = sum_{dx} sum_b small[i][j+dx][b] * Wsh[dx][(ky,px,a)][b]      (N = 32 >= 5*2*3)
// with the taps grouped by the input pixel they read: dx = 0 -> kx 1 (px 0), kx 2 (px 1); dx = -1 -> kx 3, 4;
// dx = +1 -> kx 0 (px 1): 3 shifts x 4 K steps = 12 MMAs (M 64 x N 32 x K 16) per input row, A operand = the
// SWIZZLE_128B ring slot read at a pixel-shifted start address (see tc_rowdgrad.cu).  The vertical half is the
// rolling-register scheme of the stride-1 kernel: input row i adds T_i[(ky,.)] to output row 2i+ky-1, and rows
// 2i-1 and 2i retire after input row i.
// =====================================================================================================
constexpr int R2_WS = 64, R2_B = 64;
constexpr int R2_ROWB = (R2_WS + 8) * R2_B * 2;   // 72 pixels x 128 B
constexpr int R2_RING = 6;
constexpr int R2_WELEMS = 3 * 4 * 32 * 16;        // [shift][k step][n 32][16 b]

__global__ void __launch_bounds__(RD_THREADS)
tc_rowdeconv2_kernel(const __grid_constant__ CUtensorMap tmIn, const RdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sRing = smem;                                          // RING x 9 KB, SWIZZLE_128B
  uint8_t* sW = sRing + R2_RING * R2_ROWB;                        // 12 operand blocks of 1 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + R2_WELEMS * 2);
  uint64_t* full = bars;
  uint64_t* empty = full + R2_RING;
  uint64_t* tfull = empty + R2_RING;
  uint64_t* tempty = tfull + RD_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + RD_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights -> bf16 blocks [shift][ks][n = (ky*2+px)*3+a][16 b], no-swizzle K-major cores; loads before stores
  {
    constexpr int PER = R2_WELEMS / RD_THREADS;
    static_assert(R2_WELEMS % RD_THREADS == 0, "weight operand must divide over the CTA");
    float wv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = threadIdx.x + i * RD_THREADS;
      const int k = e & 15, n = (e >> 4) & 31, ks = (e >> 9) & 3, sh = e >> 11;
      const int a = n % 3, kp = n / 3, px = kp & 1, ky = kp >> 1;
      const int kx = sh == 0 ? 1 + px : sh == 1 ? 3 + px : (px ? 0 : -1);
      wv[i] = (n < 30 && kx >= 0) ? __ldg(p.W + ((ky * 5 + kx) * 3 + a) * R2_B + ks * 16 + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = threadIdx.x + i * RD_THREADS;
      const int k = e & 15, n = (e >> 4) & 31, blk = e >> 9;
      const int off = blk * 1024 + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
      *reinterpret_cast<bf16*>(sW + off) = __float2bfloat16_rn(wv[i]);
    }
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmIn);
    for (int i = 0; i < R2_RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < RD_ACC; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 128);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int Hs = p.H / 2;
  const int nin = p.R + 2;                   // input rows I0-1 .. I0+R of a strip of R input rows

  if (warp == 4) {
    if (tc::elect_one()) {
      const uint32_t ring_addr = tc::smem_u32(sRing), full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      int slot = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, I0 = (t - n * p.strips_per_img) * p.R;
        for (int k = 0; k < nin; ++k) {
          const int y = I0 - 1 + k;
          if (y < 0 || y >= Hs) continue;
          tc::mbar_wait_addr(empty_addr + slot * 8, phase ^ 1);
          tc::mbar_expect_tx_addr(full_addr + slot * 8, (uint32_t)R2_ROWB);
          tc::tma_load_3d_addr(ring_addr + slot * R2_ROWB, &tmIn, full_addr + slot * 8, 0, -1, n * Hs + y);
          if (++slot == R2_RING) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(64, 32, 0, 0);
      const uint32_t ring_lo = tc::smem_u32(sRing) >> 4;
      const uint32_t full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      const uint32_t tfull_addr = tc::smem_u32(tfull), tempty_addr = tc::smem_u32(tempty);
      const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);                  // SWIZZLE_128B, SBO = 8 rows
      const uint32_t b_hi = (256u >> 4) | (1u << 14);                                // SBO = 2 cores
      const uint32_t b_lo0 = (tc::smem_u32(sW) >> 4) | ((128u >> 4) << 16);          // LBO = 128 B
      int slot = 0; uint32_t phase = 0;
      int acc = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, I0 = (t - n * p.strips_per_img) * p.R;
        (void)n;
        for (int k = 0; k < nin; ++k) {
          const int y = I0 - 1 + k;
          if (y < 0 || y >= Hs) continue;
          tc::mbar_wait_addr(full_addr + slot * 8, phase);
          tc::mbar_wait_addr(tempty_addr + acc * 8, aphase ^ 1);
          tc::fence_after_sync();
          const uint32_t sa_lo = (ring_lo + (uint32_t)slot * (uint32_t)(R2_ROWB >> 4)) | (1u << 16);
          const uint32_t tacc = tmem_base + (uint32_t)(acc * 32);
#pragma unroll
          for (int sh = 0; sh < 3; ++sh) {
            const int dx = sh == 0 ? 0 : sh == 1 ? -1 : 1;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_bf16_lohi(tacc, sa_lo + (uint32_t)((1 + dx) * 8 + ks * 2), a_hi,
                                b_lo0 + (uint32_t)((sh * 4 + ks) * 64), b_hi, idesc, (sh | ks) ? 1u : 0u);
          }
          tc::mma_commit_addr(tfull_addr + acc * 8);
          tc::mma_commit_addr(empty_addr + slot * 8);
          if (++slot == R2_RING) { slot = 0; phase ^= 1; }
          if (++acc == RD_ACC) { acc = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // epilogue: an M = 64 accumulator lives in lanes 0-15 of every quadrant: thread (warp q, lane l < 16) = column pair j
    const int j = warp * 16 + (lane & 15);
    const bool active = lane < 16;
    const uint32_t tlane = (uint32_t)(warp * 32) << 16;
    float bia[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) bia[a] = p.bias ? p.bias[a] : 0.f;
    int acc = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
      const int n = t / p.strips_per_img, I0 = (t - n * p.strips_per_img) * p.R;
      // r[m][px*3+a] = partial sums of output row (2y - 1 + m) while input row y is being added
      float r[5][6];
#pragma unroll
      for (int m = 0; m < 5; ++m)
#pragma unroll
        for (int c = 0; c < 6; ++c) r[m][c] = 0.f;
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < nin; ++k) {
        const int y = I0 - 1 + k;
        if (y >= 0 && y < Hs) {
          tc::mbar_wait(&tfull[acc], aphase);
          tc::fence_after_sync();
          float v[32];
          tc::tmem_ld32(tmem_base + tlane + (uint32_t)(acc * 32), v);
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tempty[acc]);
          if (++acc == RD_ACC) { acc = 0; aphase ^= 1; }
#pragma unroll
          for (int m = 0; m < 5; ++m)
#pragma unroll
            for (int c = 0; c < 6; ++c) r[m][c] += v[m * 6 + c];
        }
        // output rows 2y-1 and 2y are complete
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const int Y = 2 * y - 1 + m;
          if (active && Y >= 2 * I0 && Y < 2 * (I0 + p.R)) {
            float o[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              o[c] = r[m][c] + bia[c % 3];
              s1 += o[c]; s2 = fmaf(o[c], o[c], s2);
              if (p.act == LG_ACT_TANH) o[c] = tanhf(o[c]);
            }
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]),
                                 h2 = __floats2bfloat162_rn(o[4], o[5]);
            uint32_t* d = reinterpret_cast<uint32_t*>(p.out3 + (((int64_t)n * p.H + Y) * RD_W + 2 * j) * 3);   // 12-byte aligned
            d[0] = *reinterpret_cast<const uint32_t*>(&h0);
            d[1] = *reinterpret_cast<const uint32_t*>(&h1);
            d[2] = *reinterpret_cast<const uint32_t*>(&h2);
          }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
          for (int c = 0; c < 6; ++c) r[m][c] = r[m + 2][c];
#pragma unroll
        for (int c = 0; c < 6; ++c) { r[3][c] = 0.f; r[4][c] = 0.f; }
      }
      if (p.stats != nullptr) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { atomicAdd(&p.stats[2 * n], (double)s1); atomicAdd(&p.stats[2 * n + 1], (double)s2); }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 128);
  }
}

bool plan_rd(int Nimg, int Hb, int Wb, int A, int B, int s, RdParams* p, int* grid) {
  if (A != 3 || Wb != RD_W || Hb < 16 || (Hb & (Hb - 1))) return false;
  if (!((s == 1 && B == RD_B) || (s == 2 && B == R2_B))) return false;
  const int ctas = lg_num_sms() * (s == 1 ? RD_CTAS_PER_SM : 3);
  const int Hin = Hb / s, halo = s == 1 ? 4 : 2;                 // strips are counted in INPUT rows
  int bestR = 0; double best = -1.0;
  for (int R = Hin; R >= 4; R >>= 1) {
    const int tiles = Nimg * (Hin / R);
    const int waves = (tiles + ctas - 1) / ctas;
    const double eff = (double)tiles / ((double)waves * ctas) * (double)R / (R + halo);   // halo rows are recomputed
    if (eff > best) { best = eff; bestR = R; }
  }
  p->Nimg = Nimg; p->H = Hb; p->R = bestR; p->strips_per_img = Hin / bestR; p->total_strips = Nimg * p->strips_per_img;
  *grid = lg_even_grid(p->total_strips, ctas);
  return true;
}

}  // namespace

int lg_tc_rowdeconv_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  RdParams p; int grid;
  return plan_rd(Nimg, Hb, Wb, A, B, s, &p, &grid) ? 1 : 0;
}

// out8 (may be NULL): the same image with 8-channel zero-padded pixels.
int lg_tc_rowdeconv(const void* small, const float* W, const float* bias, void* out3, void* out8, double* stats,
                    int Nimg, int Hb, int Wb, int A, int B, int s, int act, cudaStream_t st) {
  RdParams p; int grid;
  if (!plan_rd(Nimg, Hb, Wb, A, B, s, &p, &grid) || W == nullptr) {
    lg_set_error("row-streaming RGB transposed conv: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  p.act = act; p.W = W; p.bias = bias; p.out3 = (bf16*)out3; p.out8 = (uint4*)out8; p.stats = stats;
  tc_host::EncodeTiledFn enc = tc_host::get_encode();
  if (!enc) { lg_set_error("cuTensorMapEncodeTiled entry point not available"); return LG_ERR_CUDA; }
  CUtensorMap tmIn;
  const int Ws = Wb / s, Hs = Hb / s;
  cuuint64_t dims[3] = {(cuuint64_t)B, (cuuint64_t)Ws, (cuuint64_t)Nimg * Hs};
  cuuint64_t strides[2] = {(cuuint64_t)B * 2, (cuuint64_t)Ws * B * 2};
  cuuint32_t box[3] = {(cuuint32_t)B, (cuuint32_t)(Ws + 8), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(small), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, s == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { lg_set_error("row-streaming RGB transposed conv: tensor map failed: %d", (int)r); return LG_ERR_CUDA; }
  const size_t shm = 16 * RD_KTOT * 2 + (size_t)RD_RING * RD_ROWB + 512 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_rowdeconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_rowdeconv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  if (s == 2) {
    if (out8 != nullptr) { lg_set_error("row-streaming RGB transposed conv: no padded copy for stride 2"); return LG_ERR_UNSUPPORTED; }
    const size_t shm2 = (size_t)R2_RING * R2_ROWB + R2_WELEMS * 2 + 512 + 1024;
    tc_rowdeconv2_kernel<<<grid, RD_THREADS, shm2, st>>>(tmIn, p);
    return LG_OK;
  }
  tc_rowdeconv_kernel<<<grid, RD_THREADS, shm, st>>>(tmIn, p);
  return LG_OK;
}
