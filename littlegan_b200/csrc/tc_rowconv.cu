// Row-streaming "shift GEMM" forward convolution for the layers whose big map is 128 pixels wide with few
// channels (<= 32): the encoder's first Conv2D (model.py:15, RGB image padded to 8 channels), the input
// gradient of the generator's final Conv2DTranspose(3, s1) (model.py:86) and the input gradient of the
// decoder's last Conv2DTranspose(32, s2) (model.py:39).  These layers are HBM bound (AI 60-270 FLOP/B); the
// generic implicit-GEMM kernel re-fetches its input once per tap through TMA (25 x the L2->SM traffic) and
// the in-smem im2col kernel spends its time building patches.  Here NO im2col exists anywhere:
//
//   * a CTA walks a strip of output rows of one image; every INPUT row is brought into a shared-memory
//     ring exactly once by one TMA box, de-interleaved on the fly into planes [x parity][8-channel chunk]
//     [pixel][16 B] (a 4-D tensor map whose third dimension is the 16-byte (parity, chunk) index);
//   * a plane is already the canonical no-swizzle K-major tcgen05 operand (8 rows x 16 B core matrices,
//     SBO = 128 B): the A operand of tap (ky,kx) is the SAME bytes at a start address shifted by kx
//     pixels, in the ring slot of input row s*i+ky-pad.  With 8-channel pixels the two K halves of one
//     K = 16 MMA are two neighbouring taps (LBO = 16 B); with 32 channels they are two chunk planes;
//   * one output row (M = 128 pixels, or 64 as an M = 64 MMA) = 15 (8 ch) / 50 (32 ch) MMAs into a TMEM
//     accumulator stage; rows of the image above / below the border are simply skipped.
//
// Epilogue (4 warps, one thread per pixel): + bias, per-sample sum / sum of squares (InstanceNorm
// statistics) or the fused InstanceNorm-backward pass 1 (norm_bwd.cuh) with the z rows arriving through
// their own TMA ring (swizzled, conflict free), bf16 store.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "norm_bwd.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int RC_THREADS = 352;          // warps 0-7 epilogue (two groups, alternate rows), 8 row producer, 9 MMA issuer, 10 z producer
constexpr int W_PRODUCER = 8, W_MMA = 9, W_ZPROD = 10;
constexpr int RC_ACC = 4;                // TMEM accumulator stages
constexpr int RC_ZSLOTS = 4;              // upper bound; p.zslots (a power of two) are used

struct RcParams {
  int Nimg, Hb, Wb, Hs, Ws, B, A_real;
  int R, strips_per_img, total_strips;   // R output rows per strip
  int ring, row_bytes, plane_bytes;      // ring slots; one input row = S*CH planes of Wp x 16 B
  int ktot;                              // K extent of the weight operand: 5 * NMK * 16
  int w_bytes, z_slot_bytes, zslots;     // one z slot = the 128 pixels of a TMEM stage
  const uint4* wpack;                    // bf16 weight operand in its shared-memory layout (rc_pack_kernel)
  const float* bias;
  bf16* out;                             // [N,Hs,Ws,B]
  double* stats;
};

// MMAs per kernel row (ky): 8-channel pixels pair two taps per K = 16, 32-channel pixels need 2 per tap
template <int CH> struct NMK { static constexpr int v = CH == 1 ? 3 : 10; };

// Low descriptor half (minus the slot address) of the A operand of MMA `p` of a kernel row:
// (byte offset inside the ring slot) >> 4  |  (LBO >> 4) << 16.
template <int S, int CH, int PLANEB>
__host__ __device__ constexpr uint32_t a_lo_const(int p) {
  if (CH == 1) {
    const uint32_t off = (S == 1) ? (uint32_t)(2 * p) * 16u                    // taps (2p, 2p+1): pixel j + kx
                       : (p == 0) ? (uint32_t)PLANEB                           // taps (0,2): odd-x plane, pixel j
                       : (p == 1) ? 16u                                        // taps (1,3): even-x plane, j+1
                                  : (uint32_t)PLANEB + 32u;                    // tap 4 (+ phantom): odd-x, j+2
    return (off >> 4) | ((16u >> 4) << 16);                                    // K halves = neighbouring taps
  } else {                                                                     // S == 2, 32 channels
    const int kx = p >> 1, cp = p & 1;
    const int xpar = (kx & 1) ? 0 : 1;
    const int d = (kx + 1) >> 1;                                               // 0,1,1,2,2
    // 32-channel pixels are 64-byte rows of a SWIZZLE_64B K-major operand: x-parity plane, pixel shift, K step
    const uint32_t off = (uint32_t)(xpar * CH * PLANEB + d * 64 + cp * 32);
    return (off >> 4) | (1u << 16);                                            // LBO unused
  }
}

// Which weight element feeds K index kk (0..15) of MMA g = ky*NMK + p?  Returns the row of W[25*A][B] or -1.
template <int S, int CH>
__device__ __forceinline__ int w_row(int g, int kk, int A_real) {
  const int ky = g / NMK<CH>::v, p = g - ky * NMK<CH>::v;
  if (CH == 1) {
    const int half = kk >> 3, c = kk & 7;
    int kx;
    if (S == 1) kx = 2 * p + half;
    else kx = (p == 0) ? 2 * half : (p == 1) ? 1 + 2 * half : (half == 0 ? 4 : 5);
    if (kx > 4 || c >= A_real) return -1;
    return (ky * 5 + kx) * A_real + c;
  } else {
    const int kx = p >> 1, cp = p & 1;
    return (ky * 5 + kx) * A_real + 16 * cp + kk;
  }
}

// fp32 W[(tap,a)][b] -> bf16 [b][k] canonical no-swizzle K-major operand (8 x 16 B cores, LBO = 128 B, SBO =
// ktot/8 cores), k = g*16 + kk: exactly the bytes the row kernel keeps in shared memory.
template <int S, int CH>
__global__ void rc_pack_kernel(const float* __restrict__ W, bf16* __restrict__ out, int A_real, int B, int ktot) {
  const int kcores = ktot >> 3;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ktot * B; e += gridDim.x * blockDim.x) {
    const int k = e / B, b = e - k * B;
    const int r = w_row<S, CH>(k >> 4, k & 15, A_real);
    const float v = r >= 0 ? W[(int64_t)r * B + b] : 0.f;
    out[(b >> 3) * (kcores * 64) + (k >> 3) * 64 + (b & 7) * 8 + (k & 7)] = __float2bfloat16_rn(v);
  }
}

// S: conv stride.  CH: 8-channel chunks per input pixel.  MM: MMA M (128, or 64 for 64-pixel rows).
template <int S, int CH, int MM, bool NB>
__global__ void __launch_bounds__(RC_THREADS)
tc_rowconv_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmZ, const RcParams p,
                  const NormBwdDev nb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sZ = smem;                                             // RC_ZSLOTS x z_slot_bytes (swizzled: 1024-aligned)
  uint8_t* sW = sZ + (NB ? p.zslots * p.z_slot_bytes : 0);       // B rows x ktot, no-swizzle K-major
  uint8_t* sRing = sW + ((p.w_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + (size_t)p.ring * p.row_bytes);
  uint64_t* full = bars;                     // [ring]   TMA -> MMA
  uint64_t* empty = full + p.ring;           // [ring]   MMA -> TMA
  uint64_t* tfull = empty + p.ring;          // [RC_ACC] MMA -> epilogue
  uint64_t* tempty = tfull + RC_ACC;         // [RC_ACC] epilogue -> MMA (4 arrivals)
  uint64_t* zfull = tempty + RC_ACC;         // [RC_ZSLOTS]
  uint64_t* zempty = zfull + RC_ZSLOTS;      // [RC_ZSLOTS] (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zempty + RC_ZSLOTS);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);        // B floats (16-byte aligned)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NM = NMK<CH>::v;
  constexpr int PAD = (S == 2) ? 1 : 2;
  constexpr int PLANEB = (128 / S + 8) * 16;     // one plane: (Ws + 8) pixels x 16 B
  constexpr int ROWB = S * CH * PLANEB;          // one input row in the ring
  // Output rows per TMEM stage: an M = 64 accumulator occupies lanes 0-15 of every 32-lane quadrant, or lanes
  // 16-31 when the TMEM address carries a lane offset of 16 - two 64-pixel rows share one stage, so the
  // epilogue's 32-lane tcgen05.ld returns 128 useful pixels exactly as in the M = 128 case.
  constexpr int RPS = (MM == 64) ? 2 : 1;
  const uint32_t tmem_cols = RC_ACC * p.B <= 128 ? 128 : 256;

  // weights: already bf16 in the canonical no-swizzle K-major operand layout (rc_pack_kernel)
  // (four independent 16-byte loads in flight per thread before the stores)
  for (int e0 = 0; e0 < (p.w_bytes >> 4); e0 += 4 * RC_THREADS) {
    uint4 wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * RC_THREADS + (int)threadIdx.x;
      wv[i] = e < (p.w_bytes >> 4) ? __ldg(p.wpack + e) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * RC_THREADS + (int)threadIdx.x;
      if (e < (p.w_bytes >> 4)) reinterpret_cast<uint4*>(sW)[e] = wv[i];
    }
  }
  for (int e = threadIdx.x; e < p.B; e += RC_THREADS) sbias[e] = p.bias ? p.bias[e] : 0.f;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmIn);
    if (NB) tc::tma_prefetch_desc(&tmZ);
    for (int i = 0; i < p.ring; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < RC_ACC; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    for (int i = 0; i < p.zslots; ++i) { tc::mbar_init(&zfull[i], 1); tc::mbar_init(&zempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_PRODUCER) {
    // ------------------------------------------------------------ input-row producer
    // Every strip occupies S*(R-1)+5 consecutive ring slots, rows outside the image included (no load, a
    // plain arrive): the slot of a row is then a pure function of its sequence number.
    if (tc::elect_one()) {
      const int nrows = S * (p.R - 1) + 5;
      const uint32_t ring_addr = tc::smem_u32(sRing), full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      int slot = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, i0 = (t - n * p.strips_per_img) * p.R;
        const int y0 = S * i0 - PAD, row0 = n * p.Hb;
        for (int k = 0; k < nrows; ++k) {
          const int y = y0 + k;
          tc::mbar_wait_addr(empty_addr + slot * 8, phase ^ 1);
          if (y >= 0 && y < p.Hb) {
            tc::mbar_expect_tx_addr(full_addr + slot * 8, (uint32_t)ROWB);
            tc::tma_load_4d_addr(ring_addr + slot * ROWB, &tmIn, full_addr + slot * 8, 0, -PAD, 0, row0 + y);
          } else {
            tc::mbar_arrive(&full[slot]);
          }
          if (++slot == p.ring) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_ZPROD) {
    // ------------------------------------------------------------ z producer (norm-backward epilogue):
    // the 128 pixels (RPS consecutive rows) of every TMEM stage
    if (NB && tc::elect_one()) {
      int zs = 0; uint32_t zphase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, i0 = (t - n * p.strips_per_img) * p.R;
        for (int i = i0; i < i0 + p.R; i += RPS) {
          tc::mbar_wait(&zempty[zs], zphase ^ 1);
          tc::mbar_expect_tx(&zfull[zs], (uint32_t)p.z_slot_bytes);
          tc::tma_load_2d(sZ + (size_t)zs * p.z_slot_bytes, &tmZ, &zfull[zs], 0, (n * p.Hs + i) * p.Ws);
          if (++zs == p.zslots) { zs = 0; zphase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------ MMA issuer
    // Lean single-thread loop: descriptor high halves are invariant, low halves are (slot address >> 4) plus
    // compile-time constants; ring slots advance by compare-and-subtract.
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(MM, p.B, 0, 0);
      const uint32_t ring_lo = tc::smem_u32(sRing) >> 4;
      const uint32_t full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      const uint32_t tfull_addr = tc::smem_u32(tfull), tempty_addr = tc::smem_u32(tempty);
      const uint32_t a_hi = CH == 1 ? ((128u >> 4) | (1u << 14))                     // no swizzle: SBO = 128 B, version 1
                                    : ((512u >> 4) | (1u << 14) | (4u << 29));       // SWIZZLE_64B: SBO = 8 x 64-byte rows
      const uint32_t b_hi = (((uint32_t)(p.ktot >> 3) * 128u) >> 4) | (1u << 14);   // SBO = ktot/8 cores
      const uint32_t b_lo0 = (tc::smem_u32(sW) >> 4) | ((128u >> 4) << 16);         // LBO = 128 B
      const int ring = p.ring, R = p.R;
      int s0 = 0;                          // ring slot of the strip's current top row (ky = 0)
      int wslot = 0; uint32_t wphase = 0;  // next row to wait for
      int acc = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, i0 = (t - n * p.strips_per_img) * R;
        (void)n;
        for (int r = 0; r < R; r += RPS) {
          tc::mbar_wait_addr(tempty_addr + acc * 8, aphase ^ 1);
#pragma unroll
          for (int sub = 0; sub < RPS; ++sub) {
            // input rows newly needed by this output row: 5 for the first row of a strip, S afterwards
            const int need = (r + sub == 0) ? 5 : S;
            for (int k = 0; k < need; ++k) {
              tc::mbar_wait_addr(full_addr + wslot * 8, wphase);
              if (++wslot == ring) { wslot = 0; wphase ^= 1; }
            }
            tc::fence_after_sync();
            const uint32_t tacc = tmem_base + (uint32_t)(acc * p.B) + ((uint32_t)(16 * sub) << 16);
            const int ytop = S * (i0 + r + sub) - PAD;
            const bool last = (r + sub == R - 1);
            const int sb = s0;                                // ring slot of this output row's top input row
            uint32_t accum = 0;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
              const int y = ytop + ky;
              if (y >= 0 && y < p.Hb) {
                int sl = sb + ky;
                if (sl >= ring) sl -= ring;
                const uint32_t sa_lo = ring_lo + (uint32_t)sl * (uint32_t)(ROWB >> 4);
#pragma unroll
                for (int q = 0; q < NM; ++q) {
                  tc::mma_bf16_lohi(tacc, sa_lo + a_lo_const<S, CH, PLANEB>(q), a_hi,
                                    b_lo0 + (uint32_t)(ky * NM + q) * 16u, b_hi, idesc, accum);
                  accum = 1;
                }
              }
              // the top S rows are not needed by any later output row: hand them back to the producer as soon
              // as the MMAs reading them retire (everything at the strip's end)
              if (ky == S - 1 && !last) {
#pragma unroll
                for (int k = 0; k < S; ++k) {
                  tc::mma_commit_addr(empty_addr + s0 * 8);
                  if (++s0 == ring) s0 = 0;
                }
              }
            }
            if (last) {
              for (int k = 0; k < 5; ++k) {
                tc::mma_commit_addr(empty_addr + s0 * 8);
                if (++s0 == ring) s0 = 0;
              }
            }
          }
          tc::mma_commit_addr(tfull_addr + acc * 8);
          if (++acc == RC_ACC) { acc = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: one thread per pixel of a stage;
    // warp group 0 takes the even stages of the CTA's stage sequence, group 1 the odd ones
    const int q = warp & 3, grp = warp >> 2;
    const uint32_t sbias_u32 = tc::smem_u32(sbias);
    // pixel of the stage held by this thread's TMEM lane: M = 64 rows 16q..16q+15 of output row `sub`
    const int m = (MM == 64) ? (lane >> 4) * 64 + q * 16 + (lane & 15) : q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int zrow_bytes = p.B * 2;
    const int swz = zrow_bytes == 128 ? (m & 7) : ((m >> 1) & 3);
    const int nstage = p.R / RPS;
    const int zshift = p.zslots == 4 ? 2 : 1;
    uint32_t ps = 0;                                              // stages of this CTA before the current strip
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x, ps += (uint32_t)nstage) {
      const int n = t / p.strips_per_img, i0 = (t - n * p.strips_per_img) * p.R;
      NormBwdCoef coef = {0.f, 0.f, 0.f, 0.f};
      if constexpr (NB) coef = nb_coef(nb, n);
      float s1 = 0.f, s2 = 0.f;
      for (int sg = grp; sg < nstage; sg += 2) {
        const uint32_t seq = ps + (uint32_t)sg;
        const int acc = seq & (RC_ACC - 1), zs = seq & (p.zslots - 1);
        const uint32_t aphase = (seq / RC_ACC) & 1u, zphase = (seq >> zshift) & 1u;
        bf16* orow = p.out + (((int64_t)n * p.Hs + i0 + sg * RPS) * p.Ws + m) * p.B;
        tc::mbar_wait(&tfull[acc], aphase);
        if constexpr (NB) tc::mbar_wait(&zfull[zs], zphase);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + tlane + (uint32_t)(acc * p.B);
        const uint32_t zr_u32 = tc::smem_u32(sZ) + (uint32_t)(zs * p.z_slot_bytes + m * zrow_bytes);
        for (int cb = 0; cb < p.B; cb += 32) {
          float v[32];
          tc::tmem_ld32(taddr + cb, v);
          {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t pk[8];
              float a[16];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 sb = tc::lds_v4(sbias_u32 + 4u * (uint32_t)(cb + 16 * h + e));   // shared window
                a[e] = v[16 * h + e] + sb.x; a[e + 1] = v[16 * h + e + 1] + sb.y;
                a[e + 2] = v[16 * h + e + 2] + sb.z; a[e + 3] = v[16 * h + e + 3] + sb.w;
              }
              if constexpr (NB) {
                const int c0 = (cb >> 3) + 2 * h;                 // logical 16-byte chunk of the z row
                const uint4 z0 = tc::lds_v4u(zr_u32 + (uint32_t)((c0 ^ swz) << 4));          // shared window (was LD.E.128)
                const uint4 z1 = tc::lds_v4u(zr_u32 + (uint32_t)(((c0 + 1) ^ swz) << 4));
                nb_chunk(a, z0, z1, coef, nb.alpha, s1, s2, pk);
              } else {
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  s1 += a[e] + a[e + 1];
                  s2 = fmaf(a[e], a[e], fmaf(a[e + 1], a[e + 1], s2));
                  __nv_bfloat162 hh = __floats2bfloat162_rn(a[e], a[e + 1]);
                  pk[e >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                }
              }
              uint4* dst = reinterpret_cast<uint4*>(orow + cb + 16 * h);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(&tempty[acc]);
          if (NB) tc::mbar_arrive(&zempty[zs]);
        }
      }
      double* sums = NB ? nb.red : p.stats;
      if (sums != nullptr) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { atomicAdd(&sums[2 * n], (double)s1); atomicAdd(&sums[2 * n + 1], (double)s2); }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

struct RcPlan {
  RcParams p;
  int S, CH, Cpad, Wp;
  size_t shm;
  int grid;
};

bool plan_rc(int Nimg, int Hb, int Wb, int A, int Cpad, int B, int s, bool with_nb, RcPlan* pl) {
  if (s != 1 && s != 2) return false;
  if (Wb != 128 || Hb < 8 || (Hb % s)) return false;
  const bool rgb = (Cpad == 8 && A <= 8);
  const bool c32 = (Cpad == 32 && A == 32 && s == 2);
  if (!rgb && !c32) return false;
  if (B != 32 && B != 64) return false;
  RcParams& p = pl->p;
  pl->S = s; pl->CH = Cpad / 8; pl->Cpad = Cpad;
  p.Nimg = Nimg; p.Hb = Hb; p.Wb = Wb; p.Hs = Hb / s; p.Ws = Wb / s; p.B = B; p.A_real = A;
  pl->Wp = p.Ws + 8;
  p.plane_bytes = pl->Wp * 16;
  p.row_bytes = s * pl->CH * p.plane_bytes;
  const int nmk = pl->CH == 1 ? 3 : 10;
  p.ktot = 5 * nmk * 16;
  p.w_bytes = B * p.ktot * 2;
  p.z_slot_bytes = 128 * B * 2;
  p.zslots = B <= 32 ? 4 : 2;
  const int sms = lg_num_sms();
  // ring: at least the 5 + s rows one output row touches, deeper while shared memory allows
  const size_t fixed = (with_nb ? (size_t)p.zslots * p.z_slot_bytes : 0) + ((p.w_bytes + 1023) & ~1023) + 2048 + 1024;
  int ring = 16;
  const size_t budget = (pl->CH == 1) ? 100 * 1024 : 220 * 1024;
  while (ring > 5 + s + 1 && fixed + (size_t)ring * p.row_bytes > budget) --ring;
  if (fixed + (size_t)ring * p.row_bytes > 227 * 1024) return false;
  p.ring = ring;
  pl->shm = fixed + (size_t)ring * p.row_bytes;
  const int ctas = sms * (pl->shm <= 110 * 1024 ? 2 : 1);
  // strip height: balance the persistent waves against the halo rows every strip re-reads
  int bestR = 0; double best = -1.0;
  for (int R = p.Hs; R >= 4; R >>= 1) {
    if (p.Hs % R) continue;
    const int tiles = Nimg * (p.Hs / R);
    const int waves = (tiles + ctas - 1) / ctas;
    // the stride-2 RGB layer (15 MMAs per output row) pays the 5-row fill of every strip four times as dearly as
    // the others (measured, batch 128: R = 4 / 8 / 16 / 32 -> 56.1 / 51.6 / 48.9 / 48.7 us; the 32-channel and the
    // stride-1 layers are fastest at R = 8)
    const double fill = (pl->CH == 1 && s == 2) ? 1.0 : 0.25;
    const double eff = (double)tiles / ((double)waves * ctas) * (1.0 - fill * 4.0 / (s * R + 4.0));
    if (eff > best) { best = eff; bestR = R; }
  }
  if (const char* e = getenv("LG_RC_R")) { const int R = atoi(e); if (R >= 4 && p.Hs % R == 0 && R % 2 == 0) bestR = R; }   // tuning knob
  if (!bestR) return false;
  p.R = bestR; p.strips_per_img = p.Hs / bestR; p.total_strips = Nimg * p.strips_per_img;
  pl->grid = lg_even_grid(p.total_strips, ctas);
  return true;
}

template <int S, int CH, int MM>
void launch_rc(const RcPlan& pl, const CUtensorMap& tmIn, const CUtensorMap& tmZ, const NormBwdDev& nbd, bool with_nb,
               cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_rowconv_kernel<S, CH, MM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_rowconv_kernel<S, CH, MM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  if (with_nb) {
    tc_rowconv_kernel<S, CH, MM, true><<<pl.grid, RC_THREADS, pl.shm, st>>>(tmIn, tmZ, pl.p, nbd);
  } else {
    tc_rowconv_kernel<S, CH, MM, false><<<pl.grid, RC_THREADS, pl.shm, st>>>(tmIn, tmZ, pl.p, nbd);
  }
}

}  // namespace

int lg_tc_rowconv_supported(int Nimg, int Hb, int Wb, int A, int Cpad, int B, int s) {
  RcPlan pl;
  return plan_rc(Nimg, Hb, Wb, A, Cpad, B, s, true, &pl) ? 1 : 0;
}

// big: [N,Hb,Wb,Cpad] bf16 (channels >= A are zero or ignored: their weights are zero).
// Packed weight operand of the row kernel; returns its size in bytes when W or wpack is NULL.
int lg_tc_rowconv_pack(const float* W, void* wpack, int A, int Cpad, int B, int s, cudaStream_t st) {
  RcPlan pl;
  if (!plan_rc(1, 128, 128, A, Cpad, B, s, false, &pl)) {
    lg_set_error("row-streaming conv: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  if (W == nullptr || wpack == nullptr) return pl.p.w_bytes;
  const int total = pl.p.ktot * B, blocks = (total + 255) / 256;
  if (s == 1) rc_pack_kernel<1, 1><<<blocks, 256, 0, st>>>(W, (bf16*)wpack, A, B, pl.p.ktot);
  else if (pl.CH == 1) rc_pack_kernel<2, 1><<<blocks, 256, 0, st>>>(W, (bf16*)wpack, A, B, pl.p.ktot);
  else rc_pack_kernel<2, 4><<<blocks, 256, 0, st>>>(W, (bf16*)wpack, A, B, pl.p.ktot);
  return LG_OK;
}

int lg_tc_rowconv_fprop(const void* big, const void* wpack, const float* bias, void* out, double* stats, int Nimg, int Hb,
                        int Wb, int A, int Cpad, int B, int s, const lg_norm_bwd_t* nb, cudaStream_t st) {
  RcPlan pl;
  if (!plan_rc(Nimg, Hb, Wb, A, Cpad, B, s, nb != nullptr, &pl) || wpack == nullptr) {
    lg_set_error("row-streaming conv: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  RcParams& p = pl.p;
  p.wpack = (const uint4*)wpack; p.bias = bias; p.out = (bf16*)out; p.stats = stats;
  tc_host::EncodeTiledFn enc = tc_host::get_encode();
  if (!enc) { lg_set_error("cuTensorMapEncodeTiled entry point not available"); return LG_ERR_CUDA; }
  CUtensorMap tmIn, tmZ;
  {
    // (channels of a pixel) x (pixels of one x parity) x (parity) x (rows of all images): 16-byte pixels land
    // as no-swizzle planes, 64-byte pixels as SWIZZLE_64B rows (TMA moves one inner run per pixel either way;
    // 16-byte runs of a split 64-byte pixel would quarter its throughput)
    cuuint64_t dims[4] = {(cuuint64_t)Cpad, (cuuint64_t)(Wb / s), (cuuint64_t)s, (cuuint64_t)Nimg * Hb};
    cuuint64_t strides[3] = {(cuuint64_t)s * Cpad * 2, (cuuint64_t)Cpad * 2, (cuuint64_t)Wb * Cpad * 2};
    cuuint32_t box[4] = {(cuuint32_t)Cpad, (cuuint32_t)pl.Wp, (cuuint32_t)s, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(big), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, pl.CH == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lg_set_error("row-streaming conv: input tensor map failed: %d", (int)r); return LG_ERR_CUDA; }
  }
  NormBwdDev nbd = {};
  if (nb) {
    nbd = lg_make_norm_bwd(nb, (int64_t)p.Hs * p.Ws * B);
    cuuint64_t dims[2] = {(cuuint64_t)B, (cuuint64_t)Nimg * p.Hs * p.Ws};
    cuuint64_t strides[1] = {(cuuint64_t)B * 2};
    cuuint32_t box[2] = {(cuuint32_t)B, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(nb->z), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, B == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lg_set_error("row-streaming conv: z tensor map failed: %d", (int)r); return LG_ERR_CUDA; }
  } else {
    tmZ = tmIn;
  }
  const bool wnb = nb != nullptr;
  if (s == 1) launch_rc<1, 1, 128>(pl, tmIn, tmZ, nbd, wnb, st);
  else if (pl.CH == 1) launch_rc<2, 1, 64>(pl, tmIn, tmZ, nbd, wnb, st);
  else launch_rc<2, 4, 64>(pl, tmIn, tmZ, nbd, wnb, st);
  return LG_OK;
}
