// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels (bf16 operands, fp32 accumulate).
//
//   fprop (OP_F): small[m=(n,i,j)][b] = sum_{tap,a} big[n, s*i+ky-pad, s*j+kx-pad, a] * Wf[tap][b][a]
//   dgrad (OP_T): per output phase (py,px):
//                 big[n, s*i+py, s*j+px, a] = sum_{taps of the phase, b} small[n,i+di,j+dj,b] * Wt[tap][a][b]
//
// GEMM view: M = 128 positions of the small map per tile, N = output-channel tile (<= 256), K =
// (tap, channel chunk of KC).  The A operand (activations) is fetched by ONE 4-D TMA box per
// k-block straight from the NHWC tensor: element strides (s,s) implement the conv stride, negative
// / overflowing start coordinates are zero-filled by TMA and implement the SAME padding - there is
// no im2col buffer and no zero insertion.  The B operand (weights) is a 3-D TMA box of the packed
// bf16 kernel.  Both land in 128B/64B/32B-swizzled K-major tiles that tcgen05.mma consumes directly;
// accumulators live in TMEM (double buffered) and are drained by 4 epilogue warps that add the
// bias, accumulate the per-sample InstanceNorm statistics, apply tanh where asked and store bf16.
//
// Persistent: grid = min(#tiles, #SMs); warps 0..7 = epilogue, warp 8 = TMA producer, warp 9 = MMA
// issuer + TMEM owner.  The producer and MMA-issue loops are single-thread and latency bound
// (~450 cycles per mbarrier round trip measured), so one pipeline stage carries `nsub` k-blocks
// (up to 64 KB) and the loops use 32-bit shared addresses and loop-invariant descriptor halves.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "norm_bwd.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int OP_F = 0, OP_T = 1;
constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 320;
// Warp roles: the epilogue owns warps 0-7 (TMEM lane quarter = warp id % 4, column half = warp id / 4: a
// 4-warp epilogue of a 128-column tile lasts about as long as the tile's MMAs and stalls the issuer); the
// two latency-critical single-thread roles get the HIGHEST warp ids because the SM's warp arbiter favours
// high ids.
constexpr int EPI_WARPS = 8;
constexpr int PRODUCER_WARP = 8, MMA_WARP = 9;
constexpr int MAX_STAGES = 8;
constexpr int MAX_SUB = 4;
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcParams {
  int Nimg, Hs, Ws, Hb, Wb, s, pad;
  int Kch;        // contraction channels per tap (A for fprop, B for dgrad)
  int Nch;        // real output channels (B for fprop, A for dgrad)
  int KC;         // channel chunk per k-block (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B, 16 -> SWIZZLE_32B)
  int NT;         // UMMA N (output-channel tile, multiple of 16)
  int n_tiles;    // channel tiles
  int BW, BH, BN; // tile box in small-map coordinates, BW*BH*BN == 128
  int tilesW, tilesH, tilesN;
  int m_tiles, total_tiles, phases;
  int per_phase, lg_nt, lg_tw, lg_th;   // division-free tile decode (all tile counts are powers of two)
  int pair;       // 1: CTA pairs (cta_group::2) - a scheduled tile is two neighbouring M tiles, one per CTA
  int a_bytes, sub_bytes, nsub, stage_bytes, stages;
  int act;
  const float* bias;
  bf16* out;
  double* stats;
};

// Taps of one output phase along one axis, as arithmetic (no tables):
//   dgrad, s=2: phase ph has 2+ph taps, k = (1-ph) + 2i, input shift d = ph - i
//   dgrad, s=1: 5 taps, k = i, d = 2 - i           fprop: 5 taps, k = i, d = i - pad
template <int OP, int S> __device__ __forceinline__ int taps_n(int ph) { return (OP == OP_T && S == 2) ? 2 + ph : 5; }
template <int OP, int S> __device__ __forceinline__ int tap_k(int ph, int i) { return (OP == OP_T && S == 2) ? (1 - ph) + 2 * i : i; }
template <int OP, int S> __device__ __forceinline__ int tap_d(int ph, int i) {
  if (OP == OP_F) return i - (S == 2 ? 1 : 2);
  return (S == 2) ? ph - i : 2 - i;
}

struct TileCoord { int ph_y, ph_x, nt, n0, i0, j0; };

template <bool CTA2>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                    uint32_t idesc, uint32_t accum) {
  if (CTA2) tc::mma_bf16_lohi_2sm(d, a_lo, a_hi, b_lo, b_hi, idesc, accum);
  else tc::mma_bf16_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, accum);
}

template <int OP, int S>
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t, int rank = 0) {
  TileCoord c;
  // heavier phases (more taps) first: q = 0 -> (1,1), 1 -> (1,0), 2 -> (0,1), 3 -> (0,0)
  int q = 0;
  if (OP == OP_T && S == 2) q = (t >= p.per_phase) + (t >= 2 * p.per_phase) + (t >= 3 * p.per_phase);
  const int r = t - q * p.per_phase;
  if (OP == OP_T && S == 2) { c.ph_y = (q < 2) ? 1 : 0; c.ph_x = (q & 1) ? 0 : 1; }
  else { c.ph_y = 0; c.ph_x = 0; }
  c.nt = r & ((1 << p.lg_nt) - 1);
  const int mt0 = ((r >> p.lg_nt) << p.pair) | rank;
  const int tw = mt0 & ((1 << p.lg_tw) - 1);
  const int mt1 = mt0 >> p.lg_tw;
  const int th = mt1 & ((1 << p.lg_th) - 1);
  const int tn = mt1 >> p.lg_th;
  c.n0 = tn * p.BN; c.i0 = th * p.BH; c.j0 = tw * p.BW;
  return c;
}

// NB: fuse the InstanceNorm-backward reduction of the layer below into the epilogue (norm_bwd.cuh).
// CTA2: CTA pairs.  The two CTAs of a cluster own neighbouring M tiles of the same (phase, channel tile);
// the even CTA issues 256-row tcgen05.mma.cta_group::2 instructions over both CTAs' shared memory, so each
// CTA fetches only HALF of the weight tile (the kernels are bound by the L2 -> SM fill rate, not by the
// tensor pipe).  Each CTA drains its own 128 TMEM lanes.
template <int OP, int S, bool NB, bool CTA2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p,
               const NormBwdDev nb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;                       // [stages]   TMA -> MMA
  uint64_t* empty = bars + MAX_STAGES;         // [stages]   MMA -> TMA
  uint64_t* tfull = bars + 2 * MAX_STAGES;     // [2]        MMA -> epilogue
  uint64_t* tempty = bars + 2 * MAX_STAGES + 2;// [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  float* sbias = reinterpret_cast<float*>(bars + 32);     // n_tiles * NT floats (<= 512)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CTA2 ? (int)tc::cluster_ctarank() : 0;
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t tmem_cols = (2 * p.NT <= 32) ? 32 : (2 * p.NT <= 64) ? 64 : (2 * p.NT <= 128) ? 128
                             : (2 * p.NT <= 256) ? 256 : 512;

  if (warp == PRODUCER_WARP && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], CTA2 ? 2 * EPI_WARPS : EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    if (CTA2) tc::tmem_alloc_2sm(tmem_slot, tmem_cols);
    else tc::tmem_alloc(tmem_slot, tmem_cols);
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.NT; i += NUM_THREADS) sbias[i] = (p.bias && i < p.Nch) ? p.bias[i] : 0.f;
  tc::fence_before_sync();
  if (CTA2) tc::cluster_sync_all(); else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int kc_per_tap = p.Kch / p.KC;
  const uint32_t smem_u32 = tc::smem_u32(smem);
  const uint32_t full_u32 = tc::smem_u32(full), empty_u32 = tc::smem_u32(empty);
  const uint32_t stage_bytes_u = (uint32_t)p.stage_bytes, sub_bytes_u = (uint32_t)p.sub_bytes;
  const uint32_t a_bytes_u = (uint32_t)p.a_bytes;
  const int nstages = p.stages, KCc = p.KC, nsub = p.nsub;

  if (warp == PRODUCER_WARP) {
    // ===================================== TMA producer =====================================
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      // pairs: both CTAs' loads complete on the LEADER's full barrier, which the leader arms for both
      const uint32_t full_tma = CTA2 ? tc::mapa(full_u32, 0) : full_u32;
      for (int t = tile0; t < p.total_tiles; t += tstep) {
        const TileCoord c = decode_tile<OP, S>(p, t, rank);
        const int nty = taps_n<OP, S>(c.ph_y), ntx = taps_n<OP, S>(c.ph_x);
        const int num_kb = nty * ntx * kc_per_tap;
        const int bx = (OP == OP_F) ? S * c.j0 : c.j0, by = (OP == OP_F) ? S * c.i0 : c.i0;
        const int n_off = c.nt * p.NT + (CTA2 ? rank * (p.NT >> 1) : 0);
        int iy = 0, ix = 0, kc = 0;
        for (int kb0 = 0; kb0 < num_kb; kb0 += nsub) {
          const int nvalid = min(nsub, num_kb - kb0);
          const uint32_t fb = full_tma + (uint32_t)stage * 8u;
          uint32_t sa = smem_u32 + (uint32_t)stage * stage_bytes_u;
          tc::mbar_wait_addr(empty_u32 + (uint32_t)stage * 8u, phase ^ 1);
          if (!CTA2) tc::mbar_expect_tx_addr(fb, (uint32_t)nvalid * sub_bytes_u);
          else if (rank == 0) tc::mbar_expect_tx_addr(full_u32 + (uint32_t)stage * 8u, 2u * (uint32_t)nvalid * sub_bytes_u);
#pragma unroll
          for (int j = 0; j < MAX_SUB; ++j) {
            if (j < nvalid) {
              const int tap = tap_k<OP, S>(c.ph_y, iy) * 5 + tap_k<OP, S>(c.ph_x, ix);
              if (CTA2) {
                tc::tma_load_4d_2sm(sa, &tmA, fb, kc * KCc, bx + tap_d<OP, S>(c.ph_x, ix),
                                    by + tap_d<OP, S>(c.ph_y, iy), c.n0);
                tc::tma_load_3d_2sm(sa + a_bytes_u, &tmB, fb, kc * KCc, n_off, tap);
              } else {
                tc::tma_load_4d_addr(sa, &tmA, fb, kc * KCc, bx + tap_d<OP, S>(c.ph_x, ix),
                                     by + tap_d<OP, S>(c.ph_y, iy), c.n0);
                tc::tma_load_3d_addr(sa + a_bytes_u, &tmB, fb, kc * KCc, n_off, tap);
              }
              sa += sub_bytes_u;
              if (++kc == kc_per_tap) { kc = 0; if (++ix == ntx) { ix = 0; ++iy; } }
            }
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA issuer =======================================
    if ((!CTA2 || rank == 0) && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(CTA2 ? 2 * TILE_M : TILE_M, p.NT, 0, 0);
      const uint32_t layout = (KCc == 64) ? 2u : (KCc == 32) ? 4u : 6u;     // SWIZZLE_128B / 64B / 32B
      const uint32_t sbo = 8u * (uint32_t)KCc * 2u;                        // 8 rows of KC bf16
      // K-major swizzled descriptor: lo = start>>4 | LBO(16 B)<<16 ; hi = SBO>>4 | version<<14 | layout<<29
      const uint32_t desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      const uint32_t desc_lo0 = ((smem_u32 & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t stage_units = stage_bytes_u >> 4, sub_units = sub_bytes_u >> 4, a_units = a_bytes_u >> 4;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = tile0; t < p.total_tiles; t += tstep) {
        const TileCoord c = decode_tile<OP, S>(p, t);
        const int num_kb = taps_n<OP, S>(c.ph_y) * taps_n<OP, S>(c.ph_x) * kc_per_tap;
        tc::mbar_wait_addr(tc::smem_u32(&tempty[acc]), acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.NT);
        uint32_t accum = 0;
        for (int kb0 = 0; kb0 < num_kb; kb0 += nsub) {
          const int nvalid = min(nsub, num_kb - kb0);
          tc::mbar_wait_addr(full_u32 + (uint32_t)stage * 8u, phase);
          tc::fence_after_sync();
          uint32_t a_lo = desc_lo0 + (uint32_t)stage * stage_units;
#pragma unroll
          for (int j = 0; j < MAX_SUB; ++j) {
            if (j < nvalid) {
              const uint32_t b_lo = a_lo + a_units;
              if (KCc == 64) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { mma<CTA2>(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, accum); accum = 1; }
              } else if (KCc == 32) {
#pragma unroll
                for (int k = 0; k < 2; ++k) { mma<CTA2>(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, accum); accum = 1; }
              } else {
                mma<CTA2>(d_tmem, a_lo, desc_hi, b_lo, desc_hi, idesc, accum); accum = 1;
              }
              a_lo += sub_units;
            }
          }
          // frees the smem slot (in both CTAs of a pair) when the MMAs retire
          if (CTA2) tc::mma_commit_2sm(empty_u32 + (uint32_t)stage * 8u);
          else tc::mma_commit_addr(empty_u32 + (uint32_t)stage * 8u);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (CTA2) tc::mma_commit_2sm(tc::smem_u32(&tfull[acc]));   // accumulator complete -> both epilogues
        else tc::mma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue ==========================================
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                       // row of the tile == TMEM lane
    const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
    const bool vec_ok = (p.Nch & 7) == 0;
    const uint32_t sbias_u32 = tc::smem_u32(sbias);
    // column range of this warp: half of the tile when that is a whole number of 16-column chunks
    const int half = warp >> 2;
    const bool split = (p.NT & 31) == 0;
    const int ncol = split ? (p.NT >> 1) : (half == 0 ? p.NT : 0);
    const int col0 = split ? half * ncol : 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t tempty_leader = CTA2 ? tc::mapa(tc::smem_u32(tempty), 0) : 0u;
    for (int t = tile0; t < p.total_tiles; t += tstep) {
      const TileCoord c = decode_tile<OP, S>(p, t, rank);
      const int n = c.n0 + bn, i = c.i0 + bh, j = c.j0 + bw;
      const bool valid = n < p.Nimg;
      int64_t off;
      if (OP == OP_F) off = (((int64_t)n * p.Hs + i) * p.Ws + j) * p.Nch;
      else off = (((int64_t)n * p.Hb + (S * i + c.ph_y)) * p.Wb + (S * j + c.ph_x)) * p.Nch;
      bf16* orow = p.out + off;
      const int ch0 = c.nt * p.NT;

      NormBwdCoef coef = {0.f, 0.f, 0.f, 0.f};
      NormBwdZ zc;
      const bf16* zrow = nullptr;
      if constexpr (NB) {
        // z row of this output position: first 64 channels into registers, the rest of the row towards L2,
        // all in flight while the MMAs of this tile still run
        zrow = nb.z + off + ch0 + col0;
        if (valid) coef = nb_coef(nb, n);
        nb_load(zc, zrow, ncol >> 4, valid);
        if (ncol > 64) nb_prefetch_l2(zrow + 64, (ncol - 64) * 2, valid);
        // the NEXT tile's z row towards L2 now: a whole tile period ahead of its use (the loads above miss
        // to HBM only for the first tile of a CTA)
        const int tn = t + tstep;
        if (tn < p.total_tiles) {
          const TileCoord cn = decode_tile<OP, S>(p, tn, rank);
          const int n2 = cn.n0 + bn, i2 = cn.i0 + bh, j2 = cn.j0 + bw;
          int64_t off2;
          if (OP == OP_F) off2 = (((int64_t)n2 * p.Hs + i2) * p.Ws + j2) * p.Nch;
          else off2 = (((int64_t)n2 * p.Hb + (S * i2 + cn.ph_y)) * p.Wb + (S * j2 + cn.ph_x)) * p.Nch;
          nb_prefetch_l2(nb.z + off2 + cn.nt * p.NT + col0, ncol * 2, n2 < p.Nimg);
        }
      }
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      float s1 = 0.f, s2 = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      if constexpr (NB) {
        // s1 / s2 carry (sum dy, sum dy*xhat); the host guarantees Nch % 16 == 0, no activation
        for (int sc = 0; sc < ncol; sc += 64) {
          NormBwdZ zn;
          if (sc + 64 < ncol) nb_load(zn, zrow + sc + 64, (ncol - sc - 64) >> 4, valid);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cb = col0 + sc + 16 * c;
            if (cb < col0 + ncol) {
              float v[16];
              tc::tmem_ld16(taddr + cb, v);
              tc::add_bias16(v, sbias_u32, ch0 + cb);
              uint32_t pk[8];
              nb_chunk(v, zc.v[2 * c], zc.v[2 * c + 1], coef, nb.alpha, s1, s2, pk);
              if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(orow + ch0 + cb);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
          if (sc + 64 < ncol) zc = zn;
        }
      } else {
      for (int cb = col0; cb < col0 + ncol; cb += 16) {
        float v[16];
        tc::tmem_ld16(taddr + cb, v);
        const int chb = ch0 + cb;
        if (chb >= p.Nch) continue;                       // channel padding (uniform over the CTA)
        tc::add_bias16(v, sbias_u32, chb);
        if (chb + 16 <= p.Nch && vec_ok) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float a = v[e];
            float b = v[e + 1];
            s1 += a + b; s2 += a * a + b * b;
            if (p.act == LG_ACT_TANH) { a = tanhf(a); b = tanhf(b); }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + chb);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        } else {
          // channel tail (Nch not a multiple of 16 / 8): fully unrolled with constant indices - a run-time index
          // into v[] forces the whole array into LOCAL memory (4 STL.128 per chunk and thread on every path)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (chb + e < p.Nch) {
              float a = v[e];
              s1 += a; s2 += a * a;
              if (p.act == LG_ACT_TANH) a = tanhf(a);
              if (valid) orow[chb + e] = __float2bfloat16_rn(a);
            }
          }
        }
      }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld inside tmem_ld16)
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) tc::mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        else tc::mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      double* sums = NB ? nb.red : p.stats;
      if (sums != nullptr) {
        // rows of one warp belong to one sample whenever BW*BH >= 32 (checked on the host)
        if (!valid) { s1 = 0.f; s2 = 0.f; }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0 && valid) {
          atomicAdd(&sums[2 * n], (double)s1);
          atomicAdd(&sums[2 * n + 1], (double)s2);
        }
      }
    }
  }

  tc::fence_before_sync();
  if (CTA2) tc::cluster_sync_all(); else __syncthreads();
  if (warp == MMA_WARP) {
    tc::fence_after_sync();
    if (CTA2) tc::tmem_dealloc_2sm(tmem_base, tmem_cols);
    else tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using tc_host::EncodeTiledFn;
using tc_host::get_encode;
using tc_host::is_pow2;
using tc_host::encode_act_map;

// Geometry shared by the support query and the launcher.  Returns false if not covered.
bool plan(int op, int Nimg, int Hb, int Wb, int A, int B, int s, TcParams* p, int pair_mode = 0) {
  if (s != 1 && s != 2) return false;
  const int Hs = Hb / s, Ws = Wb / s;
  if (!is_pow2(Hs) || !is_pow2(Ws) || Ws > 128 || Hs * Ws < 32) return false;
  const int Kch = (op == OP_F) ? A : B;
  const int Nch = (op == OP_F) ? B : A;
  if (Kch % 16 != 0) return false;
  const int Npad = (Nch + 15) / 16 * 16;
  int n_tiles = (Npad + 255) / 256;
  if (Npad % n_tiles != 0 || (Npad / n_tiles) % 16 != 0) return false;
  p->Nimg = Nimg; p->Hs = Hs; p->Ws = Ws; p->Hb = Hb; p->Wb = Wb; p->s = s; p->pad = (s == 2) ? 1 : 2;
  p->Kch = Kch; p->Nch = Nch; p->KC = (Kch % 64 == 0) ? 64 : (Kch % 32 == 0) ? 32 : 16;
  p->NT = Npad / n_tiles; p->n_tiles = n_tiles;
  p->BW = Ws < 128 ? Ws : 128;
  p->BH = (128 / p->BW) < Hs ? (128 / p->BW) : Hs;
  p->BN = 128 / (p->BW * p->BH);
  if (p->BW * p->BH * p->BN != 128 || p->BW * p->BH < 32) return false;
  p->tilesW = Ws / p->BW; p->tilesH = Hs / p->BH; p->tilesN = (Nimg + p->BN - 1) / p->BN;
  p->m_tiles = p->tilesW * p->tilesH * p->tilesN;
  p->phases = (op == OP_T) ? s * s : 1;
  // CTA pairs need an even number of M tiles and enough pair tiles to fill the 74 TPCs
  p->pair = (pair_mode > 0 && p->m_tiles % 2 == 0 && p->NT % 16 == 0 &&
             (pair_mode == 2 || (p->m_tiles / 2) * n_tiles * p->phases >= lg_num_sms() / 2)) ? 1 : 0;
  p->per_phase = (p->m_tiles >> p->pair) * p->n_tiles;
  p->total_tiles = p->per_phase * p->phases;
  auto lg2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  if (!is_pow2(p->n_tiles) || !is_pow2(p->tilesW) || !is_pow2(p->tilesH)) return false;
  p->lg_nt = lg2(p->n_tiles); p->lg_tw = lg2(p->tilesW); p->lg_th = lg2(p->tilesH);
  p->a_bytes = TILE_M * p->KC * 2;
  p->sub_bytes = p->a_bytes + (p->NT >> p->pair) * p->KC * 2;
  // k-blocks per pipeline stage: amortise the ~450-cycle mbarrier round trip over >= ~512 MMA cycles
  int nsub = 65536 / p->sub_bytes;
  if (nsub > MAX_SUB) nsub = MAX_SUB;
  if (nsub < 1) nsub = 1;
  while (nsub > 1 && SMEM_BUDGET / (nsub * p->sub_bytes) < 3) --nsub;
  p->nsub = nsub;
  p->stage_bytes = nsub * p->sub_bytes;
  int st = SMEM_BUDGET / p->stage_bytes;
  p->stages = st > MAX_STAGES ? MAX_STAGES : st;
  if (p->stages < 2) return false;
  if (op == OP_F && (s * p->BW > 256 || s * p->BH > 256)) return false;
  return true;
}

size_t smem_bytes(const TcParams& p) { return (size_t)p.stages * p.stage_bytes + 1024 + 256 + 2048; }

// packed weights [25][rows][cols] bf16, cols contiguous (the contraction channels)
int encode_w_map(CUtensorMap* m, const void* base, int rows, int cols, int boxCols, int boxRows,
                 CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { lg_set_error("cuTensorMapEncodeTiled entry point not available"); return LG_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 25};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  cuuint32_t box[3] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { lg_set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return LG_ERR_CUDA; }
  return LG_OK;
}

template <int OP, int S, bool NB>
void launch_kernel(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, const NormBwdDev& nb,
                   cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_conv_kernel<OP, S, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_conv_kernel<OP, S, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  if (p.pair) {
    // one CTA pair (cluster of 2 = the two SMs of a TPC) per pair tile, persistent over the 74 TPCs
    const int npairs = lg_even_grid(p.total_tiles, lg_num_sms() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * npairs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes(p);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, tc_conv_kernel<OP, S, NB, true>, tmA, tmB, p, nb);
    return;
  }
  const int grid = lg_even_grid(p.total_tiles, lg_num_sms());
  tc_conv_kernel<OP, S, NB, false><<<grid, NUM_THREADS, smem_bytes(p), st>>>(tmA, tmB, p, nb);
}

// 0 = single-CTA tiles only, 1 = pairs when the launch fills every TPC, 2 = pairs whenever possible
int g_pair_mode = -1;
int pair_mode() {
  if (g_pair_mode < 0) {
    const char* e = getenv("LG_TC_PAIRS");
    g_pair_mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return g_pair_mode;
}

template <int OP>
int launch_tc(const void* act_in, const void* wpack, const float* bias, void* out, double* stats, int Nimg, int Hb,
              int Wb, int A, int B, int s, int act, const lg_norm_bwd_t* nbh, cudaStream_t st) {
  TcParams p;
  if (!plan(OP, Nimg, Hb, Wb, A, B, s, &p, pair_mode())) {
    lg_set_error("tcgen05 path: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  NormBwdDev nb = {};
  if (nbh != nullptr) {
    if (p.Nch % 16 != 0 || act != LG_ACT_NONE || stats != nullptr) {
      lg_set_error("tcgen05 path: fused norm-backward needs Nch %% 16 == 0, no activation, no forward statistics");
      return LG_ERR_UNSUPPORTED;
    }
    nb = lg_make_norm_bwd(nbh, OP == OP_F ? (int64_t)p.Hs * p.Ws * p.Nch : (int64_t)p.Hb * p.Wb * p.Nch);
  }
  p.act = act; p.bias = bias; p.out = (bf16*)out; p.stats = stats;
  const int Ap = (A + 15) / 16 * 16, Bp = (B + 15) / 16 * 16;
  const bf16* wt = (const bf16*)wpack;                  // [25][Ap][Bp]
  const bf16* wf = wt + (size_t)25 * Ap * Bp;           // [25][Bp][Ap]
  const CUtensorMapSwizzle sw = p.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : p.KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmA, tmB;
  int e;
  if (OP == OP_F) {
    e = encode_act_map(&tmA, act_in, Nimg, Hb, Wb, A, p.KC, p.BW, p.BH, p.BN, s, sw);
    if (e) return e;
    e = encode_w_map(&tmB, wf, Bp, Ap, p.KC, p.NT >> p.pair, sw);
  } else {
    e = encode_act_map(&tmA, act_in, Nimg, p.Hs, p.Ws, B, p.KC, p.BW, p.BH, p.BN, 1, sw);
    if (e) return e;
    e = encode_w_map(&tmB, wt, Ap, Bp, p.KC, p.NT >> p.pair, sw);
  }
  if (e) return e;
  if (nbh != nullptr) {
    if (s == 1) launch_kernel<OP, 1, true>(tmA, tmB, p, nb, st);
    else launch_kernel<OP, 2, true>(tmA, tmB, p, nb, st);
  } else {
    if (s == 1) launch_kernel<OP, 1, false>(tmA, tmB, p, nb, st);
    else launch_kernel<OP, 2, false>(tmA, tmB, p, nb, st);
  }
  return LG_OK;
}

}  // namespace

int lg_tc_wgrad_supported(int Nimg, int Hb, int Wb, int A, int B, int s);   // tc_wgrad.cu

extern "C" int lg_set_cta_pairs(int mode) {
  const int prev = pair_mode();
  if (mode >= 0 && mode <= 2) g_pair_mode = mode;
  return prev;
}

int lg_tc_supported(int op, int Hb, int Wb, int A, int B, int s, int N) {
  if (op != LG_OP_DGRAD && lg_tc_cin3_supported(N, Hb, Wb, A, B, s)) return 1;
  if (op == LG_OP_WGRAD) return lg_tc_wgrad_supported(N, Hb, Wb, A, B, s);
  if (op == LG_OP_DGRAD && lg_tc_deconv_small_supported(N, Hb, Wb, A, B, s)) return 1;
  if (op == LG_OP_DGRAD && lg_tc_dgrad4_supported(N, Hb, Wb, A, B, s)) return 1;
  TcParams p;
  return plan(op == LG_OP_FPROP ? OP_F : OP_T, N, Hb, Wb, A, B, s, &p) ? 1 : 0;
}

int lg_tc_fprop(const void* big, const void* wpack, const float* bias, void* out, double* stats, int N, int Hb,
                int Wb, int A, int B, int s, const lg_norm_bwd_t* nb, cudaStream_t st) {
  return launch_tc<OP_F>(big, wpack, bias, out, stats, N, Hb, Wb, A, B, s, LG_ACT_NONE, nb, st);
}

int lg_tc_dgrad(const void* small, const void* wpack, const float* bias, void* out, double* stats, int N, int Hb,
                int Wb, int A, int B, int s, int act, const lg_norm_bwd_t* nb, cudaStream_t st) {
  return launch_tc<OP_T>(small, wpack, bias, out, stats, N, Hb, Wb, A, B, s, act, nb, st);
}
