// Stride-2 transposed convolution (conv dgrad) with <= 128 output channels, all FOUR output phases of
// a tile in one pass (tcgen05, bf16 operands, fp32 accumulate):
//
//   big[n, 2i+py, 2j+px, a] = sum_{taps of phase (py,px), b} small[n, i+di, j+dj, b] * Wt[tap][a][b]
//
// The 25 taps read the input at only 9 distinct shifts (di,dj) in {+1,0,-1}^2: shift +1 serves tap 0,
// shift 0 taps {1,2}, shift -1 taps {3,4} (per axis), and tap k belongs to phase (k+1)%2.  The generic
// kernel (tc_conv.cu) runs the phases as separate tiles and fetches the activation tile once per TAP; here
// a pipeline stage is one activation tile (one shift, one 64-channel chunk) plus the 1/2/4 weight tiles of
// the taps that use it, and the MMAs of those taps accumulate into four TMEM accumulators (one per phase).
// Shared-memory traffic per tile drops from 25 to 9 activation tiles and the tile count by 4x, which is
// what bounds these layers (operand bytes through smem, not MMA issue).
//
// Round 2: a pipeline stage is one SHIFT with up to two channel chunks (9 stages per tile for 128 input channels
// instead of 18: the kernel was bound by the latency of its stage hand-shakes - knock-outs in profiles/r2_ubench.txt -
// not by bandwidth or MMA issue), its boxes are issued by THREE producer threads (one thread gets a TMA instruction
// out only every ~60-100 ns), and the taps of a shift - which feed DIFFERENT output phases - are ONE tcgen05.mma whose
// N spans their accumulators (9 instead of 25 MMAs per k-step; an MMA re-reads its 128-row A operand from shared
// memory whatever N is).
//
// Warp roles: warps 0-3 epilogue, warp 5 MMA issuer + TMEM owner, warp 4 activation-tile producer, warps 6 / 7
// weight-tile producers (even / odd tile slot).
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "internal.h"
#include "norm_bwd.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int NUM_THREADS = 256;
constexpr int PRODUCER_WARP = 4, MMA_WARP = 5, WPROD_WARP0 = 6;
constexpr int NUM_PRODUCERS = 3;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 200 * 1024;

struct D4Params {
  int Nimg, Hs, Ws, Hb, Wb;
  int Kch, Nch, KC, NT;          // contraction channels B, output channels A, chunk, padded A
  int BW, BH, BN, lg_tw, lg_th;  // 128-position tile box in small-map coordinates
  int total_tiles;
  int a_bytes, b_bytes, stage_bytes, stages, nacc;   // nacc = 2 (double-buffered accumulator sets) or 1
  int cps;                                           // channel chunks per pipeline stage (1 or 2)
  int act;
  const float* bias;
  bf16* out;
  double* stats;
};

// per-axis shift table: shift index u = 0,1,2 <-> d = +1, 0, -1 ; taps {0}, {1,2}, {3,4}
__device__ __forceinline__ int shift_d(int u) { return 1 - u; }
__device__ __forceinline__ int shift_ntaps(int u) { return u == 0 ? 1 : 2; }
__device__ __forceinline__ int shift_tap(int u, int i) { return u == 0 ? 0 : 2 * u - 1 + i; }

// Accumulator order in TMEM: phases 1, 3, 2, 0 (column = d4_pos(ph) * NT).  That makes the phase set of every shift
// contiguous - {1,3,2,0} for the four 4-tap shifts (N = 4 NT), {1,3} / {3,2} for the 2-tap shifts (N = 2 NT), {3} for
// shift (+1,+1) - so the taps of a shift are one MMA.  The 4-tap shift (0,0) goes first: it writes all four phases,
// so one accumulate flag serves every MMA of the tile.
__device__ __forceinline__ int d4_pos(int ph) { return (0x1203 >> (4 * ph)) & 0xF; }
__device__ __forceinline__ int d4_phase(int ky, int kx) { return ((ky + 1) & 1) * 2 + ((kx + 1) & 1); }
__device__ __forceinline__ int d4_shift(int o) { return o == 0 ? 4 : (o <= 4 ? o - 1 : o); }   // 4,0,1,2,3,5,6,7,8
__device__ __forceinline__ int d4_min_pos(int uy, int ux) {       // first accumulator position of the shift
  int m = 4;
  for (int iy = 0; iy < shift_ntaps(uy); ++iy)
    for (int ix = 0; ix < shift_ntaps(ux); ++ix) {
      const int q = d4_pos(d4_phase(shift_tap(uy, iy), shift_tap(ux, ix)));
      m = q < m ? q : m;
    }
  return m;
}

template <bool NB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_dgrad4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const D4Params p,
                 const NormBwdDev nb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tfull = bars + 2 * MAX_STAGES;
  uint64_t* tempty = bars + 2 * MAX_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  float* sbias = reinterpret_cast<float*>(bars + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int set_cols = 4 * p.NT;                                   // one accumulator set: 4 phases
  const uint32_t need = (uint32_t)(p.nacc * set_cols);
  const uint32_t tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;

  if (warp == PRODUCER_WARP && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full[i], NUM_PRODUCERS); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, tmem_cols);
  for (int i = threadIdx.x; i < p.NT; i += NUM_THREADS) sbias[i] = (p.bias && i < p.Nch) ? p.bias[i] : 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int kc_per_tap = p.Kch / p.KC;
  const uint32_t smem_u32 = tc::smem_u32(smem);
  const uint32_t full_u32 = tc::smem_u32(full), empty_u32 = tc::smem_u32(empty);
  const uint32_t stage_bytes_u = (uint32_t)p.stage_bytes, a_bytes_u = (uint32_t)p.a_bytes, b_bytes_u = (uint32_t)p.b_bytes;
  const int nstages = p.stages, KCc = p.KC, nacc = p.nacc;

  const int cps = p.cps, groups = kc_per_tap / p.cps;          // chunk groups per shift
  const uint32_t chunk_w_bytes = 4u * b_bytes_u;                 // weight-tile area of one chunk: 4 slots
  if (warp == PRODUCER_WARP || warp >= WPROD_WARP0) {
    // role 0: the activation tiles of every stage; role 1 / 2: the weight tiles in even / odd slot (slot = accumulator
    // position relative to the shift's first).  Every producer waits for the stage, announces its own byte count on
    // the full barrier (count 3) and issues its boxes.
    const int role = warp == PRODUCER_WARP ? 0 : 1 + (warp - WPROD_WARP0);
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tw = t & ((1 << p.lg_tw) - 1);
        const int th = (t >> p.lg_tw) & ((1 << p.lg_th) - 1);
        const int n0 = (t >> (p.lg_tw + p.lg_th)) * p.BN;
        const int i0 = th * p.BH, j0 = tw * p.BW;
        for (int o = 0; o < 9; ++o) {
          const int sh = d4_shift(o), uy = sh / 3, ux = sh - 3 * uy;
          const int nyt = shift_ntaps(uy), nxt = shift_ntaps(ux);
          const int pos0 = d4_min_pos(uy, ux);
          int mine = 0;                                           // this producer's weight tiles per chunk
#pragma unroll
          for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int ix = 0; ix < 2; ++ix)
              if (iy < nyt && ix < nxt)
                mine += (((d4_pos(d4_phase(shift_tap(uy, iy), shift_tap(ux, ix))) - pos0) & 1) == role - 1);
          for (int g = 0; g < groups; ++g) {
            const uint32_t fb = full_u32 + (uint32_t)stage * 8u;
            const uint32_t sa = smem_u32 + (uint32_t)stage * stage_bytes_u;
            const uint32_t sw = sa + (uint32_t)cps * a_bytes_u;
            tc::mbar_wait_addr(empty_u32 + (uint32_t)stage * 8u, phase ^ 1);
            if (role == 0) {
              tc::mbar_expect_tx_addr(fb, (uint32_t)cps * a_bytes_u);
              for (int cc = 0; cc < cps; ++cc)
                tc::tma_load_4d_addr(sa + (uint32_t)cc * a_bytes_u, &tmA, fb, (g * cps + cc) * KCc, j0 + shift_d(ux),
                                     i0 + shift_d(uy), n0);
            } else {
              tc::mbar_expect_tx_addr(fb, (uint32_t)(cps * mine) * b_bytes_u);
              for (int cc = 0; cc < cps; ++cc) {
#pragma unroll
                for (int iy = 0; iy < 2; ++iy)
#pragma unroll
                  for (int ix = 0; ix < 2; ++ix)
                    if (iy < nyt && ix < nxt) {
                      const int ky = shift_tap(uy, iy), kx = shift_tap(ux, ix);
                      const int slot = d4_pos(d4_phase(ky, kx)) - pos0;
                      if ((slot & 1) == role - 1)
                        tc::tma_load_3d_addr(sw + (uint32_t)cc * chunk_w_bytes + (uint32_t)slot * b_bytes_u, &tmB, fb,
                                             (g * cps + cc) * KCc, 0, ky * 5 + kx);
                    }
              }
            }
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    if (tc::elect_one()) {
      const uint32_t idesc1 = tc::make_idesc(128, p.NT, 0, 0), idesc2 = tc::make_idesc(128, 2 * p.NT, 0, 0),
                     idesc4 = tc::make_idesc(128, 4 * p.NT, 0, 0);
      const uint32_t layout = (KCc == 64) ? 2u : 4u;
      const uint32_t sbo = 8u * (uint32_t)KCc * 2u;
      const uint32_t desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      const uint32_t desc_lo0 = ((smem_u32 & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t stage_units = stage_bytes_u >> 4, a_units = a_bytes_u >> 4, cw_units = chunk_w_bytes >> 4;
      const int ksteps = KCc / 16;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        tc::mbar_wait_addr(tc::smem_u32(&tempty[acc]), acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_set = tmem_base + (uint32_t)(acc * set_cols);
        uint32_t accum = 0;                                         // the first shift writes all four phases
        for (int o = 0; o < 9; ++o) {
          const int sh = d4_shift(o), uy = sh / 3, ux = sh - 3 * uy;
          const int ntiles = shift_ntaps(uy) * shift_ntaps(ux);
          const uint32_t idesc = ntiles == 4 ? idesc4 : ntiles == 2 ? idesc2 : idesc1;
          const uint32_t d_tmem = d_set + (uint32_t)(d4_min_pos(uy, ux) * p.NT);
          for (int g = 0; g < groups; ++g) {
            tc::mbar_wait_addr(full_u32 + (uint32_t)stage * 8u, phase);
            tc::fence_after_sync();
            const uint32_t a_lo0 = desc_lo0 + (uint32_t)stage * stage_units;
            const uint32_t b_lo0 = a_lo0 + (uint32_t)cps * a_units;
            for (int cc = 0; cc < cps; ++cc) {
              const uint32_t a_lo = a_lo0 + (uint32_t)cc * a_units, b_lo = b_lo0 + (uint32_t)cc * cw_units;
              for (int k = 0; k < ksteps; ++k) {
                tc::mma_bf16_lohi(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, accum);
                accum = 1;
              }
            }
            tc::mma_commit_addr(empty_u32 + (uint32_t)stage * 8u);
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
        tc::mma_commit(&tfull[acc]);
        if (nacc == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
        else acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
    const bool vec_ok = (p.Nch & 7) == 0;
    const uint32_t sbias_u32 = tc::smem_u32(sbias);
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int tw = t & ((1 << p.lg_tw) - 1);
      const int th = (t >> p.lg_tw) & ((1 << p.lg_th) - 1);
      const int n = (t >> (p.lg_tw + p.lg_th)) * p.BN + bn;
      const int i = th * p.BH + bh, j = tw * p.BW + bw;
      const bool valid = n < p.Nimg;
      bf16* obase = p.out + (((int64_t)n * p.Hb + 2 * i) * p.Wb + 2 * j) * p.Nch;

      NormBwdCoef coef = {0.f, 0.f, 0.f, 0.f};
      NormBwdZ zc;
      const bf16* zbase = nullptr;
      if constexpr (NB) {
        // fused InstanceNorm-backward reduction (norm_bwd.cuh): z rows of the four output phases, one phase ahead
        zbase = nb.z + (obase - p.out);
        if (valid) coef = nb_coef(nb, n);
        nb_load(zc, zbase, p.NT >> 4, valid);
        // the NEXT tile's four z rows towards L2, a whole tile period ahead of their use
        const int tnx = t + gridDim.x;
        if (tnx < p.total_tiles) {
          const int tw2 = tnx & ((1 << p.lg_tw) - 1);
          const int th2 = (tnx >> p.lg_tw) & ((1 << p.lg_th) - 1);
          const int n2 = (tnx >> (p.lg_tw + p.lg_th)) * p.BN + bn;
          const int i2 = th2 * p.BH + bh, j2 = tw2 * p.BW + bw;
          const bf16* z2 = nb.z + (((int64_t)n2 * p.Hb + 2 * i2) * p.Wb + 2 * j2) * p.Nch;
          nb_prefetch_l2(z2, 2 * p.Nch * 2, n2 < p.Nimg);                               // phases (0,0), (0,1): adjacent
          nb_prefetch_l2(z2 + (int64_t)p.Wb * p.Nch, 2 * p.Nch * 2, n2 < p.Nimg);      // phases (1,0), (1,1)
        }
      }
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      float s1 = 0.f, s2 = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * set_cols);
      if constexpr (NB) {
        for (int ph = 0; ph < 4; ++ph) {
          const int64_t poff = ((int64_t)(ph >> 1) * p.Wb + (ph & 1)) * p.Nch;
          bf16* orow = obase + poff;
          NormBwdZ zn;
          if (ph < 3) {
            const int pn = ph + 1;
            nb_load(zn, zbase + ((int64_t)(pn >> 1) * p.Wb + (pn & 1)) * p.Nch, p.NT >> 4, valid);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cb = 16 * c;
            if (cb < p.NT) {
              float v[16];
              tc::tmem_ld16(taddr + d4_pos(ph) * p.NT + cb, v);
              tc::add_bias16(v, sbias_u32, cb);
              uint32_t pk[8];
              nb_chunk(v, zc.v[2 * c], zc.v[2 * c + 1], coef, nb.alpha, s1, s2, pk);
              if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(orow + cb);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
          if (ph < 3) zc = zn;
        }
      } else {
      for (int ph = 0; ph < 4; ++ph) {
        bf16* orow = obase + ((int64_t)(ph >> 1) * p.Wb + (ph & 1)) * p.Nch;
        for (int cb = 0; cb < p.NT; cb += 16) {
          float v[16];
          tc::tmem_ld16(taddr + d4_pos(ph) * p.NT + cb, v);
          if (cb >= p.Nch) continue;
          tc::add_bias16(v, sbias_u32, cb);
          if (cb + 16 <= p.Nch && vec_ok) {
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
              uint4* dst = reinterpret_cast<uint4*>(orow + cb);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {                   // constant indices: no local-memory copy of v[]
              if (cb + e < p.Nch) {
                float a = v[e];
                s1 += a; s2 += a * a;
                if (p.act == LG_ACT_TANH) a = tanhf(a);
                if (valid) orow[cb + e] = __float2bfloat16_rn(a);
              }
            }
          }
        }
      }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (nacc == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      else acc_phase ^= 1;

      double* sums = NB ? nb.red : p.stats;
      if (sums != nullptr) {
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
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

bool plan_d4(int Nimg, int Hb, int Wb, int A, int B, D4Params* p) {
  const int Hs = Hb / 2, Ws = Wb / 2;
  if (!tc_host::is_pow2(Hs) || !tc_host::is_pow2(Ws) || Ws > 128 || Hs * Ws < 32) return false;
  if (B % 32 != 0) return false;
  const int Npad = (A + 15) / 16 * 16;
  // tiny A: tc_deconv_small.  A > 64 would need all 512 TMEM columns for one accumulator set (no double
  // buffering, 2 pipeline stages) and measured slower than the per-phase kernel (87 vs 66 us, dec2 shape).
  if (Npad > 64 || A < 16) return false;
  p->Nimg = Nimg; p->Hs = Hs; p->Ws = Ws; p->Hb = Hb; p->Wb = Wb;
  p->Kch = B; p->Nch = A; p->KC = (B % 64 == 0) ? 64 : 32; p->NT = Npad;
  p->BW = Ws < 128 ? Ws : 128;
  p->BH = (128 / p->BW) < Hs ? (128 / p->BW) : Hs;
  p->BN = 128 / (p->BW * p->BH);
  if (p->BW * p->BH * p->BN != 128 || p->BW * p->BH < 32) return false;
  const int tilesW = Ws / p->BW, tilesH = Hs / p->BH, tilesN = (Nimg + p->BN - 1) / p->BN;
  auto lg2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  p->lg_tw = lg2(tilesW); p->lg_th = lg2(tilesH);
  p->total_tiles = tilesW * tilesH * tilesN;
  p->a_bytes = 128 * p->KC * 2;
  p->b_bytes = p->NT * p->KC * 2;
  const int kc_per_tap = B / p->KC;
  p->cps = (kc_per_tap % 2 == 0) ? 2 : 1;
  p->stage_bytes = p->cps * (p->a_bytes + 4 * p->b_bytes);
  if (SMEM_BUDGET / p->stage_bytes < 2) { p->cps = 1; p->stage_bytes = p->a_bytes + 4 * p->b_bytes; }
  int st = SMEM_BUDGET / p->stage_bytes;
  p->stages = st > MAX_STAGES ? MAX_STAGES : st;
  if (p->stages < 2) return false;
  p->nacc = (8 * p->NT <= 512) ? 2 : 1;
  return true;
}

int encode_w_map3(CUtensorMap* m, const void* base, int rows, int cols, int boxCols, int boxRows, CUtensorMapSwizzle sw) {
  tc_host::EncodeTiledFn enc = tc_host::get_encode();
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

}  // namespace

int lg_tc_dgrad4_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  if (s != 2) return 0;
  D4Params p;
  return plan_d4(Nimg, Hb, Wb, A, B, &p) ? 1 : 0;
}

int lg_tc_dgrad4(const void* small, const void* wpack, const float* bias, void* out, double* stats, int Nimg, int Hb,
                 int Wb, int A, int B, int act, const lg_norm_bwd_t* nbh, cudaStream_t st) {
  D4Params p;
  if (!plan_d4(Nimg, Hb, Wb, A, B, &p)) { lg_set_error("tcgen05 dgrad4: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
  NormBwdDev nb = {};
  if (nbh != nullptr) {
    if (A % 16 != 0 || act != LG_ACT_NONE || stats != nullptr) {
      lg_set_error("tcgen05 dgrad4: fused norm-backward needs A %% 16 == 0, no activation, no forward statistics");
      return LG_ERR_UNSUPPORTED;
    }
    nb = lg_make_norm_bwd(nbh, (int64_t)Hb * Wb * A);
  }
  p.act = act; p.bias = bias; p.out = (bf16*)out; p.stats = stats;
  const int Ap = (A + 15) / 16 * 16, Bp = (B + 15) / 16 * 16;
  const CUtensorMapSwizzle sw = p.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmA, tmB;
  int e = tc_host::encode_act_map(&tmA, small, Nimg, p.Hs, p.Ws, B, p.KC, p.BW, p.BH, p.BN, 1, sw);
  if (e) return e;
  e = encode_w_map3(&tmB, wpack, Ap, Bp, p.KC, p.NT, sw);           // Wt = [25][Ap][Bp]
  if (e) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_dgrad4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_dgrad4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const size_t shm = (size_t)p.stages * p.stage_bytes + 1024 + 256 + 2048;
  const int grid = lg_even_grid(p.total_tiles, lg_num_sms());
  if (nbh != nullptr) tc_dgrad4_kernel<true><<<grid, NUM_THREADS, shm, st>>>(tmA, tmB, p, nb);
  else tc_dgrad4_kernel<false><<<grid, NUM_THREADS, shm, st>>>(tmA, tmB, p, nb);
  return LG_OK;
}
