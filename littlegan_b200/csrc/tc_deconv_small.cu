// Transposed convolution (conv dgrad) with a TINY number of output channels (A <= 5: the RGB image
// layers - the generator's final Conv2DTranspose(3, k5, s1, tanh), model.py:86, and the input-gradient
// of the encoder's first Conv2D, model.py:15) as GEMM + col2im on tcgen05:
//
//   Z[pixel][(tap,a)] = sum_b small[pixel][b] * W[tap][a][b]          (one 256 x NZ x B GEMM per tile)
//   big[n, Y, X, a]   = bias[a] + sum_{taps, s*i+ky-pad == Y, s*j+kx-pad == X} Z[(i,j)][(tap,a)]
//
// The generic dgrad kernel must re-fetch its input once per tap (25x) because every tap needs a shifted
// view of it; with 3 output channels the whole 25-tap response of a pixel is only NZ = 75 numbers, so the
// input tile (16x16 pixels incl. halo, one TMA box, zero-filled out of bounds) is read ONCE, multiplied
// by the [NZ x B] weight matrix resident in shared memory, and the overlap-add over taps happens in
// shared memory.  Epilogue: + bias, per-sample sum / sum-of-squares (InstanceNorm statistics / none for
// the image), optional tanh, bf16 store.
//
// Persistent; warp 0 = TMA producer (2-stage ring), warp 1 = MMA issuer + TMEM owner, warps 2..5 =
// epilogue (TMEM -> fp32 smem -> gather -> global).
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "internal.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int NUM_THREADS = 192;
constexpr int PRODUCER_WARP = 4, MMA_WARP = 5;   // epilogue = warps 0-3; single-thread roles get the high warp ids
constexpr int TILE = 16;                // input tile is TILE x TILE pixels = 256 GEMM rows (2 MMA blocks)
constexpr int STAGES = 2;
constexpr int ZP = 81;                   // fp32 pitch of the Z staging rows (odd -> conflict-free)

struct DsParams {
  int Nimg, Hs, Ws, Hb, Wb, s, pad, halo;
  int A, B, NZ;                          // NZ = round_up(25*A, 16) <= 128
  int out_t;                             // output pixels per tile edge: s*(TILE - 2*halo)
  int tilesH, tilesW, total_tiles;
  int a_bytes;                           // one input tile: 256 rows x B bf16
  int act;
  const float* W;                        // [25][A][B] fp32
  const float* bias;
  bf16* out;
  double* stats;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int S, int A>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_deconv_small_kernel(const __grid_constant__ CUtensorMap tmA, const DsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                               // STAGES x a_bytes
  uint8_t* sW = sA + STAGES * p.a_bytes;                            // NZ rows x B bf16, no-swizzle K-major
  const int w_bytes = p.NZ * p.B * 2;
  float* Zs = reinterpret_cast<float*>(sW + ((w_bytes + 1023) & ~1023));   // 256 x ZP fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(Zs) + 256 * ZP * 4);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [1]
  uint64_t* tempty = tfull + 1;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights -> bf16, canonical no-swizzle K-major: 8x16B core matrices, LBO (next 8 k) = 128 B,
  // SBO (next 8 rows) = (B/8)*128 B
  {
    const int kcores = p.B / 8;
    for (int e = threadIdx.x; e < p.NZ * p.B; e += NUM_THREADS) {
      const int r = e / p.B, k = e - r * p.B;
      const float v = r < 25 * p.A ? p.W[(int64_t)r * p.B + k] : 0.f;
      const int off = (r >> 3) * (kcores * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
      *reinterpret_cast<bf16*>(sW + off) = __float2bfloat16_rn(v);
    }
  }
  if (warp == PRODUCER_WARP && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(tfull, 1);
    tc::mbar_init(tempty, 4);
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, 256);
  tc::fence_proxy_async();               // generic-proxy smem writes (weights) -> visible to the tensor core
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int in_t = TILE - 2 * p.halo;    // interior input pixels per tile edge
  const int per_img = p.tilesH * p.tilesW;

  if (warp == PRODUCER_WARP) {
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n = t / per_img, r = t - n * per_img;
        const int ti = r / p.tilesW, tj = r - ti * p.tilesW;
        tc::mbar_wait(&empty[stage], phase ^ 1);
        tc::mbar_expect_tx(&full[stage], (uint32_t)p.a_bytes);
        tc::tma_load_4d(sA + stage * p.a_bytes, &tmA, &full[stage], 0, tj * in_t - p.halo, ti * in_t - p.halo, n);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, p.NZ, 0, 0);
      const uint32_t layout_a = (p.B == 64) ? 2u : 4u;              // SWIZZLE_128B / SWIZZLE_64B rows
      const uint32_t row_bytes = (uint32_t)p.B * 2u;
      const uint32_t sbo_a = 8u * row_bytes;
      const uint32_t sbo_w = (uint32_t)(p.B / 8) * 128u;
      const uint32_t sw_addr = tc::smem_u32(sW);
      int stage = 0; uint32_t phase = 0, tphase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        tc::mbar_wait(tempty, tphase ^ 1);
        tc::mbar_wait(&full[stage], phase);
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(sA + stage * p.a_bytes);
        for (int mb = 0; mb < 2; ++mb) {
          for (int k = 0; k < p.B / 16; ++k) {
            const uint64_t da = tc::make_sdesc(sa + mb * 128 * row_bytes + k * 32, 16, sbo_a, layout_a);
            const uint64_t db = tc::make_sdesc(sw_addr + k * 256, 128, sbo_w, 0u);
            tc::mma_bf16(tmem_base + mb * 128, da, db, idesc, k != 0);
          }
        }
        tc::mma_commit(&empty[stage]);
        tc::mma_commit(tfull);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        tphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int et = q * 32 + lane;        // 0..127: TMEM lane of this thread / epilogue thread id
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t tphase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int n = t / per_img, r = t - n * per_img;
      const int ti = r / p.tilesW, tj = r - ti * p.tilesW;
      tc::mbar_wait(tfull, tphase);
      tc::fence_after_sync();
      tphase ^= 1;
      // TMEM -> fp32 staging (row = input pixel of the tile, col = (tap, a))
      for (int mb = 0; mb < 2; ++mb) {
        float* zrow = Zs + (mb * 128 + et) * ZP;
        for (int cb = 0; cb < p.NZ; cb += 16) {
          float v[16];
          tc::tmem_ld16(taddr + mb * 128 + cb, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) zrow[cb + e] = v[e];
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty);                       // accumulators may be overwritten
      epi_bar_sync();                                               // staging complete

      // col2im gather: output pixel (Y0+oy, X0+ox); input row/col of tap k: (Y + pad - k)/s, local = - origin
      const int Y0 = S * ti * in_t, X0 = S * tj * in_t;
      float s1 = 0.f, s2 = 0.f;
      float bia[A];
#pragma unroll
      for (int a = 0; a < A; ++a) bia[a] = p.bias ? p.bias[a] : 0.f;
      // Outputs are walked per phase (py,px) = (Y,X) mod S so that the tap set is a compile-time
      // list: tap ky_m = pad_ - py + S*m reads tile row oy2 + HALO + (py >= pad_ ...) - m (see below).
      constexpr int IN_T = TILE - 2 * (S == 2 ? 1 : 2);            // 14 (s=2) / 12 (s=1) interior pixels
#pragma unroll
      for (int py = 0; py < S; ++py) {
#pragma unroll
        for (int px = 0; px < S; ++px) {
          // S == 1: ky = 0..4, tile row = oy + 4 - ky.   S == 2: ky = 1 - py + 2m, tile row = oy2 + 1 + py - m
          constexpr int NTY = 5, NTX = 5;
          const int nty = (S == 1) ? NTY : 2 + py, ntx = (S == 1) ? NTX : 2 + px;
          for (int o = et; o < IN_T * IN_T; o += 128) {
            const int oy2 = o / IN_T, ox2 = o - oy2 * IN_T;
            const int Y = Y0 + S * oy2 + py, X = X0 + S * ox2 + px;
            if (Y >= p.Hb || X >= p.Wb) continue;
            float acc[A];
#pragma unroll
            for (int a = 0; a < A; ++a) acc[a] = bia[a];
#pragma unroll
            for (int my = 0; my < NTY; ++my) {
              if (my < nty) {
                const int ky = (S == 1) ? my : 1 - py + 2 * my;
                const int li = (S == 1) ? oy2 + 4 - my : oy2 + 1 + py - my;
#pragma unroll
                for (int mx = 0; mx < NTX; ++mx) {
                  if (mx < ntx) {
                    const int kx = (S == 1) ? mx : 1 - px + 2 * mx;
                    const int lj = (S == 1) ? ox2 + 4 - mx : ox2 + 1 + px - mx;
                    const float* z = Zs + (li * TILE + lj) * ZP + (ky * 5 + kx) * A;
#pragma unroll
                    for (int a = 0; a < A; ++a) acc[a] += z[a];
                  }
                }
              }
            }
            bf16* dst = p.out + (((int64_t)n * p.Hb + Y) * p.Wb + X) * A;
#pragma unroll
            for (int a = 0; a < A; ++a) {
              float v = acc[a];
              s1 += v; s2 += v * v;
              if (p.act == LG_ACT_TANH) v = tanhf(v);
              dst[a] = __float2bfloat16_rn(v);
            }
          }
        }
      }
      if (p.stats != nullptr) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (tc::elect_one()) { atomicAdd(&p.stats[2 * n], (double)s1); atomicAdd(&p.stats[2 * n + 1], (double)s2); }
      }
      epi_bar_sync();                                               // Zs free for the next tile
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

bool plan_ds(int Nimg, int Hb, int Wb, int A, int B, int s, DsParams* p) {
  if (s != 1 && s != 2) return false;
  if (A != 3 || (B != 32 && B != 64)) return false;      // instantiated for the RGB layers
  const int Hs = Hb / s, Ws = Wb / s;
  if (Hs < TILE || Ws < TILE) return false;
  p->Nimg = Nimg; p->Hs = Hs; p->Ws = Ws; p->Hb = Hb; p->Wb = Wb; p->s = s;
  p->pad = (s == 2) ? 1 : 2;
  p->halo = (s == 2) ? 1 : 2;            // input pixels a tile's outputs reach beyond its interior
  p->A = A; p->B = B; p->NZ = (25 * A + 15) / 16 * 16;
  if (p->NZ > 128) return false;
  const int in_t = TILE - 2 * p->halo;
  p->out_t = s * in_t;
  p->tilesH = (Hs + in_t - 1) / in_t; p->tilesW = (Ws + in_t - 1) / in_t;
  p->total_tiles = Nimg * p->tilesH * p->tilesW;
  p->a_bytes = 256 * B * 2;
  return true;
}

}  // namespace

int lg_tc_deconv_small_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  DsParams p;
  return plan_ds(Nimg, Hb, Wb, A, B, s, &p) ? 1 : 0;
}

int lg_tc_deconv_small(const void* small, const float* W, const float* bias, void* out, double* stats, int Nimg,
                       int Hb, int Wb, int A, int B, int s, int act, cudaStream_t st) {
  DsParams p;
  if (!plan_ds(Nimg, Hb, Wb, A, B, s, &p) || W == nullptr) {
    lg_set_error("tcgen05 small-Cout dgrad: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  p.act = act; p.W = W; p.bias = bias; p.out = (bf16*)out; p.stats = stats;
  CUtensorMap tmA;
  int e = tc_host::encode_act_map(&tmA, small, Nimg, p.Hs, p.Ws, B, B, TILE, TILE, 1, 1,
                                  B == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (e) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_deconv_small_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tc_deconv_small_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const size_t shm = (size_t)STAGES * p.a_bytes + ((p.NZ * B * 2 + 1023) & ~1023) + 256 * ZP * 4 + 1024 + 256;
  const int grid = p.total_tiles < lg_num_sms() ? p.total_tiles : lg_num_sms();
  if (s == 1) tc_deconv_small_kernel<1, 3><<<grid, NUM_THREADS, shm, st>>>(tmA, p);
  else tc_deconv_small_kernel<2, 3><<<grid, NUM_THREADS, shm, st>>>(tmA, p);
  return LG_OK;
}
