// Convolutions whose "big" map is the 3-channel image (Cin = 3): the encoder's first Conv2D forward
// and weight gradient (model.py:15) and the input- / weight-gradient of the generator's final
// Conv2DTranspose(3) (model.py:86).  K per tap is only 3, so instead of 25 padded TMA boxes per tile the
// im2col row of a pixel - 75 bf16 values (ky, kx, ch), padded to 80 - is assembled ONCE in shared memory
// by 4 builder warps from a halo tile of the image, in the canonical no-swizzle core-matrix layout
// [k-chunk of 8][position][16 B].  That single tile feeds tcgen05 in both roles:
//
//   fprop : small[pos][b] = sum_k patch[pos][k] * W[k][b]            patch = K-major A operand (M = pos)
//   wgrad : dW[k][b]     += sum_pos patch[pos][k] * small[pos][b]    patch = MN-major A operand (K = pos)
//
// so the image is read ~once and the 64/32-channel map exactly once (HBM-bound layers, AI ~ 65 FLOP/B).
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "internal.h"
#include "norm_bwd.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int KPAD = 80;                    // 75 -> 80 (5 MMAs of K = 16)
constexpr int PLANE = 128 * 16;             // one k-chunk plane: 128 positions x 16 B
constexpr int MARGIN = 8;                   // zero pixels left/right of every staged image row

struct C3Params {
  int Nimg, Hb, Wb, Hs, Ws, s, pad, B;
  int BW, BH, rows_in;                      // tile = BH rows x BW cols of the small map; staged image rows
  int pitch;                                // bytes per staged image row: (Wb + 2*MARGIN) * 6
  int tiles_per_img, total_tiles;
  const bf16* img;                          // [N,Hb,Wb,3]
  const float* W;                           // [25][3][B] fp32 (fprop)
  const float* bias;
  bf16* out;                                // fprop output [N,Hs,Ws,B]
  double* stats;
  float* dW;                                // wgrad output [25][3][B]
  int b_blk, stage_bytes, stages;           // wgrad: TMA box channels / smem ring
};

// ---- im2col builder: 128 threads, thread m owns tile position m -------------------------------------
// Image halo rows s*i0 - pad ... (+rows_in), interior pixels only (margins stay zero), 16-byte vectors,
// at most 3 per thread.  Split into global->register and register->shared halves so that the loads of
// the NEXT tile are in flight while the patches of the current one are built.
constexpr int ROW_VECS = 3;
template <int S>
__device__ __forceinline__ void load_image_rows(const C3Params& p, int n, int i0, int bt, uint4 (&v)[ROW_VECS]) {
  const int vec_per_row = p.Wb * 6 / 16;
  const int y_first = S * i0 - p.pad, total = p.rows_in * vec_per_row;
#pragma unroll
  for (int q = 0; q < ROW_VECS; ++q) {
    const int e = bt + q * 128;
    v[q] = make_uint4(0, 0, 0, 0);
    if (e < total) {
      const int r = e / vec_per_row, c = e - r * vec_per_row;
      const int y = y_first + r;
      if (y >= 0 && y < p.Hb) v[q] = __ldg(reinterpret_cast<const uint4*>(p.img + ((int64_t)n * p.Hb + y) * p.Wb * 3) + c);
    }
  }
}
__device__ __forceinline__ void store_image_rows(const C3Params& p, uint8_t* simg, int bt, const uint4 (&v)[ROW_VECS]) {
  const int vec_per_row = p.Wb * 6 / 16, total = p.rows_in * vec_per_row;
#pragma unroll
  for (int q = 0; q < ROW_VECS; ++q) {
    const int e = bt + q * 128;
    if (e < total) {
      const int r = e / vec_per_row, c = e - r * vec_per_row;
      *reinterpret_cast<uint4*>(simg + r * p.pitch + MARGIN * 6 + c * 16) = v[q];
    }
  }
}

template <int S>
__device__ __forceinline__ void build_patches(const C3Params& p, const uint8_t* simg, uint8_t* sA, int m) {
  const int oi = m / p.BW, oj = m - oi * p.BW;
  const uint32_t simg_u32 = tc::smem_u32(simg);
  uint32_t pk[KPAD / 2];
#pragma unroll
  for (int i = 0; i < KPAD / 2; ++i) pk[i] = 0u;
#pragma unroll
  for (int ky = 0; ky < 5; ++ky) {
    // 15 bf16 of one kernel row, read through the shared window (behind the generic pointer these were 75 LD.E.U16
    // per position and tile)
    const uint32_t src = simg_u32 + (uint32_t)((S * oi + ky) * p.pitch + (S * oj - p.pad + MARGIN) * 6);
#pragma unroll
    for (int t = 0; t < 15; ++t) {
      const int k = ky * 15 + t;
      uint32_t v;
      asm volatile("{\n\t.reg .u16 h;\n\tld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(v) : "r"(src + 2u * (uint32_t)t));
      pk[k >> 1] |= (k & 1) ? (v << 16) : v;
    }
  }
#pragma unroll
  for (int c = 0; c < KPAD / 8; ++c)
    *reinterpret_cast<uint4*>(sA + c * PLANE + m * 16) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

__device__ __forceinline__ void bar_sync_builders() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// =====================================================================================================
// fprop: warps 0-3 builders, warp 4 MMA issuer + TMEM owner, warps 5-8 epilogue
// =====================================================================================================
constexpr int F_THREADS = 288;
constexpr int F_STAGES = 2;

template <int S, bool NB>
__global__ void __launch_bounds__(F_THREADS, 3)
cin3_fprop_kernel(const C3Params p, const NormBwdDev nb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                           // F_STAGES x 10 planes
  uint8_t* sW = sA + F_STAGES * (KPAD / 8) * PLANE;             // B rows x 80 k, no-swizzle K-major
  const int w_bytes = p.B * KPAD * 2;
  uint8_t* simg = sW + ((w_bytes + 1023) & ~1023);              // rows_in x pitch
  const int img_bytes = (p.rows_in * p.pitch + 15) & ~15;
  uint64_t* bars = reinterpret_cast<uint64_t*>(simg + ((img_bytes + 127) & ~127));
  uint64_t* full = bars;                 // [2] builders -> MMA        (128 arrivals)
  uint64_t* empty = bars + 2;            // [2] MMA -> builders
  uint64_t* tfull = bars + 4;            // [2] MMA -> epilogue
  uint64_t* tempty = bars + 6;           // [2] epilogue -> MMA        (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* sbias = reinterpret_cast<float*>(bars + 10);           // B floats (<= 128)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = 2 * p.B <= 32 ? 32 : 2 * p.B <= 64 ? 64 : 2 * p.B <= 128 ? 128 : 256;

  // weights: W[k][b] fp32 -> bf16 [b][k] no-swizzle K-major (8x16B cores, LBO = 128 B, SBO = 10*128 B)
  for (int e = threadIdx.x; e < p.B * KPAD; e += F_THREADS) {
    const int k = e / p.B, b = e - k * p.B;
    const float v = k < 75 ? p.W[(int64_t)k * p.B + b] : 0.f;
    const int off = (b >> 3) * ((KPAD / 8) * 128) + (k >> 3) * 128 + (b & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<bf16*>(sW + off) = __float2bfloat16_rn(v);
  }
  for (int e = threadIdx.x * 16; e < img_bytes; e += F_THREADS * 16)
    *reinterpret_cast<uint4*>(simg + e) = make_uint4(0, 0, 0, 0);     // margins (and everything else) start at zero
  for (int e = threadIdx.x; e < p.B; e += F_THREADS) sbias[e] = p.bias ? p.bias[e] : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&full[i], 128); tc::mbar_init(&empty[i], 1);
      tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 4) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------ builders
    const int bt = threadIdx.x;
    int stage = 0; uint32_t phase = 0;
    uint4 pre[ROW_VECS];
    if ((int)blockIdx.x < p.total_tiles)
      load_image_rows<S>(p, blockIdx.x / p.tiles_per_img, (blockIdx.x % p.tiles_per_img) * p.BH, bt, pre);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      bar_sync_builders();                                       // previous tile's patch reads of simg are done
      store_image_rows(p, simg, bt, pre);
      bar_sync_builders();
      const int tn = t + gridDim.x;                              // next tile's rows: in flight during the build
      if (tn < p.total_tiles) load_image_rows<S>(p, tn / p.tiles_per_img, (tn % p.tiles_per_img) * p.BH, bt, pre);
      tc::mbar_wait(&empty[stage], phase ^ 1);
      build_patches<S>(p, simg, sA + stage * (KPAD / 8) * PLANE, bt);
      tc::fence_proxy_async();                                   // generic writes -> async (tensor core) proxy
      tc::mbar_arrive(&full[stage]);
      if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 4) {
    // ------------------------------------------------ MMA issuer
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, p.B, 0, 0);
      const uint32_t sw_addr = tc::smem_u32(sW);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        tc::mbar_wait(&tempty[acc], aphase ^ 1);
        tc::mbar_wait(&full[stage], phase);
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(sA + stage * (KPAD / 8) * PLANE);
#pragma unroll
        for (int k = 0; k < KPAD / 16; ++k) {
          const uint64_t da = tc::make_sdesc(sa + k * 2 * PLANE, PLANE, 128, 0u);          // LBO = plane, SBO = 8 rows
          const uint64_t db = tc::make_sdesc(sw_addr + k * 256, 128, (KPAD / 8) * 128, 0u);
          tc::mma_bf16(tmem_base + acc * p.B, da, db, idesc, k != 0);
        }
        tc::mma_commit(&empty[stage]);
        tc::mma_commit(&tfull[acc]);
        if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int n = t / p.tiles_per_img, r = t - n * p.tiles_per_img;
      const int64_t pos = (int64_t)n * p.Hs * p.Ws + (int64_t)r * 128 + row;
      bf16* orow = p.out + pos * p.B;
      NormBwdCoef coef = {0.f, 0.f, 0.f, 0.f};
      NormBwdZn<2> zc;                                           // 32 channels at a time (3 CTAs/SM: few registers)
      const bf16* zrow = nullptr;
      if constexpr (NB) {                                        // fused InstanceNorm-backward reduction (norm_bwd.cuh)
        zrow = nb.z + pos * p.B;
        coef = nb_coef(nb, n);
        nb_load(zc, zrow, p.B >> 4, true);
        // the MMA of a tile is tiny here, so the z rows of this CTA's tile after next are sent towards L2
        // now (the first tile also sends the next one); without it every tile waits an HBM round trip
        for (int a = (t == (int)blockIdx.x ? 1 : 2); a <= 2; ++a) {
          const int tn = t + a * gridDim.x;
          if (tn < p.total_tiles) {
            const int n2 = tn / p.tiles_per_img, r2 = tn - n2 * p.tiles_per_img;
            nb_prefetch_l2(nb.z + ((int64_t)n2 * p.Hs * p.Ws + (int64_t)r2 * 128 + row) * p.B, p.B * 2, true);
          }
        }
      }
      tc::mbar_wait(&tfull[acc], aphase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.B);
      float s1 = 0.f, s2 = 0.f;
      if constexpr (NB) {
        for (int sc = 0; sc < p.B; sc += 32) {
          NormBwdZn<2> zn;
          if (sc + 32 < p.B) nb_load(zn, zrow + sc + 32, (p.B - sc - 32) >> 4, true);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int cb = sc + 16 * c;
            float v[16];
            tc::tmem_ld16(taddr + cb, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += sbias[cb + e];
            uint32_t pk[8];
            nb_chunk(v, zc.v[2 * c], zc.v[2 * c + 1], coef, nb.alpha, s1, s2, pk);
            uint4* dst = reinterpret_cast<uint4*>(orow + cb);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (sc + 32 < p.B) zc = zn;
        }
      } else {
      for (int cb = 0; cb < p.B; cb += 16) {
        float v[16];
        tc::tmem_ld16(taddr + cb, v);
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float a = v[e] + sbias[cb + e];
          float b = v[e + 1] + sbias[cb + e + 1];
          s1 += a + b; s2 += a * a + b * b;
          __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
          pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(orow + cb);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aphase ^= 1; }
      double* sums = NB ? nb.red : p.stats;
      if (sums != nullptr) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { atomicAdd(&sums[2 * n], (double)s1); atomicAdd(&sums[2 * n + 1], (double)s2); }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// =====================================================================================================
// wgrad: warps 0-3 builders, warp 4 MMA issuer + TMEM owner, warp 5 TMA producer (small map),
//        warps 6-9 epilogue (once, at the end: fp32 vector reductions into dW)
// =====================================================================================================
constexpr int W_THREADS = 320;
constexpr int W_APLANES = 16;               // M = 128 rows = 16 k-chunk planes (10 real + 6 zero)

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int S>
__global__ void __launch_bounds__(W_THREADS)
cin3_wgrad_kernel(const __grid_constant__ CUtensorMap tmSmall, const C3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = W_APLANES * PLANE;                         // 32 KB
  uint8_t* sStage = smem;                                        // stages x (patch planes | small tile)
  uint8_t* simg = sStage + (size_t)p.stages * p.stage_bytes;
  const int img_bytes = (p.rows_in * p.pitch + 15) & ~15;
  uint64_t* bars = reinterpret_cast<uint64_t*>(simg + ((img_bytes + 127) & ~127));
  uint64_t* full = bars;                 // [stages] 128 builders + 1 TMA expect_tx
  uint64_t* empty = bars + 4;            // [stages]
  uint64_t* tfull = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = p.B <= 32 ? 32 : p.B <= 64 ? 64 : 128;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) ++my_tiles;

  for (int e = threadIdx.x * 16; e < p.stages * p.stage_bytes + img_bytes + 128; e += W_THREADS * 16)
    if (e < p.stages * p.stage_bytes + ((img_bytes + 127) & ~127))
      *reinterpret_cast<uint4*>(smem + e) = make_uint4(0, 0, 0, 0);   // zero planes 10..15 and the image margins
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full[i], 129); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(tfull, 1);
    tc::fence_barrier_init();
    tc::tma_prefetch_desc(&tmSmall);
  }
  if (warp == 4) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    const int bt = threadIdx.x;
    int stage = 0; uint32_t phase = 0;
    uint4 pre[ROW_VECS];
    if ((int)blockIdx.x < p.total_tiles)
      load_image_rows<S>(p, blockIdx.x / p.tiles_per_img, (blockIdx.x % p.tiles_per_img) * p.BH, bt, pre);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      bar_sync_builders();
      store_image_rows(p, simg, bt, pre);
      bar_sync_builders();
      const int tn = t + gridDim.x;
      if (tn < p.total_tiles) load_image_rows<S>(p, tn / p.tiles_per_img, (tn % p.tiles_per_img) * p.BH, bt, pre);
      tc::mbar_wait(&empty[stage], phase ^ 1);
      build_patches<S>(p, simg, sStage + (size_t)stage * p.stage_bytes, bt);
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 5) {
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      const int nbox = p.B / p.b_blk;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n = t / p.tiles_per_img, r = t - n * p.tiles_per_img;
        tc::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sb = sStage + (size_t)stage * p.stage_bytes + a_bytes;
        tc::mbar_expect_tx(&full[stage], (uint32_t)(128 * p.B * 2));
        for (int bi = 0; bi < nbox; ++bi)
          tc::tma_load_4d(sb + bi * 128 * p.b_blk * 2, &tmSmall, &full[stage], bi * p.b_blk, 0, r * p.BH, n);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (my_tiles > 0 && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, p.B, 1, 1);     // both operands MN-major
      const uint32_t layout_b = p.b_blk == 64 ? 2u : 4u;
      const uint32_t sbo_b = 8u * (uint32_t)p.b_blk * 2u, lbo_b = 128u * (uint32_t)p.b_blk * 2u;
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        tc::mbar_wait(&full[stage], phase);
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(sStage + (size_t)stage * p.stage_bytes);
        const uint32_t sb = sa + (uint32_t)a_bytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {                            // 128 positions = 8 x K16
          // patches, MN-major no-swizzle: LBO = next 8 positions (128 B), SBO = next 8 k (one plane)
          const uint64_t da = tc::make_sdesc(sa + k * 256, 128, PLANE, 0u);
          const uint64_t db = tc::make_sdesc(sb + k * 2 * sbo_b, lbo_b, sbo_b, layout_b);
          tc::mma_bf16(tmem_base, da, db, idesc, !(first && k == 0));
        }
        first = false;
        tc::mma_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      tc::mma_commit(tfull);
    }
  } else if (my_tiles > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;                               // k = (tap*3 + ch)
    tc::mbar_wait(tfull, 0);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* orow = p.dW + (int64_t)row * p.B;
    for (int cb = 0; cb < p.B; cb += 16) {
      float v[16];
      tc::tmem_ld16(taddr + cb, v);
      if (row < 75) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) red_add_v4(orow + cb + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

bool plan_c3(int Nimg, int Hb, int Wb, int B, int s, C3Params* p) {
  if (s != 1 && s != 2) return false;
  const int Hs = Hb / s, Ws = Wb / s;
  if ((Wb * 6) % 16 != 0 || (B != 32 && B != 64 && B != 128)) return false;
  if (!(Ws == 128 || Ws == 64 || Ws == 32) || (Hs * Ws) % 128 != 0) return false;
  p->Nimg = Nimg; p->Hb = Hb; p->Wb = Wb; p->Hs = Hs; p->Ws = Ws; p->s = s; p->pad = (s == 2) ? 1 : 2; p->B = B;
  p->BW = Ws; p->BH = 128 / Ws;
  p->rows_in = s * (p->BH - 1) + 5;
  p->pitch = (Wb + 2 * MARGIN) * 6;
  if (p->pitch % 16 != 0 || p->rows_in * (Wb * 6 / 16) > ROW_VECS * 128) return false;
  p->tiles_per_img = Hs * Ws / 128;
  p->total_tiles = Nimg * p->tiles_per_img;
  p->b_blk = (B % 64 == 0) ? 64 : 32;
  p->stage_bytes = W_APLANES * PLANE + 128 * B * 2;
  p->stages = 3;
  return true;
}

}  // namespace

int lg_tc_cin3_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  C3Params p;
  return (A == 3 && plan_c3(Nimg, Hb, Wb, B, s, &p)) ? 1 : 0;
}

int lg_tc_cin3_fprop(const void* img, const float* W, const float* bias, void* out, double* stats, int Nimg, int Hb,
                     int Wb, int B, int s, const lg_norm_bwd_t* nbh, cudaStream_t st) {
  C3Params p;
  if (!plan_c3(Nimg, Hb, Wb, B, s, &p) || !W) { lg_set_error("cin3 fprop: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
  NormBwdDev nb = {};
  if (nbh != nullptr) {
    if (stats != nullptr) { lg_set_error("cin3 fprop: fused norm-backward excludes forward statistics"); return LG_ERR_UNSUPPORTED; }
    nb = lg_make_norm_bwd(nbh, (int64_t)p.Hs * p.Ws * B);
  }
  p.img = (const bf16*)img; p.W = W; p.bias = bias; p.out = (bf16*)out; p.stats = stats; p.dW = nullptr;
  const size_t shm = (size_t)F_STAGES * (KPAD / 8) * PLANE + ((B * KPAD * 2 + 1023) & ~1023) +
                     (size_t)p.rows_in * p.pitch + 2048 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cin3_fprop_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(cin3_fprop_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(cin3_fprop_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(cin3_fprop_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  int grid = 3 * lg_num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (nbh != nullptr) {
    if (s == 1) cin3_fprop_kernel<1, true><<<grid, F_THREADS, shm, st>>>(p, nb);
    else cin3_fprop_kernel<2, true><<<grid, F_THREADS, shm, st>>>(p, nb);
  } else {
    if (s == 1) cin3_fprop_kernel<1, false><<<grid, F_THREADS, shm, st>>>(p, nb);
    else cin3_fprop_kernel<2, false><<<grid, F_THREADS, shm, st>>>(p, nb);
  }
  return LG_OK;
}

int lg_tc_cin3_wgrad(const void* img, const void* small, float* dW, int Nimg, int Hb, int Wb, int B, int s,
                     cudaStream_t st) {
  C3Params p;
  if (!plan_c3(Nimg, Hb, Wb, B, s, &p)) { lg_set_error("cin3 wgrad: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
  p.img = (const bf16*)img; p.W = nullptr; p.bias = nullptr; p.out = nullptr; p.stats = nullptr; p.dW = dW;
  CUtensorMap tmSmall;
  int e = tc_host::encode_act_map(&tmSmall, small, Nimg, p.Hs, p.Ws, B, p.b_blk, p.BW, p.BH, 1, 1,
                                  p.b_blk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (e) return e;
  const size_t shm = (size_t)p.stages * p.stage_bytes + (size_t)p.rows_in * p.pitch + 2048;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cin3_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(cin3_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = lg_num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (s == 1) cin3_wgrad_kernel<1><<<grid, W_THREADS, shm, st>>>(tmSmall, p);
  else cin3_wgrad_kernel<2><<<grid, W_THREADS, shm, st>>>(tmSmall, p);
  return LG_OK;
}
