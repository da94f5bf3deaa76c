// tcgen05 weight-gradient kernel (bf16 operands, fp32 accumulate, fp32 atomics into dW):
//
//   dW[tap][a][b] += sum_{pos=(n,i,j)} big[n, s*i+ky-pad, s*j+kx-pad, a] * small[n,i,j,b]
//
// GEMM view: D[(tap,a)][b], M = 128 rows = (128/min(A,128)) taps x min(A,128) channels of the big
// map, N = b-tile (<= 256), K = positions.  Both operands are MN-major (channels contiguous, the
// contraction index = position has a stride), which tcgen05 consumes directly from the NHWC
// tensors: every k-block is a set of 4-D TMA boxes {64|32 channels x 64 positions}; the big-map
// boxes use element strides (s,s) and a per-tap start offset (zero-filled out of bounds), so the
// 25 shifted views never exist in memory.  Split-K over position slices fills the GPU; each CTA
// reduces its 128 x NT fp32 tile into dW with 16-byte vector reductions (red.global.add.v4.f32),
// where a thread owns a (tap,a) row and therefore contiguous b.
//
// One tile x one k-slice per CTA: warps 0-3 = epilogue, warp 5 = MMA issuer + TMEM owner, and THREE TMA producer
// warps (4: the big-map boxes, 6 / 7: the even / odd small-map boxes; a fourth producer - big-map boxes split too,
// MMA warp below the producers - was slower: 689 vs 950 TFLOP/s on enc2).  A stage is 4-6 boxes of 8 KB against four MMAs, and one
// thread gets a wait / expect / TMA instruction out only every ~60-100 ns (scripts/ubench/tma_rows.cu: the per-SM
// TMA rate doubles with a second issuing thread): with a single producer the N = 128 layers ran at the producer's
// issue rate (6 instructions = ~350 ns per stage against ~190 ns of MMAs).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int NUM_THREADS = 256;
constexpr int PRODUCER_WARP = 4, MMA_WARP = 5, PRODUCER_B_WARP = 6;   // epilogue = warps 0-3; warps 6, 7: small-map boxes
constexpr int NUM_PRODUCERS = 3;
constexpr int MAX_STAGES = 8;
constexpr int KP = 64;                   // positions per k-block
constexpr int SMEM_BUDGET = 200 * 1024;

struct WgParams {
  int Nimg, Hs, Ws, Hb, Wb, s, pad, A, B;
  int A_real;           // rows of dW per tap (A may be zero-padded storage, e.g. 3 -> 16)
  int b_blk;            // channels per small-map TMA box: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int a_blk;            // channels per big-map TMA box: 64 (SWIZZLE_128B), 32 (SWIZZLE_64B), 16 (SWIZZLE_32B)
  int rows_per_tap;     // min(A, 128)
  int taps_per_tile;    // 128 / rows_per_tap
  int a_tiles;          // ceil(A / 128)
  int groups;           // ceil(25 / taps_per_tile)
  int NT, n_tiles;      // b tile (multiple of 64)
  int BW, BH, BN;       // position box in small-map coordinates, BW*BH*BN == 64
  int pbW, pbH, pbN;    // position boxes along W, H, N
  int total_kb, kb_per_slice;
  int stages, stage_bytes, a_bytes;
  float* dW;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmBig, const __grid_constant__ CUtensorMap tmSmall,
                const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tfull = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = p.NT <= 32 ? 32 : p.NT <= 64 ? 64 : p.NT <= 128 ? 128 : 256;

  // tile decode: blockIdx.x -> (tap group, a tile, b tile); blockIdx.y -> k slice
  int t = blockIdx.x;
  const int nt = t % p.n_tiles; t /= p.n_tiles;
  const int at = t % p.a_tiles;
  const int grp = t / p.a_tiles;
  const int kb0 = blockIdx.y * p.kb_per_slice;
  const int kb1 = min(p.total_kb, kb0 + p.kb_per_slice);
  const int num_kb = kb1 - kb0;

  if (warp == PRODUCER_WARP && lane == 0) {
    tc::tma_prefetch_desc(&tmBig);
    tc::tma_prefetch_desc(&tmSmall);
    for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full[i], NUM_PRODUCERS); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(tfull, 1);
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int nbox_a = 128 / p.a_blk;                 // big-map boxes per stage
  const int nbox_b = p.NT / p.b_blk;                // small-map boxes per stage
  const int a_box_bytes = KP * p.a_blk * 2;
  const int b_box_bytes = KP * p.b_blk * 2;

  if (warp == PRODUCER_WARP) {
    if (num_kb > 0 && tc::elect_one()) {
      // per-box constants (tap offset, channel) do not depend on the k-block: hoist them
      int box_ch[8], box_dx[8], box_dy[8];
#pragma unroll
      for (int bi = 0; bi < 8; ++bi) {
        const int row0 = bi * p.a_blk;
        int tap = grp * p.taps_per_tile + row0 / p.rows_per_tap;
        if (tap > 24) tap = 24;                                     // padding rows: computed, never stored
        box_ch[bi] = at * 128 + row0 % p.rows_per_tap;
        const int ky = tap / 5;
        box_dy[bi] = ky - p.pad; box_dx[bi] = (tap - 5 * ky) - p.pad;
      }
      const uint32_t smem_u = tc::smem_u32(smem), full_u = tc::smem_u32(full), empty_u = tc::smem_u32(empty);
      const uint32_t stage_bytes_u = (uint32_t)p.stage_bytes, a_bytes_u = (uint32_t)p.a_bytes;
      const int nstages = p.stages;
      int pw = kb0 % p.pbW, ph = (kb0 / p.pbW) % p.pbH, pn = kb0 / (p.pbW * p.pbH);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int j0 = pw * p.BW, i0 = ph * p.BH, n0 = pn * p.BN;
        const uint32_t fb = full_u + (uint32_t)stage * 8u;
        const uint32_t sa = smem_u + (uint32_t)stage * stage_bytes_u;
        tc::mbar_wait_addr(empty_u + (uint32_t)stage * 8u, phase ^ 1);
        tc::mbar_expect_tx_addr(fb, a_bytes_u);
        const int bx = p.s * j0, by = p.s * i0;
#pragma unroll
        for (int bi = 0; bi < 8; ++bi)
          if (bi < nbox_a)
            tc::tma_load_4d_addr(sa + bi * a_box_bytes, &tmBig, fb, box_ch[bi], bx + box_dx[bi], by + box_dy[bi], n0);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
        if (++pw == p.pbW) { pw = 0; if (++ph == p.pbH) { ph = 0; ++pn; } }
      }
    }
  } else if (warp >= PRODUCER_B_WARP) {
    // two threads share the small-map boxes of a stage (even / odd box index)
    const int half = warp - PRODUCER_B_WARP;
    const int mine = (nbox_b - half + 1) >> 1;
    if (num_kb > 0 && tc::elect_one()) {
      const uint32_t smem_u = tc::smem_u32(smem), full_u = tc::smem_u32(full), empty_u = tc::smem_u32(empty);
      const uint32_t stage_bytes_u = (uint32_t)p.stage_bytes, a_bytes_u = (uint32_t)p.a_bytes;
      const int nstages = p.stages, b_ch0 = nt * p.NT;
      int pw = kb0 % p.pbW, ph = (kb0 / p.pbW) % p.pbH, pn = kb0 / (p.pbW * p.pbH);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int j0 = pw * p.BW, i0 = ph * p.BH, n0 = pn * p.BN;
        const uint32_t fb = full_u + (uint32_t)stage * 8u;
        const uint32_t sa = smem_u + (uint32_t)stage * stage_bytes_u;
        tc::mbar_wait_addr(empty_u + (uint32_t)stage * 8u, phase ^ 1);
        tc::mbar_expect_tx_addr(fb, (uint32_t)(mine * b_box_bytes));
#pragma unroll
        for (int bi = 0; bi < 4; ++bi)
          if (bi < nbox_b && (bi & 1) == half)
            tc::tma_load_4d_addr(sa + a_bytes_u + bi * b_box_bytes, &tmSmall, fb, b_ch0 + bi * p.b_blk, j0, i0, n0);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
        if (++pw == p.pbW) { pw = 0; if (++ph == p.pbH) { ph = 0; ++pn; } }
      }
    }
  } else if (warp == MMA_WARP) {
    if (num_kb > 0 && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, p.NT, 1, 1);        // both operands MN-major
      const uint32_t layout_a = (p.a_blk == 64) ? 2u : (p.a_blk == 32) ? 4u : 6u;
      const uint32_t sbo_a = 8u * (uint32_t)p.a_blk * 2u;            // 8 positions x a_blk channels
      const uint32_t lbo_a = (uint32_t)a_box_bytes;                  // next block of a_blk channels
      const uint32_t layout_b = (p.b_blk == 64) ? 2u : 4u;
      const uint32_t sbo_b = 8u * (uint32_t)p.b_blk * 2u, lbo_b = (uint32_t)b_box_bytes;
      // descriptors as (lo, hi): hi = SBO>>4 | version<<14 | layout<<29 (invariant); lo = start>>4 | LBO>>4<<16
      const uint32_t a_hi = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (layout_a << 29);
      const uint32_t b_hi = ((sbo_b >> 4) & 0x3FFFu) | (1u << 14) | (layout_b << 29);
      const uint32_t smem_u = tc::smem_u32(smem);
      const uint32_t a_lo0 = ((smem_u & 0x3FFFFu) >> 4) | (((lbo_a >> 4) & 0x3FFFu) << 16);
      const uint32_t b_lo0 = (((smem_u + (uint32_t)p.a_bytes) & 0x3FFFFu) >> 4) | (((lbo_b >> 4) & 0x3FFFu) << 16);
      const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
      const uint32_t a_step = (2 * sbo_a) >> 4, b_step = (2 * sbo_b) >> 4;   // 16 positions further along K
      const uint32_t full_u = tc::smem_u32(full), empty_u = tc::smem_u32(empty);
      const int nstages = p.stages;
      int stage = 0; uint32_t phase = 0, accum = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait_addr(full_u + (uint32_t)stage * 8u, phase);
        tc::fence_after_sync();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_units, b_lo = b_lo0 + (uint32_t)stage * stage_units;
#pragma unroll
        for (int k = 0; k < KP / 16; ++k) {
          tc::mma_bf16_lohi(tmem_base, a_lo + k * a_step, a_hi, b_lo + k * b_step, b_hi, idesc, accum);
          accum = 1;
        }
        tc::mma_commit_addr(empty_u + (uint32_t)stage * 8u);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      tc::mma_commit(tfull);
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tap = grp * p.taps_per_tile + row / p.rows_per_tap;
    const int a = at * 128 + row % p.rows_per_tap;
    const bool valid = tap < 25 && a < p.A_real;
    float* orow = p.dW + ((int64_t)tap * p.A_real + a) * p.B + nt * p.NT;
    tc::mbar_wait(tfull, 0);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int cb = 0; cb < p.NT; cb += 16) {
      float v[16];
      tc::tmem_ld16(taddr + cb, v);
      if (valid) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) red_add_v4(orow + cb + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
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

// Fixed cost of a CTA lifetime (prologue, pipeline fill, reduction epilogue) in k-block times, for the split-K choice.
// 12 was measured before the producer warps tripled the main loop's rate; a k-block is now so much shorter that
// the same fixed cost is worth 20-30 of them (enc4 at batch 128: 3 waves of 32 k-blocks 51.6 us, 2 waves of 64
// k-blocks 45.9 us; 20 .. 45 choose the same splits on every layer of the model).
inline int wg_overhead() {
  static const int v = getenv("LG_WG_OVERHEAD") ? atoi(getenv("LG_WG_OVERHEAD")) : 24;
  return v;
}

// A = channels of `big` as stored (16, 32, 64 or a multiple of 128); A_real <= A rows of dW per tap.
bool plan_wgrad(int Nimg, int Hb, int Wb, int A, int A_real, int B, int s, WgParams* p) {
  if (s != 1 && s != 2) return false;
  const int Hs = Hb / s, Ws = Wb / s;
  if (!tc_host::is_pow2(Hs) || !tc_host::is_pow2(Ws) || Ws > 128 || Hs * Ws < KP) return false;
  if ((B % 64 != 0 && B != 32) || A_real > A || A_real < 1) return false;
  if (A != 16 && A != 32 && A != 64 && A % 128 != 0) return false;
  p->Nimg = Nimg; p->Hs = Hs; p->Ws = Ws; p->Hb = Hb; p->Wb = Wb; p->s = s; p->pad = (s == 2) ? 1 : 2;
  p->A = A; p->A_real = A_real; p->B = B;
  p->a_blk = (A % 64 == 0) ? 64 : A;
  p->rows_per_tap = A < 128 ? A : 128;
  p->taps_per_tile = 128 / p->rows_per_tap;
  p->a_tiles = (A + 127) / 128;
  p->groups = (25 + p->taps_per_tile - 1) / p->taps_per_tile;
  p->b_blk = (B % 64 == 0) ? 64 : 32;
  int n_tiles = (B + 255) / 256;
  while (B % n_tiles != 0 || (B / n_tiles) % p->b_blk != 0) { if (++n_tiles > B / p->b_blk) return false; }
  p->n_tiles = n_tiles; p->NT = B / n_tiles;
  p->BW = Ws < KP ? Ws : KP;
  p->BH = (KP / p->BW) < Hs ? (KP / p->BW) : Hs;
  p->BN = KP / (p->BW * p->BH);
  if (p->BW * p->BH * p->BN != KP) return false;
  p->pbW = Ws / p->BW; p->pbH = Hs / p->BH; p->pbN = (Nimg + p->BN - 1) / p->BN;
  p->total_kb = p->pbW * p->pbH * p->pbN;
  p->a_bytes = KP * 128 * 2;
  p->stage_bytes = p->a_bytes + KP * p->NT * 2;
  int st = SMEM_BUDGET / p->stage_bytes;
  p->stages = st > MAX_STAGES ? MAX_STAGES : st;
  if (p->stages < 2) return false;
  // Split K so that the CTAs (one per SM: a CTA holds ~200 KB of shared memory) come in WHOLE waves: with
  // tiles x slices just above a multiple of the SM count the last wave runs a handful of CTAs on an otherwise
  // idle GPU (300 CTAs on 148 SMs took three CTA lifetimes instead of two).  Cost of a candidate =
  // waves x (k-blocks per CTA + a fixed prologue / pipeline fill / reduction epilogue of wg_overhead() k-block times).
  const int tiles = p->groups * p->a_tiles * p->n_tiles;
  const int sms = lg_num_sms();
  const int max_slices = (p->total_kb + 3) / 4;           // at least ~4 k-blocks per CTA
  int best_kbs = p->total_kb;
  long best_cost = -1;
  for (int w = 1; w <= 4; ++w) {
    int sl = (w * sms) / tiles;
    if (sl < 1) continue;
    if (sl > max_slices) sl = max_slices;
    const int kbs = (p->total_kb + sl - 1) / sl;
    const int ctas = tiles * ((p->total_kb + kbs - 1) / kbs);
    const long cost = (long)((ctas + sms - 1) / sms) * (kbs + wg_overhead());
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_kbs = kbs; }
  }
  p->kb_per_slice = best_kbs;
  return true;
}

}  // namespace

int lg_tc_wgrad_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  WgParams p;
  return plan_wgrad(Nimg, Hb, Wb, A, A, B, s, &p) ? 1 : 0;
}

int lg_tc_wgrad(const void* big, const void* small, float* dW, int Nimg, int Hb, int Wb, int A, int B, int s,
                cudaStream_t st) {
  return lg_tc_wgrad_padded(big, small, dW, Nimg, Hb, Wb, A, A, B, s, st);
}

int lg_tc_wgrad_padded(const void* big, const void* small, float* dW, int Nimg, int Hb, int Wb, int A, int A_real,
                       int B, int s, cudaStream_t st) {
  WgParams p;
  if (!plan_wgrad(Nimg, Hb, Wb, A, A_real, B, s, &p)) {
    lg_set_error("tcgen05 wgrad: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  p.dW = dW;
  CUtensorMap tmBig, tmSmall;
  const CUtensorMapSwizzle sw_a = p.a_blk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : p.a_blk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  int e = tc_host::encode_act_map(&tmBig, big, Nimg, Hb, Wb, A, p.a_blk, p.BW, p.BH, p.BN, s, sw_a);
  if (e) return e;
  e = tc_host::encode_act_map(&tmSmall, small, Nimg, p.Hs, p.Ws, B, p.b_blk, p.BW, p.BH, p.BN, 1,
                              p.b_blk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (e) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const size_t shm = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  const int tiles = p.groups * p.a_tiles * p.n_tiles;
  const int slices = (p.total_kb + p.kb_per_slice - 1) / p.kb_per_slice;
  tc_wgrad_kernel<<<dim3(tiles, slices), NUM_THREADS, shm, st>>>(tmBig, tmSmall, p);
  return LG_OK;
}
