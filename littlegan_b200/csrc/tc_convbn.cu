// General conv + folded batch-norm + ReLU unit of the Inception-2015 pool_3 graph (fid.py:36-106) on tcgen05:
// an implicit GEMM  y[pos][co] = sum_k patch[pos][k] * W[k][co]  with  k = (ky, kx, c), any kh x kw / stride / padding /
// odd map sizes (299, 149, 147, 73, 71, 35, 17, 8 do not tile into TMA boxes, so the patch matrix is gathered).
//
//   * 16 builder warps in 4 independent groups gather the A operand, each group filling a different ring item: thread
//     (position row, chunk) copies 16-byte runs (8 channels of one tap - channel counts are multiples of 8) straight
//     from the NHWC map into the no-swizzle K-major core-matrix layout [k-chunk of 8][position][16 B] with cp.async;
//     out-of-image taps and the K / M tails use its zero-fill form.  Tile setup (positions -> image coordinates, tap
//     validity bit masks) is division-free and shared across the 8 chunk lanes by shuffles.  The item's weight block,
//     prepacked in exactly its shared-memory image, arrives by ONE bulk copy (TMA unit) on the same full barrier.
//   * one elected thread issues 128 x Ncols x 16 tcgen05.mma over a 4-8 stage mbarrier ring into one of two TMEM
//     accumulators (Ncols <= 256, Cout > 256 is split into equal column tiles);
//   * 4 epilogue warps read the accumulator (tcgen05.ld), apply scale / shift / ReLU and write bf16 rows into the
//     unit's channel slice of the block's concat buffer - while the next tile's MMAs run into the other accumulator.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"

namespace {

constexpr int KC = 64;                      // k per stage: 8 chunks of 8
constexpr int PLANE = 128 * 16 + 16;        // one k-chunk plane of A: 128 positions x 16 B, pitched against bank conflicts
constexpr int A_STAGE = (KC / 8) * PLANE;   // ~16 KB
constexpr int BUILDERS = 512;               // 16 gather warps
constexpr int GROUPS = 4;                   // independent builder groups (different ring items in flight)
constexpr int GROUP_THREADS = BUILDERS / GROUPS;
constexpr int PPT = 1024 / GROUP_THREADS;   // positions per builder thread per item
static_assert(PPT == 8, "the tile setup deals one position to each of the 8 chunk lanes");
constexpr int THREADS = BUILDERS + 32 + 128;
constexpr int MMA_WARP = BUILDERS / 32;     // then 4 epilogue warps (warp & 3 = TMEM lane quarter)
constexpr int MAX_STAGES = 8;

__device__ __forceinline__ void cb_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int cb_div(int n, uint32_t mul, uint32_t shr) {   // n / d, (mul, shr) from fast_div
  return mul ? (int)(__umulhi((uint32_t)n, mul) >> shr) : n;
}
// bits k in [0, taps) with 0 <= o + k < size
__device__ __forceinline__ uint32_t cb_tap_mask(int o, int taps, int size) {
  const int lo = max(0, -o), hi = min(taps, size - o);
  return hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
}
__device__ __forceinline__ bool cb_mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking phase test
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// 1-D bulk copy global -> shared by the TMA unit, completing `bytes` on an mbarrier
__device__ __forceinline__ void cb_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct CBParams {
  const bf16* x; const bf16* wpack; const float* scale; const float* shift; bf16* y;
  int H, W, Cin, xs, Ho, Wo, kh, kw, s, ph, pw, ys, relu;
  int M, K, Cout, Ncols, n_tiles, m_tiles, KS, stages, b_stage;   // b_stage = Ncols * KC * 2 bytes
  int HoWo;
  int dbg;       // LG_CONVBN_DBG knock-outs (profiling only; scripts/convbn_knockout.sh): 1 no A copies, 2 no MMAs,
                 // 4 no epilogue, 8 no B copies, 64 plain arrive instead of tcgen05.commit
  uint32_t hw_mul, hw_shr, wo_mul, wo_shr, nt_mul, nt_shr;         // / (Ho*Wo), / Wo, / n_tiles as multiply-high + shift
  uint32_t ks_mul, ks_shr, ci_mul, ci_shr, kw_mul, kw_shr;         // / KS, / Cin, / kw
};

__global__ void __launch_bounds__(THREADS, 1) convbn_kernel(const CBParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.stages * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.stages * p.b_stage);
  uint64_t* full = bars;                      // [MAX_STAGES] builders -> MMA   (the 4 warps of one builder group + the weight block's bytes)
  uint64_t* empty = bars + MAX_STAGES;        // [MAX_STAGES] MMA -> builders
  uint64_t* tfull = bars + 2 * MAX_STAGES;    // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;               // [2] epilogue -> MMA            (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sscale = reinterpret_cast<float*>(tmem_slot + 4);
  float* sshift = sscale + p.Cout;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.Ncols) tmem_cols <<= 1;

  for (int e = threadIdx.x; e < p.Cout; e += THREADS) { sscale[e] = p.scale[e]; sshift[e] = p.shift[e]; }
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) { tc::mbar_init(&full[i], GROUP_THREADS / 32 + 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < MMA_WARP) {
    // ------------------------------------------------------------------ builders
    // The ring items (tile, k stage) are dealt to GROUPS independent groups of 4 warps: one pass of the stage loop is a
    // chain of long-latency synchronisation instructions (barrier test, vote, cp.async wait, proxy fence, arrive:
    // ~1000 cycles measured with every copy and MMA knocked out), so the groups work on different items at the same
    // time instead of all warps stepping through every stage together (item i goes to group i % GROUPS).
    // Every copy is asynchronous: the A runs by cp.async (zero-fill form for padding and tails), the item's weight
    // block by one bulk copy that completes on the item's full barrier.
    // Lane mapping: the 8 chunks of a position are 8 consecutive lanes (channel / tap runs that are contiguous in
    // memory), a warp covers 4 neighbouring positions - a warp-wide copy touches 4 full 128-byte lines instead of 32
    // partial ones.  Planes are pitched 2048 + 16 bytes so that those 32 stores spread over all banks.
    const int group = threadIdx.x / GROUP_THREADS, gtid = threadIdx.x - group * GROUP_THREADS;
    const int chunk = gtid & 7, prow = gtid >> 3;               // prow: 0 .. GROUP_THREADS/8 - 1
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_items = my_tiles * p.KS;
    const int smask = p.stages - 1, sshift_ = p.stages == 8 ? 3 : 2;      // ring sizes are 4 or 8
    uint32_t vyx[PPT];                                         // bits 0-7: in-image tap rows, bits 8-15: tap columns
    const bf16* base[PPT];                                     // &x[n, y0, x0, 0] (may lie outside; used with valid taps)
    int cur_tile = -1;
    const uint8_t* wsrc = nullptr;
    int astage = 0, pending = 0;                               // oldest item this warp has not arrived for yet
    // item i -> group i % GROUPS: with GROUPS dividing the ring size every slot is always filled by the same group, in
    // order, so the parity wait on its empty barrier can never be a whole ring generation off
    {
      for (int item = group; item < total_items; item += GROUPS) {
        const int tl = cb_div(item, p.ks_mul, p.ks_shr), ks = item - tl * p.KS;
        const int stage = item & smask;
        const uint32_t phase = (uint32_t)(item >> sshift_) & 1u;
        if (tl != cur_tile) {
          cur_tile = tl;
          const int t = (int)blockIdx.x + tl * (int)gridDim.x;
          const int mt = p.n_tiles == 1 ? t : cb_div(t, p.nt_mul, p.nt_shr), nt = t - mt * p.n_tiles;
          wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (int64_t)nt * p.KS * p.b_stage;
          // the 8 chunk lanes of a position row need the same 8 positions: lane `chunk` sets up position `chunk`,
          // the values travel by shuffles (an eighth of the integer work per thread)
          uint32_t my_vyx = 0;                                 // rows past M: no tap is valid
          int my_off = 0;                                      // element offset of &x[n, y0, x0, 0] (fits 32 bits)
          {
            const int gm = mt * 128 + prow + (GROUP_THREADS / 8) * chunk;
            if (gm < p.M) {
              const int n = cb_div(gm, p.hw_mul, p.hw_shr);
              const int r = gm - n * p.HoWo;
              const int oy = cb_div(r, p.wo_mul, p.wo_shr), ox = r - oy * p.Wo;
              const int y0 = p.s * oy - p.ph, x0 = p.s * ox - p.pw;
              my_vyx = cb_tap_mask(y0, p.kh, p.H) | (cb_tap_mask(x0, p.kw, p.W) << 8);
              my_off = ((n * p.H + y0) * p.W + x0) * p.xs;
            }
          }
#pragma unroll
          for (int q = 0; q < PPT; ++q) {
            const int src_lane = (lane & ~7) | q;
            vyx[q] = __shfl_sync(0xffffffffu, my_vyx, src_lane);
            base[q] = p.x + __shfl_sync(0xffffffffu, my_off, src_lane);
          }
        }
        // Slot not free yet = the tensor core is behind: hand over everything issued so far before blocking.
        // (Warp-uniform decision: lane 0 arrives for the warp.)
        if (!__all_sync(0xffffffffu, cb_mbar_test(&empty[stage], phase ^ 1))) {
          if (pending) {
            cb_cp_async_wait<0>();
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[astage]);
            pending = 0;
          }
          tc::mbar_wait(&empty[stage], phase ^ 1);
        }
        if (gtid == 0 && (p.dbg & 8)) tc::mbar_arrive(&full[stage]);
        if (gtid == 0 && !(p.dbg & 8)) {
          tc::mbar_expect_tx(&full[stage], (uint32_t)p.b_stage);
          cb_bulk_load(tc::smem_u32(sB + stage * p.b_stage), wsrc + (int64_t)ks * p.b_stage, (uint32_t)p.b_stage,
                       tc::smem_u32(&full[stage]));
        }
        // (ky, kx, c) of this thread's chunk: k = ks * 64 + chunk * 8 = (ky * kw + kx) * Cin + c
        const int k0 = ks * KC + chunk * 8;
        const int tap = cb_div(k0, p.ci_mul, p.ci_shr), c = k0 - tap * p.Cin;
        const int ky = cb_div(tap, p.kw_mul, p.kw_shr), kx = tap - ky * p.kw;     // ky >= kh in the K tail: no mask bit
        const int tapoff = (ky * p.W + kx) * p.xs + c;
        const uint32_t a_dst = tc::smem_u32(sA + stage * A_STAGE + chunk * PLANE + prow * 16);
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          const bool ok = ((vyx[q] >> ky) & (vyx[q] >> (8 + kx)) & 1u) != 0;
          const bf16* src = ok ? base[q] + tapoff : p.x;
          if (!(p.dbg & 1)) cb_cp_async16(a_dst + q * (GROUP_THREADS / 8) * 16, src, ok ? 16u : 0u);
        }
        cb_cp_async_commit();
        if (pending) {                                         // one item of run-ahead per warp: hand over the previous
          cb_cp_async_wait<1>();
          tc::fence_proxy_async();                             // generic-proxy writes -> async (tensor core) proxy
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full[astage]);
        }
        astage = stage; pending = 1;
      }
    }
    if (pending) {
      cb_cp_async_wait<0>();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full[astage]);
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, p.Ncols, 0, 0);
      const uint64_t desc_a0 = tc::make_sdesc(tc::smem_u32(sA), PLANE, 128, 0u);             // LBO = plane, SBO = 8 rows
      const uint64_t desc_b0 = tc::make_sdesc(tc::smem_u32(sB), 128, (KC / 8) * 128, 0u);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        tc::mbar_wait(&tempty[acc], aphase ^ 1);
        tc::fence_after_sync();
        for (int ks = 0; ks < p.KS; ++ks) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          // descriptors = ring-slot-0 descriptor + (byte offset >> 4): the address field is the low 14 bits
          const uint64_t da0 = desc_a0 + (uint64_t)((stage * A_STAGE) >> 4);
          const uint64_t db0 = desc_b0 + (uint64_t)((stage * p.b_stage) >> 4);
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            const uint64_t da = da0 + (uint64_t)((kk * 2 * PLANE) >> 4);
            const uint64_t db = db0 + (uint64_t)((kk * 256) >> 4);
            if (!(p.dbg & 2)) tc::mma_bf16(tmem_base + acc * p.Ncols, da, db, idesc, (ks | kk) != 0);
          }
          if (p.dbg & 64) tc::mbar_arrive(&empty[stage]); else tc::mma_commit(&empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (p.dbg & 64) tc::mbar_arrive(&tfull[acc]); else tc::mma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t sscale_u32 = tc::smem_u32(sscale), sshift_u32 = tc::smem_u32(sshift);
    int acc = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mt = p.n_tiles == 1 ? t : cb_div(t, p.nt_mul, p.nt_shr), nt = t - mt * p.n_tiles;
      const int gm = mt * 128 + row;
      const int c0 = nt * p.Ncols;
      bf16* orow = p.y + (int64_t)gm * p.ys + c0;
      tc::mbar_wait(&tfull[acc], aphase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Ncols);
      for (int cb = 0; cb < ((p.dbg & 4) ? 0 : p.Ncols); cb += 16) {
        float v[16];
        tc::tmem_ld16(taddr + cb, v);                          // warp-collective: every lane, also past the M tail
        uint32_t pk[8];
        // folded-BN scale / shift from shared memory through the shared window, 16 bytes at a time (behind the generic
        // pointers these were 32 scalar LD.E per chunk and thread)
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 sc = tc::lds_v4(sscale_u32 + 4u * (uint32_t)(c0 + cb + e));
          const float4 sh = tc::lds_v4(sshift_u32 + 4u * (uint32_t)(c0 + cb + e));
          float a0 = fmaf(v[e], sc.x, sh.x), a1 = fmaf(v[e + 1], sc.y, sh.y);
          float a2 = fmaf(v[e + 2], sc.z, sh.z), a3 = fmaf(v[e + 3], sc.w, sh.w);
          if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
          __nv_bfloat162 h0 = __floats2bfloat162_rn(a0, a1), h1 = __floats2bfloat162_rn(a2, a3);
          pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          pk[(e >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        if (gm < p.M) {
          uint4* dst = reinterpret_cast<uint4*>(orow + cb);
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aphase ^= 1; }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// W fp32 [K][Cout] -> per (column tile, k stage) the bf16 shared-memory image of the B operand:
// no-swizzle K-major core matrices, byte offset (b>>3)*1024 + (k>>3)*128 + (b&7)*16 + (k&7)*2 inside a stage block
__global__ void convbn_pack_kernel(const float* __restrict__ W, bf16* __restrict__ out, int K, int Cout, int Ncols,
                                   int n_tiles, int KS) {
  const int64_t total = (int64_t)n_tiles * KS * Ncols * KC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kl = (int)(e % KC);
    int64_t t = e / KC;
    const int b = (int)(t % Ncols); t /= Ncols;
    const int ks = (int)(t % KS);
    const int nt = (int)(t / KS);
    const int k = ks * KC + kl;
    const float v = k < K ? W[(int64_t)k * Cout + nt * Ncols + b] : 0.f;
    const int64_t block = ((int64_t)nt * KS + ks) * Ncols * KC;
    const int off = (b >> 3) * 512 + (kl >> 3) * 64 + (b & 7) * 8 + (kl & 7);      // in bf16 elements
    out[block + off] = __float2bfloat16_rn(v);
  }
}

// q = n / d for 0 <= n < 2^31 as (n * mul) >> (32 + shr)
inline void fast_div(uint32_t d, uint32_t* mul, uint32_t* shr) {
  if (d == 1) { *mul = 0; *shr = 0; return; }                  // mul 0 = "divisor is 1" (handled in cb_div)
  uint32_t l = 0;
  while ((1u << l) < d) ++l;                                   // ceil(log2 d)
  const uint64_t p2 = 1ull << (31 + l);
  *mul = (uint32_t)((p2 + d - 1) / d);
  *shr = l - 1;
}

struct Plan { int Ncols, n_tiles, KS; };
inline bool plan(int Cin, int kh, int kw, int Cout, Plan* pl) {
  if (Cin % 8 != 0 || Cout % 16 != 0) return false;
  const int nt = (Cout + 255) / 256;
  if (Cout % (16 * nt) != 0) return false;
  pl->n_tiles = nt; pl->Ncols = Cout / nt; pl->KS = (kh * kw * Cin + KC - 1) / KC;
  return true;
}

}  // namespace

int lg_tc_convbn_supported(int Cin, int x_stride, int x_off, int kh, int kw, int Cout, int y_stride, int y_off) {
  Plan pl;
  return plan(Cin, kh, kw, Cout, &pl) && x_stride % 8 == 0 && x_off % 8 == 0 && y_stride % 8 == 0 && y_off % 8 == 0;
}

int64_t lg_tc_convbn_pack_bytes(int Cin, int kh, int kw, int Cout) {
  Plan pl;
  if (!plan(Cin, kh, kw, Cout, &pl)) return 0;
  return (int64_t)pl.n_tiles * pl.KS * pl.Ncols * KC * 2;
}

int lg_tc_convbn_pack(const float* W, void* wpack, int Cin, int kh, int kw, int Cout, cudaStream_t st) {
  Plan pl;
  if (!plan(Cin, kh, kw, Cout, &pl)) return LG_ERR_UNSUPPORTED;
  const int64_t total = (int64_t)pl.n_tiles * pl.KS * pl.Ncols * KC;
  int g = (int)((total + 255) / 256);
  if (g > lg_num_sms() * 16) g = lg_num_sms() * 16;
  convbn_pack_kernel<<<g, 256, 0, st>>>(W, (bf16*)wpack, kh * kw * Cin, Cout, pl.Ncols, pl.n_tiles, pl.KS);
  return LG_OK;
}

int lg_tc_convbn(const void* x, const void* wpack, const float* scale, const float* shift, void* y, int N, int H, int Wd,
                 int Cin, int xs, int xo, int kh, int kw, int s, int ph, int pw, int Cout, int ys, int yo, int relu,
                 cudaStream_t st) {
  Plan pl;
  if (!plan(Cin, kh, kw, Cout, &pl)) return LG_ERR_UNSUPPORTED;
  CBParams p;
  p.x = (const bf16*)x + xo; p.wpack = (const bf16*)wpack; p.scale = scale; p.shift = shift; p.y = (bf16*)y + yo;
  p.H = H; p.W = Wd; p.Cin = Cin; p.xs = xs; p.kh = kh; p.kw = kw; p.s = s; p.ph = ph; p.pw = pw; p.ys = ys; p.relu = relu;
  p.Ho = (H + 2 * ph - kh) / s + 1; p.Wo = (Wd + 2 * pw - kw) / s + 1;
  p.M = N * p.Ho * p.Wo; p.K = kh * kw * Cin; p.Cout = Cout;
  p.Ncols = pl.Ncols; p.n_tiles = pl.n_tiles; p.KS = pl.KS; p.m_tiles = (p.M + 127) / 128;
  p.b_stage = pl.Ncols * KC * 2;
  fast_div((uint32_t)(p.Ho * p.Wo), &p.hw_mul, &p.hw_shr);
  fast_div((uint32_t)p.Wo, &p.wo_mul, &p.wo_shr);
  fast_div((uint32_t)p.n_tiles, &p.nt_mul, &p.nt_shr);
  fast_div((uint32_t)p.KS, &p.ks_mul, &p.ks_shr);
  fast_div((uint32_t)Cin, &p.ci_mul, &p.ci_shr);
  fast_div((uint32_t)kw, &p.kw_mul, &p.kw_shr);
  p.HoWo = p.Ho * p.Wo;
  static const int dbg_env = getenv("LG_CONVBN_DBG") ? atoi(getenv("LG_CONVBN_DBG")) : 0;
  p.dbg = dbg_env;
  p.stages = 8 * (A_STAGE + p.b_stage) <= 196 * 1024 ? 8 : 4;        // power of two: slot = item & (stages - 1)
  const size_t shm = (size_t)p.stages * (A_STAGE + p.b_stage) + 1024 + 256 + 2 * (size_t)Cout * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(convbn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = lg_even_grid(tiles, lg_num_sms());
  convbn_kernel<<<grid, THREADS, shm, st>>>(p);
  return LG_OK;
}
