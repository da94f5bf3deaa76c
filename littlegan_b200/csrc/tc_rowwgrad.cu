// Row-streaming weight gradient for the decoder's last Conv2DTranspose(32, k5, s2) (model.py:39; dec4):
//
//   dW[ky][kx][a][b] += sum_{n,i,j} big[n][2i+ky-1][2j+kx-1][a] * small[n][i][j][b]      big 128x128x32, small 64x64x64
//
// The generic wgrad kernel (tc_wgrad.cu) re-fetches the big map once per tap (25 shifted TMA boxes per k-block)
// and is bound by L2->SM bandwidth on this layer (24 KB per 128 MMA cycles).  Here a CTA walks a strip of small-map
// rows; every big-map row is fetched ONCE into a shared-memory ring as PIXEL PAIRS ([pair][128 B] = even pixel |
// odd pixel, SWIZZLE_128B: one 128-byte TMA run per pair, and the stride-2 gather becomes a unit-stride walk over
// pairs) and every small-map row once ([pixel][128 B], SWIZZLE_128B).  Both are then MN-major tcgen05 operands as
// they lie (K = the 64 positions of the row, 4 MMAs of K = 16):
//   * a 64-element M block of the A operand is one pixel pair = two taps (x = 2(j+d), 2(j+d)+1 <-> kx = 2d+1, 2d+2);
//     consecutive pairs are consecutive M blocks (LBO = 128 B), so ONE M = 128 MMA covers the taps kx = 1..4 of a
//     kernel row and an M = 64 MMA on the pair before covers kx = 0 (its even-pixel half is unused);
//   * the B operand is the small row (N = 64 channels).
// All 25 tap accumulators stay in TMEM for the whole kernel (5 x 64 columns for the M = 128 groups, 3 x 64 for the
// M = 64 groups, two per column block through the TMEM lane offset 16): 512 columns.  At the end every CTA adds
// its partial dW with bulk reduce-adds (cp.reduce.async.bulk .add.f32, one 256-byte row per thread).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int RW_THREADS = 192;          // warps 0-3 final reduction, 4 TMA producer, 5 MMA issuer
constexpr int RW_A = 32, RW_B = 64, RW_WS = 64;
constexpr int RW_BIGB = (RW_WS + 8) * 2 * RW_A * 2;  // one big row: 72 pixel pairs x 128 B
constexpr int RW_SMALLB = RW_WS * RW_B * 2;          // one small row: 64 pixels x 128 B
constexpr int RW_RING = 12, RW_SRING = 4;

struct RwParams {
  int Nimg, Hb, Hs, R, strips_per_img, total_strips;
  float* dW;                             // [25][32][64] fp32, accumulated into
};

__device__ __forceinline__ void rw_red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(RW_THREADS, 1)
tc_rowwgrad_kernel(const __grid_constant__ CUtensorMap tmBig, const __grid_constant__ CUtensorMap tmSmall,
                   const RwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sBig = smem;                                           // RING x 9 KB (1024-aligned)
  uint8_t* sSmall = sBig + RW_RING * RW_BIGB;                     // SRING x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSmall + RW_SRING * RW_SMALLB);
  uint64_t* full = bars;                     // [RING]
  uint64_t* empty = full + RW_RING;          // [RING]
  uint64_t* sfull = empty + RW_RING;         // [SRING]
  uint64_t* sempty = sfull + RW_SRING;       // [SRING]
  uint64_t* done = sempty + RW_SRING;        // [1] all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmBig);
    tc::tma_prefetch_desc(&tmSmall);
    for (int i = 0; i < RW_RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < RW_SRING; ++i) { tc::mbar_init(&sfull[i], 1); tc::mbar_init(&sempty[i], 1); }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nrows = 2 * (p.R - 1) + 5;       // big rows a strip walks (rows outside the image are virtual)

  if (warp == 4) {
    // ------------------------------------------------------------ producer: big rows, then the small row they complete
    if (tc::elect_one()) {
      const uint32_t big_addr = tc::smem_u32(sBig), small_addr = tc::smem_u32(sSmall);
      const uint32_t full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      const uint32_t sfull_addr = tc::smem_u32(sfull), sempty_addr = tc::smem_u32(sempty);
      int slot = 0; uint32_t phase = 0;
      int ss = 0; uint32_t sphase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, i0 = (t - n * p.strips_per_img) * p.R;
        const int y0 = 2 * i0 - 1;
        int next_small = 0;                                       // small rows issued so far in this strip
        for (int k = 0; k < nrows; ++k) {
          const int y = y0 + k;
          tc::mbar_wait_addr(empty_addr + slot * 8, phase ^ 1);
          if (y >= 0 && y < p.Hb) {
            tc::mbar_expect_tx_addr(full_addr + slot * 8, (uint32_t)RW_BIGB);
            tc::tma_load_3d_addr(big_addr + slot * RW_BIGB, &tmBig, full_addr + slot * 8, 0, -1, n * p.Hb + y);
          } else {
            tc::mbar_arrive(&full[slot]);
          }
          if (++slot == RW_RING) { slot = 0; phase ^= 1; }
          // small row r is usable once big rows up to 2r+4 (strip-relative) are in flight
          while (next_small < p.R && 2 * next_small + 4 <= k) {
            tc::mbar_wait_addr(sempty_addr + ss * 8, sphase ^ 1);
            tc::mbar_expect_tx_addr(sfull_addr + ss * 8, (uint32_t)RW_SMALLB);
            tc::tma_load_3d_addr(small_addr + ss * RW_SMALLB, &tmSmall, sfull_addr + ss * 8, 0, 0,
                                 n * p.Hs + i0 + next_small);
            if (++ss == RW_SRING) { ss = 0; sphase ^= 1; }
            ++next_small;
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    if (tc::elect_one()) {
      const uint32_t idesc128 = tc::make_idesc(128, RW_B, 1, 1), idesc64 = tc::make_idesc(64, RW_B, 1, 1);
      const uint32_t big_lo = tc::smem_u32(sBig) >> 4, small_lo = tc::smem_u32(sSmall) >> 4;
      const uint32_t full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      const uint32_t sfull_addr = tc::smem_u32(sfull), sempty_addr = tc::smem_u32(sempty);
      // MN-major descriptors: LBO = stride between blocks along M / N, SBO = stride between 8-position groups
      const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);                  // SWIZZLE_128B, 8 x 128-byte rows
      const uint32_t b_hi = a_hi;
      constexpr uint32_t A_LBO = (128u >> 4) << 16;                                  // next 64-element block = next pixel pair
      constexpr uint32_t B_LBO = (uint32_t)(RW_SMALLB >> 4) << 16;                   // single 64-channel block: unused
      int s0 = 0;
      int wslot = 0; uint32_t wphase = 0;
      int ss = 0; uint32_t sphase = 0;
      uint32_t started = 0;                  // bit ky: group A accumulator written; bit 8+ky: group B
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, i0 = (t - n * p.strips_per_img) * p.R;
        (void)n;
        for (int r = 0; r < p.R; ++r) {
          const int need = (r == 0) ? 5 : 2;
          for (int k = 0; k < need; ++k) {
            tc::mbar_wait_addr(full_addr + wslot * 8, wphase);
            if (++wslot == RW_RING) { wslot = 0; wphase ^= 1; }
          }
          tc::mbar_wait_addr(sfull_addr + ss * 8, sphase);
          tc::fence_after_sync();
          const bool last = (r == p.R - 1);
          const int sb = s0;
          const int ytop = 2 * (i0 + r) - 1;
          const uint32_t sm_lo = (small_lo + (uint32_t)ss * (uint32_t)(RW_SMALLB >> 4)) | B_LBO;
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
            const int y = ytop + ky;
            if (y >= 0 && y < p.Hb) {
              int sl = sb + ky;
              if (sl >= RW_RING) sl -= RW_RING;
              const uint32_t row_lo = (big_lo + (uint32_t)sl * (uint32_t)(RW_BIGB >> 4)) | A_LBO;
              const uint32_t stA = (started >> ky) & 1u, stB = (started >> (8 + ky)) & 1u;
              // group A: pixel pairs j+1, j+2 of the ring row (pair p holds x = 2(p-1), 2(p-1)+1) = taps kx 1..4: M = 128
              const uint32_t tA = tmem_base + (uint32_t)(ky * 64);
              // group B: pixel pair j = tap kx 0 in its odd-pixel half: M = 64; two ky per column block
              const uint32_t tB = tmem_base + (uint32_t)(320 + (ky >> 1) * 64) + ((uint32_t)(16 * (ky & 1)) << 16);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                tc::mma_bf16_lohi(tA, row_lo + (uint32_t)(8 + ks * 128), a_hi, sm_lo + (uint32_t)(ks * 128), b_hi, idesc128,
                                  ks == 0 ? stA : 1u);
                tc::mma_bf16_lohi(tB, row_lo + (uint32_t)(ks * 128), a_hi, sm_lo + (uint32_t)(ks * 128), b_hi, idesc64,
                                  ks == 0 ? stB : 1u);
              }
              started |= (1u << ky) | (1u << (8 + ky));
            }
            if (ky == 1 && !last) {                            // big rows 2i-1 and 2i are not needed by later small rows
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                tc::mma_commit_addr(empty_addr + s0 * 8);
                if (++s0 == RW_RING) s0 = 0;
              }
            }
          }
          if (last) {
            for (int k = 0; k < 5; ++k) {
              tc::mma_commit_addr(empty_addr + s0 * 8);
              if (++s0 == RW_RING) s0 = 0;
            }
          }
          tc::mma_commit_addr(sempty_addr + ss * 8);
          if (++ss == RW_SRING) { ss = 0; sphase ^= 1; }
        }
      }
      tc::mma_commit(done);
    }
  } else {
    // ------------------------------------------------------------ final reduction of the CTA's partial dW:
    // every thread stages its TMEM lane (one (tap, a) row of 64 floats) in shared memory (the ring is idle now;
    // 272-byte pitch = conflict-free 16-byte stores) and hands the row to the TMA unit as ONE 256-byte bulk
    // reduce-add instead of sixteen scattered 16-byte atomics.
    const int q = warp;
    tc::mbar_wait(done, 0);
    tc::fence_after_sync();
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    float* stage = reinterpret_cast<float*>(sBig + (size_t)threadIdx.x * (4 * 272));   // 4 staging rows per thread
    // group A (columns ky*64): lane m = 64 d + 32 par + a -> tap kx = 2d + 1 + par;  group B (columns 320 + kp*64):
    // M = 64 rows r = 16 q + (lane & 15) = 32 par + a -> tap kx = 0 for par = 1 (par = 0 unused), lanes 16-31 hold
    // ky = 2 kp + 1
#pragma unroll 1
    for (int blk = 0; blk < 8; ++blk) {
      int ky, kx, a;
      if (blk < 5) { ky = blk; kx = 2 * (q >> 1) + 1 + (q & 1); a = lane; }
      else { const int r = 16 * q + (lane & 15); ky = 2 * (blk - 5) + (lane >> 4); kx = (r >> 5) ? 0 : 5; a = r & 31; }
      const bool valid = ky < 5 && kx < 5;
      float* st = stage + (blk & 3) * 68;
      if (blk == 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // rows 0-3 have been read
#pragma unroll
      for (int cb = 0; cb < RW_B; cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + tlane + (uint32_t)(blk * 64 + cb), v);
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(st + cb + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
      tc::fence_proxy_async();
      if (valid) tc::bulk_reduce_add_f32(p.dW + ((int64_t)((ky * 5 + kx) * RW_A + a)) * RW_B, st, RW_B * 4);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

bool plan_rw(int Nimg, int Hb, int Wb, int A, int B, int s, RwParams* p, int* grid) {
  if (s != 2 || A != RW_A || B != RW_B || Wb != 2 * RW_WS || Hb < 16 || (Hb & (Hb - 1))) return false;
  const int Hs = Hb / 2;
  const int ctas = lg_num_sms();
  int bestR = 0; double best = -1.0;
  for (int R = Hs; R >= 4; R >>= 1) {
    const int tiles = Nimg * (Hs / R);
    const int waves = (tiles + ctas - 1) / ctas;
    const double eff = (double)tiles / ((double)waves * ctas) * (1.0 - 0.15 * 3.0 / (2.0 * R + 3.0));
    if (eff > best) { best = eff; bestR = R; }
  }
  if (const char* e = getenv("LG_RW_R")) { const int R = atoi(e); if (R >= 4 && Hs % R == 0) bestR = R; }
  p->Nimg = Nimg; p->Hb = Hb; p->Hs = Hs; p->R = bestR; p->strips_per_img = Hs / bestR;
  p->total_strips = Nimg * p->strips_per_img;
  *grid = lg_even_grid(p->total_strips, ctas);
  return true;
}

}  // namespace

int lg_tc_rowwgrad_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  RwParams p; int grid;
  return plan_rw(Nimg, Hb, Wb, A, B, s, &p, &grid) ? 1 : 0;
}

int lg_tc_rowwgrad(const void* big, const void* small, float* dW, int Nimg, int Hb, int Wb, int A, int B, int s,
                   cudaStream_t st) {
  RwParams p; int grid;
  if (!plan_rw(Nimg, Hb, Wb, A, B, s, &p, &grid)) {
    lg_set_error("row-streaming wgrad: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  p.dW = dW;
  tc_host::EncodeTiledFn enc = tc_host::get_encode();
  if (!enc) { lg_set_error("cuTensorMapEncodeTiled entry point not available"); return LG_ERR_CUDA; }
  CUtensorMap tmBig, tmSmall;
  {
    cuuint64_t dims[3] = {(cuuint64_t)2 * A, (cuuint64_t)(Wb / 2), (cuuint64_t)Nimg * Hb};      // pixel pairs
    cuuint64_t strides[2] = {(cuuint64_t)2 * A * 2, (cuuint64_t)Wb * A * 2};
    cuuint32_t box[3] = {(cuuint32_t)(2 * A), (cuuint32_t)(RW_WS + 8), 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmBig, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(big), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lg_set_error("row-streaming wgrad: big tensor map failed: %d", (int)r); return LG_ERR_CUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)B, (cuuint64_t)RW_WS, (cuuint64_t)Nimg * p.Hs};
    cuuint64_t strides[2] = {(cuuint64_t)B * 2, (cuuint64_t)RW_WS * B * 2};
    cuuint32_t box[3] = {(cuuint32_t)B, (cuuint32_t)RW_WS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmSmall, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(small), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lg_set_error("row-streaming wgrad: small tensor map failed: %d", (int)r); return LG_ERR_CUDA; }
  }
  const size_t shm = (size_t)RW_RING * RW_BIGB + (size_t)RW_SRING * RW_SMALLB + 512 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_rowwgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  tc_rowwgrad_kernel<<<grid, RW_THREADS, shm, st>>>(tmBig, tmSmall, p);
  return LG_OK;
}
