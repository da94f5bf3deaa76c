// Row-streaming stride-2 transposed convolution for the decoder's last Conv2DTranspose(32, k5, s2)
// (model.py:39; dec4: 64x64x64 -> 128x128x32, the largest activation of the model).  The generic four-phase
// dgrad kernel fetches a shifted TMA box per tap; here every INPUT row is fetched once into a shared-memory
// ring ([pixel][128 B] exactly as TMA writes it with SWIZZLE_128B) and each of the 25 taps reads it in place
// through a K-major descriptor whose start address is shifted by whole pixels (the swizzle is a function of
// the absolute shared-memory address, so any 128-byte row is a valid start).
//
//   out[2i+py][2j+px][a] = bias[a] + sum_{ky in K(py), kx in K(px)} sum_b small[i+dy(ky)][j+dx(kx)][b] W[ky][kx][a][b]
//   K(0) = {1,3}, K(1) = {0,2,4}, dy(k) = dx(k) = 1 - (k+1)/2
//
// One input row pair position i = "row pair": 4 output phases x 32 channels = 128 TMEM columns, M = 64 pixels
// (lanes 0-15 of every quadrant); the next row pair goes to lanes 16-31 (TMEM lane offset 16), so one TMEM
// stage = 2 row pairs = 4 output rows and the epilogue's 32-lane loads are fully used.  Taps kx = 1|2 and 3|4
// read the SAME input pixel (dx = 0 resp. -1) for the two x phases, so each pair is ONE N = 64 MMA whose weight
// operand stacks both taps (columns [px 0 | px 1]); kx = 0 is an N = 32 MMA: 60 MMAs per row pair instead of
// 100 (the single-thread issue rate, not the tensor pipe, bounds these small MMAs).  The 100 KB weight operand
// is resident in shared memory.
// Epilogue (8 warps, two groups alternating stages): + bias, per-sample sum / sum of squares (InstanceNorm
// statistics), bf16 store - the two x phases of a thread are 128 contiguous bytes.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int RG_THREADS = 320;          // warps 0-7 epilogue, 8 row producer, 9 MMA issuer
constexpr int RG_WPROD = 8, RG_WMMA = 9;
constexpr int RG_ACC = 4;                // TMEM stages of 128 columns
constexpr int RG_WS = 64;                // input row width = MMA M
constexpr int RG_B = 64, RG_A = 32;      // input / output channels
constexpr int RG_ROWB = (RG_WS + 8) * RG_B * 2;      // 72 pixels x 128 B
constexpr int RG_WBYTES = 25 * (RG_B / 16) * RG_A * 16 * 2;   // 100 KB: [tap][k step][a][16 b]
constexpr int RG_RING = 8;

struct RgParams {
  int Nimg, Hs, R, strips_per_img, total_strips;   // R input rows (= row pairs) per strip
  const uint4* wpack;
  const float* bias;
  bf16* out;                             // [N,2Hs,128,32]
  double* stats;
};

__host__ __device__ constexpr int rg_d(int k) { return 1 - (k + 1) / 2; }       // input offset of tap k

// fp32 W[ky][kx][a][b] -> bf16 operand blocks [ky][shift][ks][n][16 b], no-swizzle K-major cores (8 n x 8 b = 128 B):
// shift 0 (dx = 0): n = (px, a) = taps kx 1 | 2;  shift 1 (dx = -1): taps kx 3 | 4;  shift 2 (dx = +1): n = a, tap kx 0.
__global__ void rg_pack_kernel(const float* __restrict__ W, bf16* __restrict__ out) {
  const int total = 25 * RG_A * RG_B;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int b = e % RG_B, a = (e / RG_B) % RG_A, tap = e / (RG_A * RG_B);
    const int ky = tap / 5, kx = tap - 5 * ky;
    const int shift = kx == 0 ? 2 : (kx - 1) >> 1, half = kx == 0 ? 0 : (kx - 1) & 1;
    const int n = half * RG_A + a, ks = b >> 4, k = b & 15;
    const int blk = shift == 2 ? 512 : 1024;                       // elements per (ks) block
    const int off = ky * 10240 + shift * 4096 + ks * blk + (n >> 3) * 128 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7);
    out[off] = __float2bfloat16_rn(W[e]);
  }
}

__global__ void __launch_bounds__(RG_THREADS, 1)
tc_rowdgrad_kernel(const __grid_constant__ CUtensorMap tmIn, const RgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sRing = smem;                                          // RING x 9 KB (1024-aligned: SWIZZLE_128B)
  uint8_t* sW = sRing + RG_RING * RG_ROWB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + RG_WBYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + RG_RING;
  uint64_t* tfull = empty + RG_RING;
  uint64_t* tempty = tfull + RG_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + RG_ACC);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int e0 = 0; e0 < (RG_WBYTES >> 4); e0 += 4 * RG_THREADS) {
    uint4 wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * RG_THREADS + (int)threadIdx.x;
      wv[i] = e < (RG_WBYTES >> 4) ? __ldg(p.wpack + e) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * RG_THREADS + (int)threadIdx.x;
      if (e < (RG_WBYTES >> 4)) reinterpret_cast<uint4*>(sW)[e] = wv[i];
    }
  }
  if (threadIdx.x < RG_A) sbias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmIn);
    for (int i = 0; i < RG_RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < RG_ACC; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == RG_WMMA) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nin = p.R + 2;                   // input rows a strip walks: I0-1 .. I0+R (virtual outside the image)

  if (warp == RG_WPROD) {
    if (tc::elect_one()) {
      const uint32_t ring_addr = tc::smem_u32(sRing), full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      int slot = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, I0 = (t - n * p.strips_per_img) * p.R;
        for (int k = 0; k < nin; ++k) {
          const int y = I0 - 1 + k;
          tc::mbar_wait_addr(empty_addr + slot * 8, phase ^ 1);
          if (y >= 0 && y < p.Hs) {
            tc::mbar_expect_tx_addr(full_addr + slot * 8, (uint32_t)RG_ROWB);
            tc::tma_load_3d_addr(ring_addr + slot * RG_ROWB, &tmIn, full_addr + slot * 8, 0, -1, n * p.Hs + y);
          } else {
            tc::mbar_arrive(&full[slot]);
          }
          if (++slot == RG_RING) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == RG_WMMA) {
    if (tc::elect_one()) {
      const uint32_t idesc64 = tc::make_idesc(64, 2 * RG_A, 0, 0), idesc32 = tc::make_idesc(64, RG_A, 0, 0);
      const uint32_t ring_lo = tc::smem_u32(sRing) >> 4;
      const uint32_t full_addr = tc::smem_u32(full), empty_addr = tc::smem_u32(empty);
      const uint32_t tfull_addr = tc::smem_u32(tfull), tempty_addr = tc::smem_u32(tempty);
      const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);                  // SWIZZLE_128B: SBO = 8 x 128-byte rows
      const uint32_t b_hi = (256u >> 4) | (1u << 14);                                // no swizzle: SBO = 2 cores
      const uint32_t b_lo0 = (tc::smem_u32(sW) >> 4) | ((128u >> 4) << 16);          // LBO = 128 B
      int s0 = 0;                          // ring slot of the current row pair's top input row (i - 1)
      int wslot = 0; uint32_t wphase = 0;
      int acc = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x) {
        const int n = t / p.strips_per_img, I0 = (t - n * p.strips_per_img) * p.R;
        (void)n;
        for (int r = 0; r < p.R; r += 2) {
          tc::mbar_wait_addr(tempty_addr + acc * 8, aphase ^ 1);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int need = (r + sub == 0) ? 3 : 1;
            for (int k = 0; k < need; ++k) {
              tc::mbar_wait_addr(full_addr + wslot * 8, wphase);
              if (++wslot == RG_RING) { wslot = 0; wphase ^= 1; }
            }
            tc::fence_after_sync();
            const int i = I0 + r + sub;
            const bool last = (r + sub == p.R - 1);
            const int sb = s0;
            const uint32_t tacc = tmem_base + (uint32_t)(acc * 128) + ((uint32_t)(16 * sub) << 16);
            uint32_t started0 = 0, started1 = 0;              // has output-row parity py an accumulated term yet?
#pragma unroll
            for (int d = 0; d < 3; ++d) {                      // input row i + d - 1
              const int y = i + d - 1;
              if (y >= 0 && y < p.Hs) {
                int sl = sb + d;
                if (sl >= RG_RING) sl -= RG_RING;
                const uint32_t sa_lo = (ring_lo + (uint32_t)sl * (uint32_t)(RG_ROWB >> 4)) | (1u << 16);
#pragma unroll
                for (int ky = 0; ky < 5; ++ky) {
                  if (rg_d(ky) != d - 1) continue;
                  const int py = (ky & 1) ? 0 : 1;
                  const uint32_t st = py ? started1 : started0;
                  const uint32_t tcol = tacc + (uint32_t)(py * 64);
                  const uint32_t wb = b_lo0 + (uint32_t)(ky * 1280);            // 20 KB of operand blocks per ky
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)                               // dx = 0: taps kx 1 | 2, both x phases
                    tc::mma_bf16_lohi(tcol, sa_lo + (uint32_t)(8 + ks * 2), a_hi, wb + (uint32_t)(ks * 128), b_hi,
                                      idesc64, ks == 0 ? st : 1u);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)                               // dx = -1: taps kx 3 | 4
                    tc::mma_bf16_lohi(tcol, sa_lo + (uint32_t)(ks * 2), a_hi, wb + (uint32_t)(512 + ks * 128), b_hi,
                                      idesc64, 1u);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)                               // dx = +1: tap kx 0, x phase 1 only
                    tc::mma_bf16_lohi(tcol + 32u, sa_lo + (uint32_t)(16 + ks * 2), a_hi, wb + (uint32_t)(1024 + ks * 64),
                                      b_hi, idesc32, 1u);
                  if (py) started1 = 1; else started0 = 1;
                }
              }
              if (d == 0 && !last) {                           // row i - 1 is not needed by any later row pair
                tc::mma_commit_addr(empty_addr + s0 * 8);
                if (++s0 == RG_RING) s0 = 0;
              }
            }
            if (last) {
              for (int k = 0; k < 3; ++k) {
                tc::mma_commit_addr(empty_addr + s0 * 8);
                if (++s0 == RG_RING) s0 = 0;
              }
            }
          }
          tc::mma_commit_addr(tfull_addr + acc * 8);
          if (++acc == RG_ACC) { acc = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3, grp = warp >> 2;
    const uint32_t sbias_u32 = tc::smem_u32(sbias);
    const int sub = lane >> 4, j = q * 16 + (lane & 15);
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int nstage = p.R / 2;
    const int Hb = 2 * p.Hs;
    uint32_t ps = 0;
    for (int t = blockIdx.x; t < p.total_strips; t += gridDim.x, ps += (uint32_t)nstage) {
      const int n = t / p.strips_per_img, I0 = (t - n * p.strips_per_img) * p.R;
      float s1 = 0.f, s2 = 0.f;
      for (int sg = grp; sg < nstage; sg += 2) {
        const uint32_t seq = ps + (uint32_t)sg;
        const int acc = seq & (RG_ACC - 1);
        const uint32_t aphase = (seq / RG_ACC) & 1u;
        const int i = I0 + 2 * sg + sub;
        tc::mbar_wait(&tfull[acc], aphase);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + tlane + (uint32_t)(acc * 128);
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          bf16* orow = p.out + ((((int64_t)n * Hb + 2 * i + py) * 128) + 2 * j) * RG_A;
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            float v[32];
            tc::tmem_ld32(taddr + (uint32_t)((py * 2 + px) * 32), v);
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 sb4 = tc::lds_v4(sbias_u32 + 4u * (uint32_t)e);     // shared window, not a generic LD.E
              const float a0 = v[e] + sb4.x, a1 = v[e + 1] + sb4.y, a2 = v[e + 2] + sb4.z, a3 = v[e + 3] + sb4.w;
              s1 += (a0 + a1) + (a2 + a3);
              s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2))));
              __nv_bfloat162 h0 = __floats2bfloat162_rn(a0, a1), h1 = __floats2bfloat162_rn(a2, a3);
              pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h0);
              pk[(e >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
            }
            uint4* dst = reinterpret_cast<uint4*>(orow + px * RG_A);
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      }
      if (p.stats != nullptr) {
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { atomicAdd(&p.stats[2 * n], (double)s1); atomicAdd(&p.stats[2 * n + 1], (double)s2); }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == RG_WMMA) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

bool plan_rg(int Nimg, int Hb, int Wb, int A, int B, int s, RgParams* p, int* grid) {
  if (s != 2 || A != RG_A || B != RG_B || Wb != 2 * RG_WS || Hb < 16 || (Hb & (Hb - 1))) return false;
  const int Hs = Hb / 2;
  const int ctas = lg_num_sms();
  int bestR = 0; double best = -1.0;
  for (int R = Hs; R >= 4; R >>= 1) {
    const int tiles = Nimg * (Hs / R);
    const int waves = (tiles + ctas - 1) / ctas;
    const double eff = (double)tiles / ((double)waves * ctas) * (1.0 - 0.25 * 2.0 / (R + 2.0));
    if (eff > best) { best = eff; bestR = R; }
  }
  if (const char* e = getenv("LG_RG_R")) { const int R = atoi(e); if (R >= 4 && Hs % R == 0 && R % 2 == 0) bestR = R; }
  p->Nimg = Nimg; p->Hs = Hs; p->R = bestR; p->strips_per_img = Hs / bestR; p->total_strips = Nimg * p->strips_per_img;
  *grid = lg_even_grid(p->total_strips, ctas);
  return true;
}

}  // namespace

int lg_tc_rowdgrad_supported(int Nimg, int Hb, int Wb, int A, int B, int s) {
  RgParams p; int grid;
  return plan_rg(Nimg, Hb, Wb, A, B, s, &p, &grid) ? 1 : 0;
}

// Packed weight operand; returns its size in bytes when W or wpack is NULL.
int lg_tc_rowdgrad_pack(const float* W, void* wpack, int A, int B, cudaStream_t st) {
  if (A != RG_A || B != RG_B) { lg_set_error("row-streaming dgrad: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
  if (W == nullptr || wpack == nullptr) return RG_WBYTES;
  rg_pack_kernel<<<(25 * RG_A * RG_B + 255) / 256, 256, 0, st>>>(W, (bf16*)wpack);
  return LG_OK;
}

int lg_tc_rowdgrad(const void* small, const void* wpack, const float* bias, void* out, double* stats, int Nimg, int Hb,
                   int Wb, int A, int B, int s, cudaStream_t st) {
  RgParams p; int grid;
  if (!plan_rg(Nimg, Hb, Wb, A, B, s, &p, &grid) || wpack == nullptr) {
    lg_set_error("row-streaming dgrad: unsupported geometry");
    return LG_ERR_UNSUPPORTED;
  }
  p.wpack = (const uint4*)wpack; p.bias = bias; p.out = (bf16*)out; p.stats = stats;
  tc_host::EncodeTiledFn enc = tc_host::get_encode();
  if (!enc) { lg_set_error("cuTensorMapEncodeTiled entry point not available"); return LG_ERR_CUDA; }
  CUtensorMap tmIn;
  cuuint64_t dims[3] = {(cuuint64_t)B, (cuuint64_t)RG_WS, (cuuint64_t)Nimg * p.Hs};
  cuuint64_t strides[2] = {(cuuint64_t)B * 2, (cuuint64_t)RG_WS * B * 2};
  cuuint32_t box[3] = {(cuuint32_t)B, (cuuint32_t)(RG_WS + 8), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(small), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { lg_set_error("row-streaming dgrad: tensor map failed: %d", (int)r); return LG_ERR_CUDA; }
  const size_t shm = (size_t)RG_RING * RG_ROWB + RG_WBYTES + 1024 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_rowdgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  tc_rowdgrad_kernel<<<grid, RG_THREADS, shm, st>>>(tmIn, p);
  return LG_OK;
}
