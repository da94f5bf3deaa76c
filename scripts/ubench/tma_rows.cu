// How fast does one SM's TMA engine move data from L2 as a function of the box shape?
//   mode 0: cp.async.bulk (1-D, no tensor map)                         chunk bytes = rows * 128
//   mode 1: cp.async.bulk.tensor.2d, box = (64 bf16 = 128 B inner) x rows, SWIZZLE_128B
//   mode 2: cp.async.bulk.tensor.2d, box = (inner_elems) x rows, no swizzle  (inner up to 256 elems = 512 B)
//   mode 3: one SWIZZLE_128B tensor box AND one 1-D bulk copy of the same size per ring slot (do the two paths add?)
// Every CTA (one per SM) keeps DEPTH boxes in flight over an L2-resident [R x C] bf16 matrix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/tma_rows scripts/ubench/tma_rows.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

#ifndef DEPTH
#define DEPTH 6
#endif

#ifndef NPROD
#define NPROD 1
#endif
__global__ void __launch_bounds__(32 * NPROD + 32) k(const __grid_constant__ CUtensorMap tm, const uint8_t* buf, int mode, int rows, int inner_bytes,
                                        int R, int C, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_base[];
  uint8_t* smem = smem_base;
  __shared__ uint64_t bars_all[NPROD][DEPTH];
  if (threadIdx.x == 0) {
    for (int w = 0; w < NPROD; ++w)
      for (int i = 0; i < DEPTH; ++i) mbar_init(&bars_all[w][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x & 31) != 0 || (threadIdx.x >> 5) >= NPROD) return;
  const int pw = threadIdx.x >> 5;                      // producer warp: its own ring of DEPTH slots
  uint64_t* bars = bars_all[pw];
  smem += (size_t)pw * DEPTH * (((size_t)rows * inner_bytes * (mode == 3 ? 2 : 1) + 1023) & ~(size_t)1023);
  const uint32_t box_bytes = (uint32_t)rows * inner_bytes * (mode == 3 ? 2 : 1);
  const uint32_t half = (uint32_t)rows * inner_bytes;
  const int boxes_y = R / rows, boxes_x = (C * 2) / inner_bytes;
  const int nbox = boxes_y * boxes_x;
  int c = (int)((blockIdx.x * 7919u + pw * 104729u) % nbox);
  uint32_t phase[DEPTH] = {0};
  auto issue = [&](int s) {
    mbar_expect(&bars[s], box_bytes);
    uint8_t* dst = smem + (size_t)s * ((box_bytes + 1023) & ~1023u);
    if (mode == 0) bulk_g2s(dst, buf + (size_t)c * box_bytes, box_bytes, &bars[s]);
    else if (mode == 3) {
      tma_2d(dst, &tm, &bars[s], (c % boxes_x) * (inner_bytes / 2), (c / boxes_x) * rows);
      bulk_g2s(dst + half, buf + (size_t)((c * 31) % nbox) * half, half, &bars[s]);
    } else tma_2d(dst, &tm, &bars[s], (c % boxes_x) * (inner_bytes / 2), (c / boxes_x) * rows);
    c = (c + 1) % nbox;
  };
  for (int i = 0; i < DEPTH; ++i) issue(i);
  for (int it = DEPTH; it < iters; ++it) {
    const int s = it % DEPTH;
    mbar_wait(&bars[s], phase[s]);
    phase[s] ^= 1;
    issue(s);
  }
  for (int s = 0; s < DEPTH; ++s) mbar_wait(&bars[s], phase[s]);
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  const int rows = argc > 2 ? atoi(argv[2]) : 128;
  const int inner_bytes = argc > 3 ? atoi(argv[3]) : 128;
  const int R = 32768, C = 512;                 // 32 MB bf16 matrix, L2 resident
  uint8_t* buf;
  cudaMalloc(&buf, (size_t)R * C * 2);
  cudaMemset(buf, 1, (size_t)R * C * 2);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)(inner_bytes / 2), (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      (mode == 1 || mode == 3) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (mode != 0 && r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = argc > 4 ? atoi(argv[4]) : prop.multiProcessorCount;
  const size_t box_bytes = (size_t)rows * inner_bytes * (mode == 3 ? 2 : 1);
  const size_t shm = (size_t)NPROD * DEPTH * ((box_bytes + 1023) & ~(size_t)1023) + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  const int iters = 3000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<<<sms, 32 * NPROD + 32, shm>>>(tm, buf, mode, rows, inner_bytes, R, C, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double bytes = (double)sms * iters * box_bytes * NPROD;
  const double ns_per_box = best * 1e6 / iters;
  printf("prod %d depth %d mode %d  box %4d rows x %4d B = %6zu B : %.2f TB/s, %.1f GB/s/SM, %.0f ns/box, %.2f ns/row  (%s)\n", NPROD, DEPTH, mode, rows,
         inner_bytes, box_bytes, bytes / best / 1e9, bytes / best / 1e6 / sms, ns_per_box, ns_per_box / rows,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
