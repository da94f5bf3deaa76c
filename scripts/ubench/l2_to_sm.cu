// L2 -> SM bandwidth ceiling on this GPU: every CTA streams an L2-resident buffer into its shared memory with
// bulk async copies (the TMA data path, cp.async.bulk + mbarrier), `depth` copies in flight per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/l2_to_sm scripts/ubench/l2_to_sm.cu
//   ./l2_to_sm [region_MB=48] [chunk_KB=16] [ctas_per_sm=1]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(
          smem_u32(b)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int DEPTH = 6;

__global__ void __launch_bounds__(64) stream_kernel(const uint8_t* __restrict__ buf, size_t region, uint32_t chunk, int iters) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[DEPTH];
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t nchunks = region / chunk;
  size_t c = ((size_t)blockIdx.x * 7919u) % nchunks;      // decorrelate the CTAs' positions
  uint32_t phase[DEPTH] = {0};
  for (int i = 0; i < DEPTH; ++i) {
    mbar_expect(&bars[i], chunk);
    bulk_g2s(smem + (size_t)i * chunk, buf + c * chunk, chunk, &bars[i]);
    c = (c + 1) % nchunks;
  }
  for (int it = DEPTH; it < iters; ++it) {
    const int s = it % DEPTH;
    mbar_wait(&bars[s], phase[s]);
    phase[s] ^= 1;
    mbar_expect(&bars[s], chunk);
    bulk_g2s(smem + (size_t)s * chunk, buf + c * chunk, chunk, &bars[s]);
    c = (c + 1) % nchunks;
  }
  for (int s = 0; s < DEPTH; ++s) mbar_wait(&bars[s], phase[s]);
}

int main(int argc, char** argv) {
  const size_t region = (size_t)(argc > 1 ? atoi(argv[1]) : 48) << 20;
  const uint32_t chunk = (uint32_t)(argc > 2 ? atoi(argv[2]) : 16) << 10;
  const int per_sm = argc > 3 ? atoi(argv[3]) : 1;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  uint8_t* buf;
  cudaMalloc(&buf, region);
  cudaMemset(buf, 1, region);
  const size_t shm = (size_t)DEPTH * chunk;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    stream_kernel<<<sms * per_sm, 64, shm>>>(buf, region, chunk, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)sms * per_sm * iters * chunk;
    printf("region %zu MB, chunk %u KB, %d CTAs/SM x %d SMs, depth %d: %.3f ms, %.2f TB/s L2->SM (%s)\n", region >> 20,
           chunk >> 10, per_sm, sms, DEPTH, ms, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
