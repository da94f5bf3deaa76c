// Host-side TMA descriptor helpers shared by the tcgen05 translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace tc_host {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// 4-D map over an NHWC bf16 activation tensor; box = {boxC, boxW, boxH, boxN} OUTPUT elements, taken with
// element stride `estride` along W and H (the conv stride); out-of-bounds elements read as zero.
static inline int encode_act_map(CUtensorMap* m, const void* base, int Nimg, int H, int W, int C, int boxC, int boxW,
                                 int boxH, int boxN, int estride, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { lg_set_error("cuTensorMapEncodeTiled entry point not available"); return LG_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)(boxW * estride), (cuuint32_t)(boxH * estride), (cuuint32_t)boxN};
  cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { lg_set_error("cuTensorMapEncodeTiled(activation) failed: %d", (int)r); return LG_ERR_CUDA; }
  return LG_OK;
}

}  // namespace tc_host
