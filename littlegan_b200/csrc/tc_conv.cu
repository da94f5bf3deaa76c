// placeholder - replaced by the tcgen05 implementation
#include "common.cuh"
#include "internal.h"
int lg_tc_supported(int, int, int, int, int, int, int) { return 0; }
int lg_tc_fprop(const void*, const void*, const float*, void*, double*, int, int, int, int, int, int, cudaStream_t) { lg_set_error("tcgen05 path: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
int lg_tc_dgrad(const void*, const void*, const float*, void*, double*, int, int, int, int, int, int, int, cudaStream_t) { lg_set_error("tcgen05 path: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
int lg_tc_wgrad(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t) { lg_set_error("tcgen05 path: unsupported geometry"); return LG_ERR_UNSUPPORTED; }
