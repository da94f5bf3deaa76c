// extern "C" surface: argument validation, error reporting and SIMT / tcgen05 dispatch for the
// convolution family and the dense GEMM.  See include/littlegan_b200.h for the contracts.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "internal.h"

static thread_local char g_err[512] = "";

void lg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int lg_abi_version(void) { return 1; }
extern "C" const char* lg_last_error(void) { return g_err; }

extern "C" int lg_tensor_core_path_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int lg_conv2d_tc_supported(int op, int N, int Hb, int Wb, int A, int B, int stride) {
  return lg_tc_supported(op, Hb, Wb, A, B, stride, N);
}

static int check_geom(const char* fn, int N, int Hb, int Wb, int A, int B, int s, int dtype) {
  if (N <= 0 || Hb <= 0 || Wb <= 0 || A <= 0 || B <= 0 || (s != 1 && s != 2) || (Hb % s) || (Wb % s) ||
      (dtype != LG_F32 && dtype != LG_BF16)) {
    lg_set_error("%s: invalid geometry N=%d Hb=%d Wb=%d A=%d B=%d stride=%d dtype=%d", fn, N, Hb, Wb, A, B, s,
                 dtype);
    return LG_ERR_INVALID;
  }
  return LG_OK;
}

// Would lg_conv2d_fprop / lg_conv2d_dgrad (tensor-core path) accept a fused norm-backward epilogue here?
extern "C" int lg_conv2d_norm_bwd_supported(int op, int N, int Hb, int Wb, int A, int B, int stride) {
  if (!lg_tc_supported(op, Hb, Wb, A, B, stride, N)) return 0;
  if (op == LG_OP_FPROP) return B % 16 == 0 ? 1 : 0;                      // output channels = B
  if (op == LG_OP_DGRAD) return (A % 16 == 0 && !lg_tc_deconv_small_supported(N, Hb, Wb, A, B, stride)) ? 1 : 0;
  return 0;
}

extern "C" int lg_conv2d_fprop(const void* big, const float* W, const void* wpack, const float* bias,
                               void* small_out, double* stats, int N, int Hb, int Wb, int A, int B, int stride,
                               int dtype, int use_tc, const lg_norm_bwd_t* norm_bwd, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, dtype)) return e;
  LG_REQUIRE(big && small_out, "NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc) {
    LG_REQUIRE(dtype == LG_BF16 && wpack, "tcgen05 path needs LG_BF16 activations and packed weights");
    int e;
    if (W && lg_tc_cin3_supported(N, Hb, Wb, A, B, stride))              // RGB input: in-smem im2col
      e = lg_tc_cin3_fprop(big, W, bias, small_out, stats, N, Hb, Wb, B, stride, norm_bwd, st);
    else
      e = lg_tc_fprop(big, wpack, bias, small_out, stats, N, Hb, Wb, A, B, stride, norm_bwd, st);
    if (e) return e;
  } else {
    LG_REQUIRE(W, "NULL weights");
    LG_REQUIRE(!norm_bwd, "the fused norm-backward epilogue exists on the tensor-core path only");
    lg_simt_fprop(big, W, bias, small_out, stats, N, Hb, Wb, A, B, stride, dtype, st);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_fprop_rows_supported(int N, int Hb, int Wb, int A_big, int A, int B, int stride) {
  return lg_tc_rowconv_supported(N, Hb, Wb, A, A_big, B, stride);
}

extern "C" int lg_pack_rowconv_weights(const float* W, void* wpack, int A_big, int A, int B, int stride,
                                       void* stream) {
  int e = lg_tc_rowconv_pack(W, wpack, A, A_big, B, stride, (cudaStream_t)stream);
  if (e < 0 || !W || !wpack) return e;
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_fprop_rows(const void* big, int A_big, const void* wpack, const float* bias, void* small_out,
                                    double* stats, int N, int Hb, int Wb, int A, int B, int stride,
                                    const lg_norm_bwd_t* norm_bwd, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, LG_BF16)) return e;
  LG_REQUIRE(big && small_out && wpack, "NULL tensor");
  LG_REQUIRE(!(norm_bwd && stats), "norm_bwd excludes stats");
  int e = lg_tc_rowconv_fprop(big, wpack, bias, small_out, stats, N, Hb, Wb, A, A_big, B, stride, norm_bwd,
                              (cudaStream_t)stream);
  if (e) return e;
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_dgrad(const void* small, const float* W, const void* wpack, const float* bias,
                               void* big_out, double* stats, int N, int Hb, int Wb, int A, int B, int stride,
                               int act, int dtype, int use_tc, const lg_norm_bwd_t* norm_bwd, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, dtype)) return e;
  LG_REQUIRE(small && big_out, "NULL tensor");
  LG_REQUIRE(act == LG_ACT_NONE || act == LG_ACT_TANH, "unsupported activation");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc) {
    LG_REQUIRE(dtype == LG_BF16 && wpack, "tcgen05 path needs LG_BF16 activations and packed weights");
    int e;
    if (W && lg_tc_rowdeconv_supported(N, Hb, Wb, A, B, stride)) {       // final RGB layer: row streaming
      LG_REQUIRE(!norm_bwd, "no fused norm-backward epilogue on the RGB transposed-conv kernels");
      e = lg_tc_rowdeconv(small, W, bias, big_out, nullptr, stats, N, Hb, Wb, A, B, stride, act, st);
    } else if (W && lg_tc_deconv_small_supported(N, Hb, Wb, A, B, stride)) {    // RGB layers: GEMM + col2im
      LG_REQUIRE(!norm_bwd, "no fused norm-backward epilogue on the RGB transposed-conv kernel");
      e = lg_tc_deconv_small(small, W, bias, big_out, stats, N, Hb, Wb, A, B, stride, act, st);
    } else if (lg_tc_dgrad4_supported(N, Hb, Wb, A, B, stride))          // <= 128 channels: 4 phases per pass
      e = lg_tc_dgrad4(small, wpack, bias, big_out, stats, N, Hb, Wb, A, B, act, norm_bwd, st);
    else
      e = lg_tc_dgrad(small, wpack, bias, big_out, stats, N, Hb, Wb, A, B, stride, act, norm_bwd, st);
    if (e) return e;
  } else {
    LG_REQUIRE(W, "NULL weights");
    LG_REQUIRE(!norm_bwd, "the fused norm-backward epilogue exists on the tensor-core path only");
    lg_simt_dgrad(small, W, bias, big_out, stats, N, Hb, Wb, A, B, stride, act, dtype, st);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_dgrad_rows_supported(int N, int Hb, int Wb, int A, int B, int stride) {
  return lg_tc_rowdgrad_supported(N, Hb, Wb, A, B, stride);
}

extern "C" int lg_pack_rowdgrad_weights(const float* W, void* wpack, int A, int B, void* stream) {
  int e = lg_tc_rowdgrad_pack(W, wpack, A, B, (cudaStream_t)stream);
  if (e < 0 || !W || !wpack) return e;
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_dgrad_rows(const void* small, const void* wpack, const float* bias, void* big_out,
                                    double* stats, int N, int Hb, int Wb, int A, int B, int stride, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, LG_BF16)) return e;
  LG_REQUIRE(small && big_out && wpack, "NULL tensor");
  int e = lg_tc_rowdgrad(small, wpack, bias, big_out, stats, N, Hb, Wb, A, B, stride, (cudaStream_t)stream);
  if (e) return e;
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_dgrad_rgb_supported(int N, int Hb, int Wb, int A, int B, int stride) {
  return lg_tc_rowdeconv_supported(N, Hb, Wb, A, B, stride);
}

extern "C" int lg_conv2d_dgrad_rgb(const void* small, const float* W, const float* bias, void* big_out,
                                   void* big_out_pad8, double* stats, int N, int Hb, int Wb, int A, int B, int stride,
                                   int act, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, LG_BF16)) return e;
  LG_REQUIRE(small && big_out && W, "NULL tensor");
  LG_REQUIRE(act == LG_ACT_NONE || act == LG_ACT_TANH, "unsupported activation");
  int e = lg_tc_rowdeconv(small, W, bias, big_out, big_out_pad8, stats, N, Hb, Wb, A, B, stride, act,
                          (cudaStream_t)stream);
  if (e) return e;
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_wgrad(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A,
                               int B, int stride, int dtype, int use_tc, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, dtype)) return e;
  LG_REQUIRE(big && small && dW, "NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc) {
    LG_REQUIRE(dtype == LG_BF16, "tcgen05 path needs LG_BF16 activations");
    int e;
    if (lg_tc_rowwgrad_supported(N, Hb, Wb, A, B, stride))               // dec4: row streaming, all taps resident
      e = lg_tc_rowwgrad(big, small, dW, N, Hb, Wb, A, B, stride, st);
    else if (lg_tc_cin3_supported(N, Hb, Wb, A, B, stride))
      e = lg_tc_cin3_wgrad(big, small, dW, N, Hb, Wb, B, stride, st);
    else
      e = lg_tc_wgrad(big, small, dW, N, Hb, Wb, A, B, stride, st);
    if (e) return e;
  } else {
    lg_simt_wgrad(big, small, dW, N, Hb, Wb, A, B, stride, dtype, st);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_wgrad_padded(const void* big, const void* small, float* dW, int N, int Hb, int Wb,
                                      int A_big, int A, int B, int stride, void* stream) {
  if (int e = check_geom(__func__, N, Hb, Wb, A, B, stride, LG_BF16)) return e;
  LG_REQUIRE(big && small && dW && A_big >= A, "bad arguments");
  int e = lg_tc_wgrad_padded(big, small, dW, N, Hb, Wb, A_big, A, B, stride, (cudaStream_t)stream);
  if (e) return e;
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_transpose_fprop(const void* x_small, const float* W, const void* wpack,
                                         const float* bias, void* y_big, double* stats, int N, int Hb, int Wb,
                                         int A_out, int B_in, int stride, int act, int dtype, int use_tc,
                                         void* stream) {
  return lg_conv2d_dgrad(x_small, W, wpack, bias, y_big, stats, N, Hb, Wb, A_out, B_in, stride, act, dtype, use_tc,
                         nullptr, stream);
}

extern "C" int lg_conv2d_transpose_dgrad(const void* dy_big, const float* W, const void* wpack, void* dx_small,
                                         int N, int Hb, int Wb, int A_out, int B_in, int stride, int dtype,
                                         int use_tc, void* stream) {
  return lg_conv2d_fprop(dy_big, W, wpack, nullptr, dx_small, nullptr, N, Hb, Wb, A_out, B_in, stride, dtype, use_tc,
                         nullptr, stream);
}

extern "C" int lg_gemm(const void* A, const float* Bm, const float* bias, void* C, int M, int N, int K,
                       int transA, int transB, int accumulate, int a_dtype, int c_dtype, void* stream) {
  LG_REQUIRE(A && Bm && C && M > 0 && N > 0 && K > 0, "bad arguments");
  LG_REQUIRE(!accumulate || c_dtype == LG_F32, "accumulate needs an fp32 C");
  LG_REQUIRE(!accumulate || !bias, "bias is only applied when accumulate == 0");
  lg_simt_dense(A, Bm, bias, C, M, N, K, transA, transB, accumulate, a_dtype, c_dtype, (cudaStream_t)stream);
  LG_LAUNCH_CHECK();
  return LG_OK;
}
