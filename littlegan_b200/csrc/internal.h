// Internal (non-exported) entry points shared between the translation units.
#pragma once
#include <cuda_runtime.h>

#include "littlegan_b200.h"

int lg_simt_fprop(const void* big, const float* W, const float* bias, void* out, double* stats, int N,
                  int Hb, int Wb, int A, int B, int s, int dtype, cudaStream_t st);
int lg_simt_dgrad(const void* small, const float* W, const float* bias, void* out, double* stats, int N,
                  int Hb, int Wb, int A, int B, int s, int act, int dtype, cudaStream_t st);
int lg_simt_wgrad(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A, int B,
                  int s, int dtype, cudaStream_t st);
int lg_simt_dense(const void* A, const float* Bm, const float* bias, void* C, int M, int N, int K, int tA,
                  int tB, int acc, int a_dtype, int c_dtype, cudaStream_t st);
int lg_simt_conv_bn(const void* x, const float* W, const float* scale, const float* shift, void* y, int N, int H,
                    int Wd, int Cin, int xs, int xo, int kh, int kw, int s, int ph, int pw, int Cout, int ys, int yo,
                    int relu, int dtype, cudaStream_t st);
// tcgen05 conv + BN + ReLU unit (tc_convbn.cu): bf16 maps, Cin % 8 == 0, Cout % 16 == 0
int lg_tc_convbn_supported(int Cin, int x_stride, int x_off, int kh, int kw, int Cout, int y_stride, int y_off);
int64_t lg_tc_convbn_pack_bytes(int Cin, int kh, int kw, int Cout);
int lg_tc_convbn_pack(const float* W, void* wpack, int Cin, int kh, int kw, int Cout, cudaStream_t st);
int lg_tc_convbn(const void* x, const void* wpack, const float* scale, const float* shift, void* y, int N, int H, int Wd,
                 int Cin, int xs, int xo, int kh, int kw, int s, int ph, int pw, int Cout, int ys, int yo, int relu,
                 cudaStream_t st);

// tcgen05 / TMA path (tc_conv.cu).  Return LG_ERR_UNSUPPORTED when the geometry is not covered.
// `nb` (may be NULL): fuse the InstanceNorm-backward reduction of the layer below into the epilogue.
int lg_tc_fprop(const void* big, const void* wpack, const float* bias, void* out, double* stats, int N,
                int Hb, int Wb, int A, int B, int s, const lg_norm_bwd_t* nb, cudaStream_t st);
int lg_tc_dgrad(const void* small, const void* wpack, const float* bias, void* out, double* stats, int N,
                int Hb, int Wb, int A, int B, int s, int act, const lg_norm_bwd_t* nb, cudaStream_t st);
int lg_tc_wgrad(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A, int B,
                int s, cudaStream_t st);
int lg_tc_wgrad_padded(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A, int A_real,
                       int B, int s, cudaStream_t st);
int lg_tc_deconv_small_supported(int N, int Hb, int Wb, int A, int B, int s);
int lg_tc_deconv_small(const void* small, const float* W, const float* bias, void* out, double* stats, int N,
                       int Hb, int Wb, int A, int B, int s, int act, cudaStream_t st);
int lg_tc_dgrad4_supported(int N, int Hb, int Wb, int A, int B, int s);
int lg_tc_dgrad4(const void* small, const void* wpack, const float* bias, void* out, double* stats, int N, int Hb,
                 int Wb, int A, int B, int act, const lg_norm_bwd_t* nb, cudaStream_t st);
int lg_tc_cin3_supported(int N, int Hb, int Wb, int A, int B, int s);
int lg_tc_cin3_fprop(const void* img, const float* W, const float* bias, void* out, double* stats, int N, int Hb,
                     int Wb, int B, int s, const lg_norm_bwd_t* nb, cudaStream_t st);
int lg_tc_cin3_wgrad(const void* img, const void* small, float* dW, int N, int Hb, int Wb, int B, int s,
                     cudaStream_t st);
int lg_tc_rowconv_supported(int N, int Hb, int Wb, int A, int Cpad, int B, int s);
int lg_tc_rowconv_pack(const float* W, void* wpack, int A, int Cpad, int B, int s, cudaStream_t st);
int lg_tc_rowconv_fprop(const void* big, const void* wpack, const float* bias, void* out, double* stats, int N, int Hb,
                        int Wb, int A, int Cpad, int B, int s, const lg_norm_bwd_t* nb, cudaStream_t st);
int lg_tc_rowdeconv_supported(int N, int Hb, int Wb, int A, int B, int s);
int lg_tc_rowdeconv(const void* small, const float* W, const float* bias, void* out3, void* out8, double* stats,
                    int N, int Hb, int Wb, int A, int B, int s, int act, cudaStream_t st);
int lg_tc_rowdgrad_supported(int N, int Hb, int Wb, int A, int B, int s);
int lg_tc_rowdgrad_pack(const float* W, void* wpack, int A, int B, cudaStream_t st);
int lg_tc_rowdgrad(const void* small, const void* wpack, const float* bias, void* out, double* stats, int N, int Hb,
                   int Wb, int A, int B, int s, cudaStream_t st);
int lg_tc_rowwgrad_supported(int N, int Hb, int Wb, int A, int B, int s);
int lg_tc_rowwgrad(const void* big, const void* small, float* dW, int N, int Hb, int Wb, int A, int B, int s,
                   cudaStream_t st);
int lg_tc_supported(int op, int Hb, int Wb, int A, int B, int s, int N);
