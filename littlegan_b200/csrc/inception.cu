// Inception-2015 pool_3 forward for the FID statistics pass (reference fid.py:36-106): the pieces around the
// convolution GEMMs - TF-1.x bilinear resize + input normalisation, 3x3 max / average pools that write straight into
// a channel slice of the block's concat output, the final 8x8 mean - and the C-ABI entry points of the family.
// All of them are HBM-bound gathers: one thread per output element with the channel index fastest, so a warp reads
// and writes contiguous channel runs.
#include "common.cuh"
#include "internal.h"

namespace {

// TF-1.x ResizeBilinear (align_corners = false, no half-pixel centres): src = dst * in / out, then (v - sub) * mul.
// The output pixel has Cp >= C channels, the extra ones zero (RGB -> 8 channels = one 16-byte run per pixel, which
// puts the first conv unit on the tensor-core path).
template <typename S, typename T>
__global__ void resize_bilinear_kernel(const S* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Cp,
                                       int Ho, int Wo, float sy, float sx, float sub, float mul) {
  const int64_t total = (int64_t)N * Ho * Wo * Cp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % Cp);
    if (c >= C) { y[e] = from_f<T>(0.f); continue; }
    int64_t t = e / Cp;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const float fy = oy * sy, fx = ox * sx;
    const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float wy = fy - y0, wx = fx - x0;
    const S* base = x + (int64_t)n * H * W * C + c;
    const float v00 = (float)base[((int64_t)y0 * W + x0) * C], v01 = (float)base[((int64_t)y0 * W + x1) * C];
    const float v10 = (float)base[((int64_t)y1 * W + x0) * C], v11 = (float)base[((int64_t)y1 * W + x1) * C];
    const float top = v00 + (v01 - v00) * wx, bot = v10 + (v11 - v10) * wx;
    y[e] = from_f<T>((top + (bot - top) * wy - sub) * mul);
  }
}

// RGB -> 8-channel bf16 pixels: one thread per output pixel, one 16-byte store
template <typename S>
__global__ void resize_rgb8_kernel(const S* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int Ho, int Wo,
                                   float sy, float sx, float sub, float mul) {
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int ox = (int)(e % Wo);
    int64_t t = e / Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const float fy = oy * sy, fx = ox * sx;
    const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float wy = fy - y0, wx = fx - x0;
    const S* base = x + (int64_t)n * H * W * 3;
    const S* p00 = base + ((int64_t)y0 * W + x0) * 3;
    const S* p01 = base + ((int64_t)y0 * W + x1) * 3;
    const S* p10 = base + ((int64_t)y1 * W + x0) * 3;
    const S* p11 = base + ((int64_t)y1 * W + x1) * 3;
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v00 = (float)p00[c], v01 = (float)p01[c], v10 = (float)p10[c], v11 = (float)p11[c];
      const float top = v00 + (v01 - v00) * wx, bot = v10 + (v11 - v10) * wx;
      o[c] = (top + (bot - top) * wy - sub) * mul;
    }
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
    h[0] = __floats2bfloat162_rn(o[0], o[1]);
    h[1] = __floats2bfloat162_rn(o[2], 0.f);
    h[2] = __floats2bfloat162_rn(0.f, 0.f);
    h[3] = h[2];
    reinterpret_cast<uint4*>(y)[e] = pk;
  }
}

// mode 0: max; 1: average over the in-bounds taps only (the 2015 graph's AvgPool with SAME padding);
// 2: average with the padding counted (divisor k*k)
template <typename T>
__global__ void pool2d_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xs, int k,
                              int s, int pad, int Ho, int Wo, int mode, int ys) {
  const int64_t total = (int64_t)N * Ho * Wo * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    int64_t t = e / C;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float acc = mode == 0 ? -INFINITY : 0.f;
    int cnt = 0;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * s - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * s - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const float v = to_f(x[(((int64_t)n * H + iy) * W + ix) * xs + c]);
        acc = mode == 0 ? fmaxf(acc, v) : acc + v;
        ++cnt;
      }
    }
    if (mode == 1) acc /= (float)cnt;
    else if (mode == 2) acc /= (float)(k * k);
    y[(((int64_t)n * Ho + oy) * Wo + ox) * ys + c] = from_f<T>(acc);
  }
}

// bf16 maps whose channel counts / slice offsets are multiples of 8: one thread per (output pixel, 8 channels),
// 16-byte loads and stores (the scalar kernel above moves 2 bytes per instruction)
__global__ void pool2d_vec8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C8, int xs,
                                   int k, int s, int pad, int Ho, int Wo, int mode, int ys) {
  // grid: x = blocks over (ox, c8) of one output row, y = oy, z = n  (no 64-bit divisions per element)
  const int row_elems = Wo * C8;
  const int oy = blockIdx.y, n = blockIdx.z;
  const bf16* xin = x + (int64_t)n * H * W * xs;
  bf16* yout = y + ((int64_t)n * Ho + oy) * Wo * ys;
  const int iy0 = oy * s - pad;
  const int ky_lo = max(0, -iy0), ky_hi = min(k, H - iy0);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < row_elems; e += gridDim.x * blockDim.x) {
    const int ox = e / C8, c = (e - ox * C8) * 8;
    const int ix0 = ox * s - pad;
    const int kx_lo = max(0, -ix0), kx_hi = min(k, W - ix0);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = mode == 0 ? -INFINITY : 0.f;
    for (int ky = ky_lo; ky < ky_hi; ++ky) {
      const bf16* row = xin + ((int64_t)(iy0 + ky) * W + ix0) * xs + c;
      for (int kx = kx_lo; kx < kx_hi; ++kx) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + kx * xs));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          if (mode == 0) { acc[2 * i] = fmaxf(acc[2 * i], f.x); acc[2 * i + 1] = fmaxf(acc[2 * i + 1], f.y); }
          else { acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        }
      }
    }
    const float div = mode == 1 ? (float)((ky_hi - ky_lo) * (kx_hi - kx_lo)) : (float)(k * k);
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      oh[i] = mode == 0 ? __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1])
                        : __floats2bfloat162_rn(acc[2 * i] / div, acc[2 * i + 1] / div);
    *reinterpret_cast<uint4*>(yout + ox * ys + c) = o;
  }
}

// 3x3 pools as a separable sliding window: a thread owns (n, ox, 8 channels) and walks a segment of output rows,
// keeping the horizontal max / sum of the three input rows of the window in registers - 3 (stride 1) or 6 (stride 2)
// 16-byte loads per output instead of 9.
template <int S>
__global__ void pool3_slide_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int H, int W, int C8, int xs, int pad,
                                   int Ho, int Wo, int mode, int ys, int seg) {
  const int row_elems = Wo * C8;
  const int n = blockIdx.z;
  const int oy_begin = blockIdx.y * seg, oy_end = min(Ho, oy_begin + seg);
  const bf16* xin = x + (int64_t)n * H * W * xs;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < row_elems; e += gridDim.x * blockDim.x) {
    const int ox = e / C8, c = (e - ox * C8) * 8;
    const int ix0 = ox * S - pad;
    const bool cv0 = ix0 >= 0 && ix0 < W, cv1 = ix0 + 1 >= 0 && ix0 + 1 < W, cv2 = ix0 + 2 >= 0 && ix0 + 2 < W;
    const int ncols = (int)cv0 + (int)cv1 + (int)cv2;
    const float fill = mode == 0 ? -INFINITY : 0.f;
    float h[3][8];
    bool hv[3];
    auto load_row = [&](int r, float (&o)[8]) -> bool {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fill;
      if (r < 0 || r >= H) return false;
      const bf16* row = xin + ((int64_t)r * W + ix0) * xs + c;
      uint4 v[3];
      v[0] = cv0 ? __ldg(reinterpret_cast<const uint4*>(row)) : make_uint4(0, 0, 0, 0);
      v[1] = cv1 ? __ldg(reinterpret_cast<const uint4*>(row + xs)) : make_uint4(0, 0, 0, 0);
      v[2] = cv2 ? __ldg(reinterpret_cast<const uint4*>(row + 2 * xs)) : make_uint4(0, 0, 0, 0);
      const bool cv[3] = {cv0, cv1, cv2};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!cv[j]) continue;
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&v[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(hh[i]);
          if (mode == 0) { o[2 * i] = fmaxf(o[2 * i], f.x); o[2 * i + 1] = fmaxf(o[2 * i + 1], f.y); }
          else { o[2 * i] += f.x; o[2 * i + 1] += f.y; }
        }
      }
      return true;
    };
    int iy0 = oy_begin * S - pad;
    hv[0] = load_row(iy0, h[0]);
    hv[1] = load_row(iy0 + 1, h[1]);
    hv[2] = load_row(iy0 + 2, h[2]);
    for (int oy = oy_begin; oy < oy_end; ++oy) {
      const int nrows = (int)hv[0] + (int)hv[1] + (int)hv[2];
      const float div = mode == 1 ? (float)(nrows * ncols) : 9.f;
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a, b;
        if (mode == 0) {
          a = fmaxf(fmaxf(h[0][2 * i], h[1][2 * i]), h[2][2 * i]);
          b = fmaxf(fmaxf(h[0][2 * i + 1], h[1][2 * i + 1]), h[2][2 * i + 1]);
        } else {
          a = (h[0][2 * i] + h[1][2 * i] + h[2][2 * i]) / div;
          b = (h[0][2 * i + 1] + h[1][2 * i + 1] + h[2][2 * i + 1]) / div;
        }
        oh[i] = __floats2bfloat162_rn(a, b);
      }
      *reinterpret_cast<uint4*>(y + (((int64_t)n * Ho + oy) * Wo + ox) * ys + c) = o;
      if (oy + 1 < oy_end) {                                   // slide the window down by S rows
        iy0 += S;
        if (S == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { h[0][i] = h[1][i]; h[1][i] = h[2][i]; }
          hv[0] = hv[1]; hv[1] = hv[2];
          hv[2] = load_row(iy0 + 2, h[2]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) h[0][i] = h[2][i];
          hv[0] = hv[2];
          hv[1] = load_row(iy0 + 1, h[1]);
          hv[2] = load_row(iy0 + 2, h[2]);
        }
      }
    }
  }
}

// pool_3: mean over the HW positions of the last map, fp32 out
template <typename T>
__global__ void global_avgpool_kernel(const T* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
  const int64_t total = (int64_t)N * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int n = (int)(e / C);
    const T* p = x + (int64_t)n * HW * C + c;
    float acc = 0.f;
    for (int i = 0; i < HW; ++i) acc += to_f(p[(int64_t)i * C]);
    y[e] = acc / (float)HW;
  }
}

inline int grid_for(int64_t total) {
  int64_t need = (total + 255) / 256, cap = (int64_t)lg_num_sms() * 16;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace

extern "C" int lg_pack_conv_bn_weights(const float* W, void* wpack, int Cin, int kh, int kw, int Cout, void* stream) {
  LG_REQUIRE(Cin > 0 && kh > 0 && kw > 0 && Cout > 0, "invalid geometry");
  if (W == nullptr || wpack == nullptr) return (int)lg_tc_convbn_pack_bytes(Cin, kh, kw, Cout);
  int r = lg_tc_convbn_pack(W, wpack, Cin, kh, kw, Cout, (cudaStream_t)stream);
  if (r != LG_OK) { lg_set_error("%s: geometry has no tensor-core form", __func__); return r; }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_conv2d_bn_relu(const void* x, const float* W, const void* wpack, const float* scale, const float* shift,
                                 void* y, int N, int H, int Wd, int Cin, int x_stride, int x_off, int kh, int kw,
                                 int stride, int ph, int pw, int Cout, int y_stride, int y_off, int relu, int dtype,
                                 void* stream) {
  LG_REQUIRE(N > 0 && H > 0 && Wd > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && ph >= 0 && pw >= 0,
             "invalid geometry");
  LG_REQUIRE(H + 2 * ph >= kh && Wd + 2 * pw >= kw, "kernel larger than the padded input");
  LG_REQUIRE(x_off >= 0 && x_off + Cin <= x_stride && y_off >= 0 && y_off + Cout <= y_stride, "channel slice out of range");
  LG_REQUIRE(dtype == LG_F32 || dtype == LG_BF16, "dtype must be LG_F32 or LG_BF16");
  LG_REQUIRE(x && (W || wpack) && scale && shift && y, "null pointer");
  const int64_t M = (int64_t)N * ((H + 2 * ph - kh) / stride + 1) * ((Wd + 2 * pw - kw) / stride + 1);
  LG_REQUIRE(M < (1ll << 31) - 128 && (int64_t)kh * kw * Cin < (1ll << 31), "problem too large");
  if (dtype == LG_BF16 && wpack != nullptr && (int64_t)N * H * Wd * x_stride < (1ll << 31) &&   // 32-bit element offsets
      lg_tc_convbn_supported(Cin, x_stride, x_off, kh, kw, Cout, y_stride, y_off) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(wpack)) & 15) == 0) {
    lg_tc_convbn(x, wpack, scale, shift, y, N, H, Wd, Cin, x_stride, x_off, kh, kw, stride, ph, pw, Cout, y_stride, y_off,
                 relu, (cudaStream_t)stream);
  } else {
    LG_REQUIRE(W != nullptr, "this geometry / dtype runs the fp32-accumulate SIMT path and needs the fp32 kernel W");
    lg_simt_conv_bn(x, W, scale, shift, y, N, H, Wd, Cin, x_stride, x_off, kh, kw, stride, ph, pw, Cout, y_stride, y_off,
                    relu, dtype, (cudaStream_t)stream);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_pool2d(const void* x, void* y, int N, int H, int W, int C, int x_stride, int x_off, int k, int stride,
                         int pad, int mode, int y_stride, int y_off, int dtype, void* stream) {
  LG_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && pad < k, "invalid geometry");
  LG_REQUIRE(H + 2 * pad >= k && W + 2 * pad >= k, "window larger than the padded input");
  LG_REQUIRE(mode >= 0 && mode <= 2, "mode: 0 max, 1 average over valid taps, 2 average incl. padding");
  LG_REQUIRE(x_off >= 0 && x_off + C <= x_stride && y_off >= 0 && y_off + C <= y_stride, "channel slice out of range");
  LG_REQUIRE(dtype == LG_F32 || dtype == LG_BF16, "dtype must be LG_F32 or LG_BF16");
  LG_REQUIRE(x && y, "null pointer");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int g = grid_for((int64_t)N * Ho * Wo * C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LG_BF16 && ((C | x_stride | x_off | y_stride | y_off) & 7) == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
  {
    LG_REQUIRE(Ho <= 65535 && N <= 65535, "map too tall / batch too large for the pooling grid");
    const int row_elems = Wo * (C / 8);
    if (k == 3 && (stride == 1 || stride == 2)) {
      // row segments: enough threads to fill the GPU, as long as possible otherwise (less halo re-reading)
      int64_t per_seg = (int64_t)N * row_elems;
      int nseg = (int)((300000 + per_seg - 1) / per_seg);
      nseg = nseg < 1 ? 1 : (nseg > Ho ? Ho : nseg);
      const int seg = (Ho + nseg - 1) / nseg;
      const int threads = row_elems >= 128 ? 128 : ((row_elems + 31) / 32) * 32;
      dim3 grid((row_elems + threads - 1) / threads, (Ho + seg - 1) / seg, N);
      if (stride == 1)
        pool3_slide_kernel<1><<<grid, threads, 0, st>>>((const bf16*)x + x_off, (bf16*)y + y_off, H, W, C / 8, x_stride,
                                                        pad, Ho, Wo, mode, y_stride, seg);
      else
        pool3_slide_kernel<2><<<grid, threads, 0, st>>>((const bf16*)x + x_off, (bf16*)y + y_off, H, W, C / 8, x_stride,
                                                        pad, Ho, Wo, mode, y_stride, seg);
      LG_LAUNCH_CHECK();
      return LG_OK;
    }
    const int threads = row_elems >= 256 ? 256 : ((row_elems + 31) / 32) * 32;
    dim3 grid((row_elems + threads - 1) / threads, Ho, N);
    pool2d_vec8_kernel<<<grid, threads, 0, st>>>((const bf16*)x + x_off, (bf16*)y + y_off, N, H, W, C / 8, x_stride, k,
                                                 stride, pad, Ho, Wo, mode, y_stride);
  }
  else if (dtype == LG_BF16)
    pool2d_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x + x_off, (bf16*)y + y_off, N, H, W, C, x_stride, k, stride, pad,
                                           Ho, Wo, mode, y_stride);
  else
    pool2d_kernel<float><<<g, 256, 0, st>>>((const float*)x + x_off, (float*)y + y_off, N, H, W, C, x_stride, k, stride,
                                            pad, Ho, Wo, mode, y_stride);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_global_avgpool(const void* x, float* y, int N, int HW, int C, int dtype, void* stream) {
  LG_REQUIRE(x && y && N > 0 && HW > 0 && C > 0, "bad arguments");
  LG_REQUIRE(dtype == LG_F32 || dtype == LG_BF16, "dtype must be LG_F32 or LG_BF16");
  const int g = grid_for((int64_t)N * C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LG_BF16) global_avgpool_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, y, N, HW, C);
  else global_avgpool_kernel<float><<<g, 256, 0, st>>>((const float*)x, y, N, HW, C);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_resize_bilinear_norm(const void* x, void* y, int N, int H, int W, int C, int Cpad, int Ho, int Wo,
                                       float sub, float mul, int src_is_u8, int dtype, void* stream) {
  LG_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && Cpad >= C && Ho > 0 && Wo > 0, "bad arguments");
  LG_REQUIRE(dtype == LG_F32 || dtype == LG_BF16, "dtype must be LG_F32 or LG_BF16");
  const int g = grid_for((int64_t)N * Ho * Wo * Cpad);
  cudaStream_t st = (cudaStream_t)stream;
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  if (dtype == LG_BF16 && C == 3 && Cpad == 8 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const int g8 = grid_for((int64_t)N * Ho * Wo);
    if (src_is_u8) resize_rgb8_kernel<uint8_t><<<g8, 256, 0, st>>>((const uint8_t*)x, (bf16*)y, N, H, W, Ho, Wo, sy, sx, sub, mul);
    else resize_rgb8_kernel<float><<<g8, 256, 0, st>>>((const float*)x, (bf16*)y, N, H, W, Ho, Wo, sy, sx, sub, mul);
  } else if (src_is_u8) {
    if (dtype == LG_BF16) resize_bilinear_kernel<uint8_t, bf16><<<g, 256, 0, st>>>((const uint8_t*)x, (bf16*)y, N, H, W, C, Cpad, Ho, Wo, sy, sx, sub, mul);
    else resize_bilinear_kernel<uint8_t, float><<<g, 256, 0, st>>>((const uint8_t*)x, (float*)y, N, H, W, C, Cpad, Ho, Wo, sy, sx, sub, mul);
  } else {
    if (dtype == LG_BF16) resize_bilinear_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)x, (bf16*)y, N, H, W, C, Cpad, Ho, Wo, sy, sx, sub, mul);
    else resize_bilinear_kernel<float, float><<<g, 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C, Cpad, Ho, Wo, sy, sx, sub, mul);
  }
  LG_LAUNCH_CHECK();
  return LG_OK;
}
