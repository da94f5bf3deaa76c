// Input augmentation of the train step (eager_trainer.py:127-131), the reference's chain of five tf.image / tf.random
// ops on `real_image_1` fused into two launches:
//
//   new = flip_left_right(x)            per image, with probability 1/2            (tf.image.random_flip_left_right)
//   new = new + db                      db ~ U[-0.02, 0.02), one draw per batch    (tf.image.random_brightness)
//   new = (new - mean_c) * fc + mean_c  fc ~ U[0.75, 1.003), mean per image and channel over H x W  (random_contrast)
//   new = hue_rotate(new, dh)           dh ~ U[-0.03, 0.03) of a turn              (tf.image.random_hue(x, 0.03, -0.03):
//                                       the third positional argument of the reference's call is the SEED)
//   new = new + 0.1 * N(0, 0.2)         per element                                (tf.random.normal)
//
// Pass 1 (lg_augment_prepare, one CTA per image): per-channel means of the raw image (the flip does not change them
// and the brightness shifts them by db) and, when asked, this step's draws from Philox4x32-10(seed, step) - the step
// counter lives on the device (pass 2 advances it), so the launches are CUDA-graph replayable.  Pass 2 (lg_augment_apply): one read of x, one
// write of the result in the activation dtype; the hue rotation works on (hue, min, max) of the pixel - the form
// of TF's fused AdjustHue kernel, which is defined for any value range (the images are in [-1, 1], not [0, 1]).
#include <curand_kernel.h>

#include "common.cuh"

namespace {

constexpr int AT = 256;

// params: [0] db, [1] fc, [2] dh, [3] step (bit pattern); per image i at 4 + 4 i: mean_r, mean_g, mean_b, flip
__global__ void __launch_bounds__(AT) augment_prepare_kernel(const float* __restrict__ x, int HW,
                                                             unsigned long long* state, float* params, float max_b,
                                                             float c_lo, float c_hi, float max_h, int draw) {
  __shared__ double sh[3][AT / 32];
  const int n = blockIdx.x;
  const float* img = x + (int64_t)n * HW * 3;
  double s[3] = {0.0, 0.0, 0.0};
  // 4 pixels (12 floats, three aligned 16-byte loads) per thread and iteration
  const float4* v = reinterpret_cast<const float4*>(img);
  for (int g = threadIdx.x; g < HW / 4; g += AT) {
    const float4 a = __ldg(v + 3 * g), b = __ldg(v + 3 * g + 1), c = __ldg(v + 3 * g + 2);
    s[0] += (double)a.x + a.w + b.z + c.y;
    s[1] += (double)a.y + b.x + b.w + c.z;
    s[2] += (double)a.z + b.y + c.x + c.w;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    s[c] = warp_sum(s[c]);
    if (lane == 0) sh[c][w] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < AT / 32; ++i) t += sh[threadIdx.x][i];
    params[4 + 4 * n + threadIdx.x] = (float)(t / (double)HW);
  }
  if (state != nullptr && threadIdx.x == 0) {
    const unsigned long long seed = state[0], step = state[1];
    if (n == 0) params[3] = __uint_as_float((unsigned)step);  // pass 2 derives its noise stream from it
    if (draw) {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, /*subsequence*/ (unsigned long long)n, /*offset*/ 4ull * step, &st);
      const float4 u = curand_uniform4(&st);                  // (0, 1]
      params[4 + 4 * n + 3] = u.x <= 0.5f ? 1.f : 0.f;
      if (n == 0) {
        params[0] = (2.f * (1.f - u.y) - 1.f) * max_b;        // [-max_b, max_b)
        params[1] = c_lo + (1.f - u.z) * (c_hi - c_lo);
        params[2] = (2.f * (1.f - u.w) - 1.f) * max_h;
      }
    }
  }
}

// Hue rotation by dh turns, on (hue in [0,6), min, max) of the pixel.
__device__ __forceinline__ void hue_rotate(float& r, float& g, float& b, float dh) {
  const float vmax = fmaxf(r, fmaxf(g, b)), vmin = fminf(r, fminf(g, b));
  const float range = vmax - vmin;
  if (!(range > 0.f)) return;
  float h;
  if (r == vmax) h = (g - b) / range;                // [-1, 1]
  else if (g == vmax) h = 2.f + (b - r) / range;     // [1, 3]
  else h = 4.f + (r - g) / range;                    // [3, 5]
  h += 6.f * dh;
  h -= 6.f * floorf(h * (1.f / 6.f));                // [0, 6)
  if (h >= 6.f) h = 0.f;
  const float f = h - 2.f * floorf(h * 0.5f);        // h mod 2
  const float xm = vmin + range * (1.f - fabsf(f - 1.f));
  const int sector = (int)h;
  switch (sector) {
    case 0: r = vmax; g = xm; b = vmin; break;
    case 1: r = xm; g = vmax; b = vmin; break;
    case 2: r = vmin; g = vmax; b = xm; break;
    case 3: r = vmin; g = xm; b = vmax; break;
    case 4: r = xm; g = vmin; b = vmax; break;
    default: r = vmax; g = vmin; b = xm; break;
  }
}

template <typename TO>
__global__ void __launch_bounds__(AT) augment_apply_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                                           const float* __restrict__ noise,
                                                           unsigned long long* state, float noise_std,
                                                           TO* __restrict__ out, int N, int H, int W) {
  const int groups_per_row = W / 4;
  const int64_t total = (int64_t)N * H * groups_per_row;
  const int64_t gid = (int64_t)blockIdx.x * AT + threadIdx.x;
  if (gid >= total) return;
  const int gx = (int)(gid % groups_per_row);
  const int64_t row = gid / groups_per_row;                  // n * H + y
  const int n = (int)(row / H);
  const float db = params[0], fc = params[1], dh = params[2];
  const float* pm = params + 4 + 4 * n;
  const float mean[3] = {pm[0], pm[1], pm[2]};
  const bool flip = pm[3] > 0.5f;
  // source pixel group: mirrored group, pixels in reverse order
  const int sgx = flip ? groups_per_row - 1 - gx : gx;
  const float4* src = reinterpret_cast<const float4*>(x + (row * W + 4 * sgx) * 3);
  const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
  float px[4][3] = {{a.x, a.y, a.z}, {a.w, b.x, b.y}, {b.z, b.w, c.x}, {c.y, c.z, c.w}};
  if (flip) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float t = px[0][ch]; px[0][ch] = px[3][ch]; px[3][ch] = t;
      t = px[1][ch]; px[1][ch] = px[2][ch]; px[2][ch] = t;
    }
  }
  float nz[12];
  if (noise != nullptr) {
    const float4* nsrc = reinterpret_cast<const float4*>(noise + (row * W + 4 * gx) * 3);
    const float4 n0 = __ldg(nsrc), n1 = __ldg(nsrc + 1), n2 = __ldg(nsrc + 2);
    nz[0] = n0.x; nz[1] = n0.y; nz[2] = n0.z; nz[3] = n0.w; nz[4] = n1.x; nz[5] = n1.y; nz[6] = n1.z; nz[7] = n1.w;
    nz[8] = n2.x; nz[9] = n2.y; nz[10] = n2.z; nz[11] = n2.w;
  } else {
    const unsigned step = __float_as_uint(params[3]);
    // the step counter advances here: every thread takes the step from params[3], which pass 1 wrote
    if (gid == 0) state[1] += 1ull;
    curandStatePhilox4_32_10_t st;
    // subsequences below 2^32 belong to the per-image draws of pass 1.  The offset counts 32-bit Philox outputs:
    // a thread consumes 3 x curand_normal4 = 12 of them per step, so step k owns outputs [12k, 12k + 12)
    curand_init(state[0], (1ull << 32) + (unsigned long long)gid, 12ull * step, &st);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float4 g4 = curand_normal4(&st);
      nz[4 * i] = g4.x * noise_std; nz[4 * i + 1] = g4.y * noise_std;
      nz[4 * i + 2] = g4.z * noise_std; nz[4 * i + 3] = g4.w * noise_std;
    }
  }
  float o[12];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float r = (px[p][0] - mean[0]) * fc + mean[0] + db;
    float g = (px[p][1] - mean[1]) * fc + mean[1] + db;
    float bl = (px[p][2] - mean[2]) * fc + mean[2] + db;
    hue_rotate(r, g, bl, dh);
    o[3 * p] = r + nz[3 * p]; o[3 * p + 1] = g + nz[3 * p + 1]; o[3 * p + 2] = bl + nz[3 * p + 2];
  }
  TO* dst = out + (row * W + 4 * gx) * 3;
  if constexpr (sizeof(TO) == 4) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = make_float4(o[0], o[1], o[2], o[3]);
    d4[1] = make_float4(o[4], o[5], o[6], o[7]);
    d4[2] = make_float4(o[8], o[9], o[10], o[11]);
  } else {
    uint2* d2 = reinterpret_cast<uint2*>(dst);              // 12 bf16 = 24 bytes, 8-byte aligned
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[4 * i], o[4 * i + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(o[4 * i + 2], o[4 * i + 3]);
      d2[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
  }
}

// out[i] ~ N(0, 1): Philox4x32-10(seed, subsequence = thread, offset = 4 * calls-per-thread * step), four values per
// thread and pass; the step counter lives on the device and advances inside the launch (CUDA-graph replayable).
__global__ void __launch_bounds__(256) normal_fill_kernel(float* __restrict__ out, int64_t n, unsigned long long* state) {
  const unsigned long long seed = state[0], step = state[1];
  const int64_t nthreads = (int64_t)gridDim.x * 256;
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t per = (n + 4 * nthreads - 1) / (4 * nthreads);          // float4 draws per thread and launch
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (2ull << 32) + (unsigned long long)gid, 4ull * (unsigned long long)per * step, &st);
  for (int64_t q = 0; q < per; ++q) {
    const int64_t e = (q * nthreads + gid) * 4;
    const float4 g = curand_normal4(&st);
    if (e + 3 < n) *reinterpret_cast<float4*>(out + e) = g;
    else {
      if (e < n) out[e] = g.x;
      if (e + 1 < n) out[e + 1] = g.y;
      if (e + 2 < n) out[e + 2] = g.z;
    }
  }
  // every thread has read the step by the time the LAST CTA increments it: a separate tiny launch would be
  // simpler, a grid-wide ticket keeps it one launch
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(&state[2], 1ull);
    last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) { state[2] = 0ull; state[1] = step + 1ull; }
}

}  // namespace

// Standard-normal fill of a fp32 buffer (the generator noise of eager_trainer.py:125: tf.random.normal).
// state: device uint64[3] = {seed, step, 0}; the launch advances `step`.
extern "C" int lg_normal_fill(float* out, int64_t n, void* state, void* stream) {
  LG_REQUIRE(out && state && n > 0, "bad arguments");
  LG_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  int64_t need = (n / 4 + 255) / 256;
  int64_t cap = (int64_t)lg_num_sms() * 4;
  const unsigned grid = (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
  normal_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, (unsigned long long*)state);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_augment_prepare(const float* x, int N, int H, int W, void* state, float* params, float max_brightness,
                                  float contrast_lo, float contrast_hi, float max_hue, int draw, void* stream) {
  LG_REQUIRE(x && params && N > 0 && H > 0 && W > 0, "bad arguments");
  LG_REQUIRE(!draw || state, "drawing needs the {seed, step} state");
  if ((H * W) % 4 != 0) { lg_set_error("lg_augment_prepare: H*W must be a multiple of 4"); return LG_ERR_UNSUPPORTED; }
  augment_prepare_kernel<<<N, AT, 0, (cudaStream_t)stream>>>(x, H * W, (unsigned long long*)state, params,
                                                            max_brightness, contrast_lo, contrast_hi, max_hue, draw);
  LG_LAUNCH_CHECK();
  return LG_OK;
}

extern "C" int lg_augment_apply(const float* x, const float* params, const float* noise, void* state, float noise_std,
                                void* out, int N, int H, int W, int out_dtype, void* stream) {
  LG_REQUIRE(x && params && out && N > 0 && H > 0 && W > 0, "bad arguments");
  LG_REQUIRE(noise || state, "either a noise tensor or the {seed, step} state");
  LG_REQUIRE(out_dtype == LG_F32 || out_dtype == LG_BF16, "bad dtype");
  if (W % 4 != 0) { lg_set_error("lg_augment_apply: W must be a multiple of 4"); return LG_ERR_UNSUPPORTED; }
  const int64_t total = (int64_t)N * H * (W / 4);
  const unsigned grid = (unsigned)((total + AT - 1) / AT);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == LG_F32)
    augment_apply_kernel<float><<<grid, AT, 0, st>>>(x, params, noise, (unsigned long long*)state, noise_std,
                                                     (float*)out, N, H, W);
  else
    augment_apply_kernel<bf16><<<grid, AT, 0, st>>>(x, params, noise, (unsigned long long*)state, noise_std,
                                                    (bf16*)out, N, H, W);
  LG_LAUNCH_CHECK();
  return LG_OK;
}
