// InstanceNormalization backward, pass 1, fused into the epilogue of the convolution kernel that PRODUCES
// the gradient (instance.py:105-128; backward derived in SURVEY 8 a5).
//
// A tensor-core conv kernel of the backward pass writes g = dL/d(a), where a = LeakyReLU(IN(z)) is the
// output of the previous layer.  The norm backward of that layer needs the per-sample sums
//     red[n] = ( sum dy , sum dy * xhat ),   dy = g * LeakyReLU'(y),  y = gamma*xhat + beta,  xhat = (z - mu)/s
// before it can form dz.  Instead of a separate pass that re-reads g and z, the epilogue reads the matching
// z row (the only extra traffic), accumulates the two sums from the fp32 accumulators and stores dy, so the
// follow-up apply kernel is a single read-read-write pass (lg_instnorm_act_bwd_apply with dy_ready).
#pragma once
#include "common.cuh"

struct NormBwdDev {
  const bf16* z;          // pre-norm conv output of the layer below, same shape as this kernel's output
  const double* stats;    // [N][2] (sum z, sum z^2)
  const float* gamma;
  const float* beta;
  double* red;            // [N][2] += (sum dy, sum dy*xhat)
  float eps, alpha;
  double inv_M;           // 1 / (elements per sample)
};

static inline NormBwdDev lg_make_norm_bwd(const lg_norm_bwd_t* h, int64_t elems_per_sample) {
  NormBwdDev d;
  d.z = (const bf16*)h->z; d.stats = h->stats; d.gamma = h->gamma; d.beta = h->beta; d.red = h->red;
  d.eps = h->eps; d.alpha = h->alpha; d.inv_M = 1.0 / (double)elems_per_sample;
  return d;
}

struct NormBwdCoef { float k1, k0, y1, y0; };   // xhat = fma(z, k1, k0) ; y = fma(z, y1, y0)

__device__ __forceinline__ NormBwdCoef nb_coef(const NormBwdDev& nb, int n) {
  const double mu = nb.stats[2 * n] * nb.inv_M;
  const double var = nb.stats[2 * n + 1] * nb.inv_M - mu * mu;
  const double sigma = var > 0.0 ? sqrt(var) : 0.0;
  const float inv = (float)(1.0 / (sigma + (double)nb.eps));
  const float ga = nb.gamma[0], be = nb.beta[0];
  NormBwdCoef c;
  c.k1 = inv; c.k0 = -(float)mu * inv; c.y1 = ga * c.k1; c.y0 = fmaf(ga, c.k0, be);
  return c;
}

// Up to NCH 16-channel chunks (NCH x 32 B) of one z row, kept in registers.
template <int NCH>
struct NormBwdZn { uint4 v[2 * NCH]; };
typedef NormBwdZn<4> NormBwdZ;

template <int NCH>
__device__ __forceinline__ void nb_load(NormBwdZn<NCH>& b, const bf16* zrow, int nchunks, bool valid) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (valid && c < nchunks) {
      b.v[2 * c] = __ldg(reinterpret_cast<const uint4*>(zrow + 16 * c));
      b.v[2 * c + 1] = __ldg(reinterpret_cast<const uint4*>(zrow + 16 * c + 8));
    } else {
      b.v[2 * c] = make_uint4(0, 0, 0, 0);
      b.v[2 * c + 1] = make_uint4(0, 0, 0, 0);
    }
  }
}

__device__ __forceinline__ void nb_prefetch_l2(const bf16* p, int bytes, bool valid) {
  if (!valid) return;
  for (int o = 0; o < bytes; o += 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(p) + o));
}

// One 16-channel chunk: a[] = fp32 gradient w.r.t. the activation; returns dy packed as 8 x bf16x2.
__device__ __forceinline__ void nb_chunk(const float (&a)[16], const uint4& z0, const uint4& z1, const NormBwdCoef& c,
                                         float alpha, float& r1, float& r2, uint32_t (&pk)[8]) {
  const uint32_t zw[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float2 zf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zw[e]));
    const float d0 = a[2 * e] * (fmaf(zf.x, c.y1, c.y0) > 0.f ? 1.f : alpha);
    const float d1 = a[2 * e + 1] * (fmaf(zf.y, c.y1, c.y0) > 0.f ? 1.f : alpha);
    r1 += d0 + d1;
    r2 = fmaf(d0, fmaf(zf.x, c.k1, c.k0), r2);
    r2 = fmaf(d1, fmaf(zf.y, c.k1, c.k0), r2);
    __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
    pk[e] = *reinterpret_cast<uint32_t*>(&h);
  }
}
