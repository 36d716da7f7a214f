// Device helpers of the compositing kernels (composite.cu), shared with the batched render (render_batch.cu).
#pragma once
#include "common.cuh"

namespace snb {

constexpr int kMaxChunks = 64;  // S <= 2048

struct SampleTerms {
  float alpha, t, e, delta, sr;
};

// The reference's fp32 rounding sequence, literally: a = 1 - exp(-relu(s)*d); t = (1 - a) + 1e-10.
__device__ __forceinline__ SampleTerms sample_terms(float s, float zk, float znext, bool last, bool relu) {
  SampleTerms r;
  r.delta = last ? 1e10f : (znext - zk);
  r.sr = relu ? fmaxf(s, 0.f) : s;
  r.e = expf(-r.sr * r.delta);
  r.alpha = 1.f - r.e;
  r.t = (1.f - r.alpha) + 1e-10f;
  return r;
}

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= u;
  }
  return v;
}

__device__ __forceinline__ float warp_rev_incl_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += u;
  }
  return v;
}

template <int LPR>
__device__ __forceinline__ float seg_excl_prod(float v, int sl, float* total) {
  float incl = v;
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) {
    const float u = __shfl_up_sync(0xffffffffu, incl, o, LPR);
    if (sl >= o) incl *= u;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1, LPR);
  if (sl == 0) excl = 1.f;
  *total = __shfl_sync(0xffffffffu, incl, LPR - 1, LPR);
  return excl;
}
template <int LPR>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LPR);
  return v;
}
// sum over the strictly later lanes of the segment
template <int LPR>
__device__ __forceinline__ float seg_rev_excl_sum(float v, int sl) {
  float incl = v;
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) {
    const float u = __shfl_down_sync(0xffffffffu, incl, o, LPR);
    if (sl + o < LPR) incl += u;
  }
  return incl - v;
}

struct Lane4 {
  float s[4], z[5], c[4][3];
  bool valid;   // the lane's 4 samples exist (k0 < S)
};

template <int LPR>
__device__ __forceinline__ Lane4 load_lane4(const float* __restrict__ sg, const float* __restrict__ cg, const float* __restrict__ zr,
                                            int k0, int S) {
  Lane4 L;
  L.valid = k0 < S;
  if (L.valid) {
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(sg + k0));
    const float4 z4 = zr ? __ldg(reinterpret_cast<const float4*>(zr + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);   // zr == NULL: the caller fills z
    const float4 a = __ldg(reinterpret_cast<const float4*>(cg + 3 * k0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(cg + 3 * k0 + 4));
    const float4 c = __ldg(reinterpret_cast<const float4*>(cg + 3 * k0 + 8));
    L.s[0] = s4.x; L.s[1] = s4.y; L.s[2] = s4.z; L.s[3] = s4.w;
    L.z[0] = z4.x; L.z[1] = z4.y; L.z[2] = z4.z; L.z[3] = z4.w;
    L.z[4] = (zr && k0 + 4 < S) ? __ldg(zr + k0 + 4) : 0.f;
    L.c[0][0] = a.x; L.c[0][1] = a.y; L.c[0][2] = a.z; L.c[1][0] = a.w;
    L.c[1][1] = b.x; L.c[1][2] = b.y; L.c[2][0] = b.z; L.c[2][1] = b.w;
    L.c[3][0] = c.y; L.c[3][1] = c.z; L.c[3][2] = c.w; L.c[2][2] = c.x;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) { L.s[j] = 0.f; L.z[j] = 0.f; L.c[j][0] = L.c[j][1] = L.c[j][2] = 0.f; }
    L.z[4] = 0.f;
  }
  return L;
}

}  // namespace snb
