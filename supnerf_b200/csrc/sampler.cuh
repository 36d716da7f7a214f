// Device helpers of the ray / box-sampler kernels (sampler.cu), shared with the batched render (render_batch.cu).
#pragma once
#include "common.cuh"
#include <math.h>

namespace snb {

// torch.minimum/maximum propagate NaN (utils.py:308-312).
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }

struct Slab {
  float t_near, t_far;
  bool hit;
  float tmin[3], tmax[3], inv[3];
};

// utils.py:303-319 in the reference's op order: reciprocal; (aabb - o) * inv; min/max; two compares.
__device__ __forceinline__ Slab slab_test2(const float o[3], const float d[3], const float lo[3], const float hi[3]) {
  Slab r;
  float t0[3], t1[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    r.inv[a] = __frcp_rn(d[a]);
    r.tmin[a] = __fmul_rn(__fsub_rn(lo[a], o[a]), r.inv[a]);
    r.tmax[a] = __fmul_rn(__fsub_rn(hi[a], o[a]), r.inv[a]);
    t0[a] = nan_min(r.tmin[a], r.tmax[a]);
    t1[a] = nan_max(r.tmin[a], r.tmax[a]);
  }
  r.t_near = nan_max(nan_max(t0[0], t0[1]), t0[2]);
  r.t_far = nan_min(nan_min(t1[0], t1[1]), t1[2]);
  bool inside = r.t_far > r.t_near;
  float m = inside ? 1.f : 0.f;
  r.hit = inside && (__fmul_rn(r.t_far, m) > 0.f);
  return r;
}

__device__ __forceinline__ Slab slab_test(const float o[3], const float d[3], const float half[3]) {
  const float lo[3] = {-half[0], -half[1], -half[2]};
  return slab_test2(o, d, lo, half);
}

// torch.maximum / torch.minimum backward: the selected operand gets the gradient, ties split it.
__device__ __forceinline__ void pick_max(float a, float b, float g, float& ga, float& gb) {
  if (a == b) { ga = 0.5f * g; gb = 0.5f * g; }
  else if (a > b) { ga = g; gb = 0.f; }
  else { ga = 0.f; gb = g; }
}
__device__ __forceinline__ void pick_min(float a, float b, float g, float& ga, float& gb) {
  if (a == b) { ga = 0.5f * g; gb = 0.5f * g; }
  else if (a < b) { ga = g; gb = 0.f; }
  else { ga = 0.f; gb = g; }
}

// gradients of (t_near, t_far) w.r.t. origin, direction and the two box corners, per axis
__device__ __forceinline__ void slab_backward(const Slab& sl, const float o[3], const float lo[3], const float hi[3],
                                              float gnear, float gfar, float go[3], float gd[3], float glo[3], float ghi[3]) {
  float t0[3], t1[3], g0[3], g1[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { t0[a] = nan_min(sl.tmin[a], sl.tmax[a]); t1[a] = nan_max(sl.tmin[a], sl.tmax[a]); }
  float gxy, gz_;
  pick_max(nan_max(t0[0], t0[1]), t0[2], gnear, gxy, gz_);
  g0[2] = gz_;
  pick_max(t0[0], t0[1], gxy, g0[0], g0[1]);
  pick_min(nan_min(t1[0], t1[1]), t1[2], gfar, gxy, gz_);
  g1[2] = gz_;
  pick_min(t1[0], t1[1], gxy, g1[0], g1[1]);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float gmin_a, gmax_a, ga, gb;
    pick_min(sl.tmin[a], sl.tmax[a], g0[a], gmin_a, gmax_a);  // t0 = min(tmin, tmax)
    pick_max(sl.tmin[a], sl.tmax[a], g1[a], ga, gb);          // t1 = max(tmin, tmax)
    gmin_a += ga; gmax_a += gb;
    // tmin = (lo - o) * inv ; tmax = (hi - o) * inv ; inv = 1/d
    go[a] = -(gmin_a + gmax_a) * sl.inv[a];
    glo[a] = gmin_a * sl.inv[a];
    ghi[a] = gmax_a * sl.inv[a];
    const float ginv = gmin_a * (lo[a] - o[a]) + gmax_a * (hi[a] - o[a]);
    gd[a] = -ginv * sl.inv[a] * sl.inv[a];
  }
}

}  // namespace snb
