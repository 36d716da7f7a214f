// K3 / K3b — alpha compositing along rays as a warp-level segmented scan, forward and reverse-scan
// backward.  Replaces renderer.py:43-65, :355-379 and utils.py:187-233 of the reference.
//
// One warp owns one ray at a time; lane l owns sample (chunk*32 + l).  Transmittance is an exclusive
// product scan across lanes (5 shuffles) carried across 32-sample chunks; the backward recomputes
// the forward quantities and runs the suffix sum as a reverse scan (no total-minus-prefix
// cancellation).  HBM-bound: 20*S+20 B/ray forward, 40*S+20 read + 20*S written backward.
#include "common.cuh"
#include "../../include/supnerf_b200.h"

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

__global__ void __launch_bounds__(256) composite_fwd_kernel(
    const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ z,
    int64_t rays_per_zrow, int64_t n_rays, int S, int flags,
    float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_acc) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool relu = flags & SNB_SIGMA_RELU, white = flags & SNB_WHITE_BKGD;
  for (int64_t ray = warp; ray < n_rays; ray += nwarps) {
    const float* sg = sigma + ray * S;
    const float* cg = rgb + ray * S * 3;
    const float* zr = z + (ray / rays_per_zrow) * S;
    float carry = 1.f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, A = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int k = base + lane;
      const bool valid = k < S;
      float s = 0.f, zk = 0.f, zn = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (valid) {
        s = __ldg(sg + k);
        zk = __ldg(zr + k);
        zn = (k + 1 < S) ? __ldg(zr + k + 1) : 0.f;
        c0 = __ldg(cg + 3 * k);
        c1 = __ldg(cg + 3 * k + 1);
        c2 = __ldg(cg + 3 * k + 2);
      }
      SampleTerms q = sample_terms(s, zk, zn, k == S - 1, relu);
      if (!valid) { q.alpha = 0.f; q.t = 1.f; }
      const float incl = warp_incl_prod(q.t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      const float T = carry * excl;
      const float w = q.alpha * T;
      ar += w * c0; ag += w * c1; ab += w * c2; ad += w * zk; aw += w;
      if (k == S - 1) A = T;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad); aw = warp_sum(aw);
    A = warp_sum(A);  // exactly one lane holds it
    if (lane == 0) {
      if (white) { ar = ar + 1.f - aw; ag = ag + 1.f - aw; ab = ab + 1.f - aw; }
      out_rgb[ray * 3 + 0] = ar; out_rgb[ray * 3 + 1] = ag; out_rgb[ray * 3 + 2] = ab;
      out_depth[ray] = ad;
      out_acc[ray] = A;
    }
  }
}

__global__ void __launch_bounds__(256) composite_bwd_kernel(
    const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ z,
    int64_t rays_per_zrow, int64_t n_rays, int S, int flags,
    const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_acc,
    float* __restrict__ g_sigma, float* __restrict__ g_rgbs, float* __restrict__ g_z) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool relu = flags & SNB_SIGMA_RELU, white = flags & SNB_WHITE_BKGD;
  const int nchunks = (S + 31) >> 5;
  float carries[kMaxChunks];  // transmittance at the start of each chunk (kept in local memory only for S > 32*regs)
  for (int64_t ray = warp; ray < n_rays; ray += nwarps) {
    const float* sg = sigma + ray * S;
    const float* cg = rgb + ray * S * 3;
    const float* zr = z + (ray / rays_per_zrow) * S;
    const float gc0 = __ldg(g_rgb + ray * 3), gc1 = __ldg(g_rgb + ray * 3 + 1), gc2 = __ldg(g_rgb + ray * 3 + 2);
    const float gD = __ldg(g_depth + ray), gA = __ldg(g_acc + ray);
    const float gsum = gc0 + gc1 + gc2;
    // pass 1: per-chunk starting transmittance and A = T_{S-1}
    float carry = 1.f, A = 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      const int k = c * 32 + lane;
      const bool valid = k < S;
      float s = 0.f, zk = 0.f, zn = 0.f;
      if (valid) { s = __ldg(sg + k); zk = __ldg(zr + k); zn = (k + 1 < S) ? __ldg(zr + k + 1) : 0.f; }
      SampleTerms q = sample_terms(s, zk, zn, k == S - 1, relu);
      if (!valid) q.t = 1.f;
      const float incl = warp_incl_prod(q.t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      carries[c] = carry;
      if (k == S - 1) A = carry * excl;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    A = warp_sum(A);
    const float gAA = gA * A;
    // pass 2: chunks in reverse; suffix = sum_{k>j} gw_k w_k as a reverse scan
    float suffix_carry = 0.f;  // sum over all samples of later chunks
#pragma unroll 1
    for (int c = nchunks - 1; c >= 0; --c) {
      const int k = c * 32 + lane;
      const bool valid = k < S;
      float s = 0.f, zk = 0.f, zn = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (valid) {
        s = __ldg(sg + k); zk = __ldg(zr + k); zn = (k + 1 < S) ? __ldg(zr + k + 1) : 0.f;
        c0 = __ldg(cg + 3 * k); c1 = __ldg(cg + 3 * k + 1); c2 = __ldg(cg + 3 * k + 2);
      }
      const bool last = (k == S - 1);
      SampleTerms q = sample_terms(s, zk, zn, last, relu);
      if (!valid) { q.alpha = 0.f; q.t = 1.f; }
      const float incl = warp_incl_prod(q.t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      const float T = carries[c] * excl;
      const float w = q.alpha * T;
      float gw = gc0 * c0 + gc1 * c1 + gc2 * c2 + gD * zk;
      if (white) gw -= gsum;
      const float x = valid ? gw * w : 0.f;
      const float rincl = warp_rev_incl_sum(x, lane);
      const float suffix = (rincl - x) + suffix_carry;  // strictly-later samples: later lanes + later chunks
      suffix_carry += __shfl_sync(0xffffffffu, rincl, 0);
      const float g_t = (suffix + (last ? 0.f : gAA)) / q.t;
      const float g_alpha = gw * T - g_t;
      float gs = g_alpha * q.delta * q.e;
      if (relu && !(s > 0.f)) gs = 0.f;
      const float gdel = (last || !valid) ? 0.f : g_alpha * q.sr * q.e;
      if (valid) {
        g_sigma[ray * S + k] = gs;
        g_rgbs[(ray * S + k) * 3 + 0] = w * gc0;
        g_rgbs[(ray * S + k) * 3 + 1] = w * gc1;
        g_rgbs[(ray * S + k) * 3 + 2] = w * gc2;
      }
      if (g_z != nullptr) {
        // g_z_k = w_k gD + g_delta_{k-1} - g_delta_k.  g_delta_{k-1} of lane 0 lives in chunk c-1, which is
        // processed next: its lane 31 adds it then (same warp, ordered by __syncwarp).
        float gprev = __shfl_up_sync(0xffffffffu, gdel, 1);
        if (lane == 0) gprev = 0.f;
        if (valid) g_z[ray * S + k] = w * gD + gprev - gdel;
        __syncwarp();
        if (lane == 31 && k + 1 < S) g_z[ray * S + k + 1] += gdel;
        __syncwarp();
      }
    }
  }
}

}  // namespace snb

using namespace snb;

static int composite_grid(int64_t n_rays, int warps_per_block) {
  int sms = sm_count();
  if (sms <= 0) return -1;
  int64_t blocks = ceil_div(n_rays, warps_per_block);
  int64_t cap = (int64_t)sms * 8;  // 8 resident 256-thread CTAs per SM = full occupancy, grid a multiple of the SM count
  return (int)(blocks < cap ? blocks : cap);
}

extern "C" int snb_composite_fwd(const float* sigma, const float* rgb, const float* z, int64_t rays_per_zrow,
                                 int64_t n_rays, int32_t n_samples, int32_t flags,
                                 float* out_rgb, float* out_depth, float* out_acc, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1 && rays_per_zrow >= 1, "composite_fwd: bad sizes");
  if (n_rays == 0) return 0;
  int grid = composite_grid(n_rays, 8);
  SNB_REQUIRE(grid > 0, "composite_fwd: no CUDA device");
  composite_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sigma, rgb, z, rays_per_zrow, n_rays, n_samples, flags,
                                                               out_rgb, out_depth, out_acc);
  SNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int snb_composite_bwd(const float* sigma, const float* rgb, const float* z, int64_t rays_per_zrow,
                                 int64_t n_rays, int32_t n_samples, int32_t flags,
                                 const float* g_rgb, const float* g_depth, const float* g_acc,
                                 float* g_sigma, float* g_rgbs, float* g_z, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1 && rays_per_zrow >= 1, "composite_bwd: bad sizes");
  SNB_REQUIRE(n_samples <= 32 * kMaxChunks, "composite_bwd: n_samples > %d unsupported", 32 * kMaxChunks);
  if (n_rays == 0) return 0;
  int grid = composite_grid(n_rays, 8);
  SNB_REQUIRE(grid > 0, "composite_bwd: no CUDA device");
  composite_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sigma, rgb, z, rays_per_zrow, n_rays, n_samples, flags,
                                                               g_rgb, g_depth, g_acc, g_sigma, g_rgbs, g_z);
  SNB_LAUNCH_CHECK();
  return 0;
}
