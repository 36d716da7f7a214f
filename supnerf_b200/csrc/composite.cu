// K3 / K3b — alpha compositing along rays as a warp-level segmented scan, forward and reverse-scan
// backward.  Replaces renderer.py:43-65, :355-379 and utils.py:187-233 of the reference.
//
// One warp owns one ray at a time; lane l owns sample (chunk*32 + l).  Transmittance is an exclusive
// product scan across lanes (5 shuffles) carried across 32-sample chunks; the backward recomputes
// the forward quantities and runs the suffix sum as a reverse scan (no total-minus-prefix
// cancellation).  HBM-bound: 20*S+20 B/ray forward, 40*S+20 read + 20*S written backward.
#include "common.cuh"
#include "composite.cuh"
#include "../../include/supnerf_b200.h"

namespace snb {

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


// ---------------------------------------------------------------------------------------------------------------
// Fast path (S % 4 == 0, S <= 128, 16-byte aligned rows): LPR lanes per ray, 4 consecutive samples per lane in
// registers (128-bit loads), so one warp instruction serves 32/LPR rays x 4 samples and the per-ray scan/reduction
// shuffles are log2(LPR) deep.  ncu on the one-sample-per-lane kernels above showed them ISSUE-bound (65 % issue
// slots, 33 % of HBM peak): 325 warp instructions per 64-sample ray forward, 667 backward.
template <int LPR>
__global__ void __launch_bounds__(256) composite_fwd4_kernel(
    const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ z,
    int64_t rays_per_zrow, int64_t n_rays, int S, int flags,
    float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_acc) {
  constexpr int RPW = 32 / LPR;   // rays per warp
  const int lane = threadIdx.x & 31, sl = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool relu = flags & SNB_SIGMA_RELU, white = flags & SNB_WHITE_BKGD;
  const int k0 = sl * 4;
  for (int64_t ray0 = warp * RPW; ray0 < n_rays; ray0 += nwarps * RPW) {
    const int64_t ray = ray0 + sub;
    const bool rv = ray < n_rays;
    const int64_t rr = rv ? ray : n_rays - 1;
    const Lane4 L = load_lane4<LPR>(sigma + rr * S, rgb + rr * S * 3, z + (rr / rays_per_zrow) * S, k0, S);
    float al[4], tl[4], tp = 1.f;   // tl[j]: product of t over the lane's samples before j
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      SampleTerms q = sample_terms(L.s[j], L.z[j], L.z[j + 1], k0 + j == S - 1, relu);
      if (!L.valid) { q.alpha = 0.f; q.t = 1.f; }
      al[j] = q.alpha; tl[j] = tp; tp *= q.t;
    }
    float total;
    const float T0 = seg_excl_prod<LPR>(tp, sl, &total);
    float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, A = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float T = T0 * tl[j];
      const float w = al[j] * T;
      ar += w * L.c[j][0]; ag += w * L.c[j][1]; ab += w * L.c[j][2]; ad += w * L.z[j]; aw += w;
      if (k0 + j == S - 1) A = T;
    }
    ar = seg_sum<LPR>(ar); ag = seg_sum<LPR>(ag); ab = seg_sum<LPR>(ab); ad = seg_sum<LPR>(ad); aw = seg_sum<LPR>(aw);
    A = __shfl_sync(0xffffffffu, A, (S - 1) / 4, LPR);   // the lane that owns the last sample
    if (sl == 0 && rv) {
      if (white) { ar = ar + 1.f - aw; ag = ag + 1.f - aw; ab = ab + 1.f - aw; }
      out_rgb[ray * 3 + 0] = ar; out_rgb[ray * 3 + 1] = ag; out_rgb[ray * 3 + 2] = ab;
      out_depth[ray] = ad;
      out_acc[ray] = A;
    }
  }
}

template <int LPR>
__global__ void __launch_bounds__(256) composite_bwd4_kernel(
    const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ z,
    int64_t rays_per_zrow, int64_t n_rays, int S, int flags,
    const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_acc,
    float* __restrict__ g_sigma, float* __restrict__ g_rgbs, float* __restrict__ g_z) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sl = lane % LPR, sub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool relu = flags & SNB_SIGMA_RELU, white = flags & SNB_WHITE_BKGD;
  const int k0 = sl * 4;
  for (int64_t ray0 = warp * RPW; ray0 < n_rays; ray0 += nwarps * RPW) {
    const int64_t ray = ray0 + sub;
    const bool rv = ray < n_rays;
    const int64_t rr = rv ? ray : n_rays - 1;
    const Lane4 L = load_lane4<LPR>(sigma + rr * S, rgb + rr * S * 3, z + (rr / rays_per_zrow) * S, k0, S);
    const float gc0 = __ldg(g_rgb + rr * 3), gc1 = __ldg(g_rgb + rr * 3 + 1), gc2 = __ldg(g_rgb + rr * 3 + 2);
    const float gD = __ldg(g_depth + rr), gA = __ldg(g_acc + rr);
    const float gsum = gc0 + gc1 + gc2;
    SampleTerms q[4];
    float tl[4], tp = 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = sample_terms(L.s[j], L.z[j], L.z[j + 1], k0 + j == S - 1, relu);
      if (!L.valid) { q[j].alpha = 0.f; q[j].t = 1.f; }
      tl[j] = tp; tp *= q[j].t;
    }
    float total;
    const float T0 = seg_excl_prod<LPR>(tp, sl, &total);
    float T[4], w[4], gw[4], x[4], A = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T[j] = T0 * tl[j];
      w[j] = q[j].alpha * T[j];
      gw[j] = gc0 * L.c[j][0] + gc1 * L.c[j][1] + gc2 * L.c[j][2] + gD * L.z[j];
      if (white) gw[j] -= gsum;
      x[j] = L.valid ? gw[j] * w[j] : 0.f;
      if (k0 + j == S - 1) A = T[j];
    }
    A = __shfl_sync(0xffffffffu, A, (S - 1) / 4, LPR);
    const float gAA = gA * A;
    // suffix_j = sum_{k > j} gw_k w_k: later samples of this lane, then every later lane (reverse scan, no total-minus-prefix)
    const float later = seg_rev_excl_sum<LPR>((x[0] + x[1]) + (x[2] + x[3]), sl);
    float suf[4];
    suf[3] = later; suf[2] = suf[3] + x[3]; suf[1] = suf[2] + x[2]; suf[0] = suf[1] + x[1];
    float gs[4], gdel[4], gcol[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool last = (k0 + j == S - 1);
      const float g_t = (suf[j] + (last ? 0.f : gAA)) / q[j].t;
      const float g_alpha = gw[j] * T[j] - g_t;
      gs[j] = g_alpha * q[j].delta * q[j].e;
      if (relu && !(L.s[j] > 0.f)) gs[j] = 0.f;
      gdel[j] = (last || !L.valid) ? 0.f : g_alpha * q[j].sr * q[j].e;
      gcol[j][0] = w[j] * gc0; gcol[j][1] = w[j] * gc1; gcol[j][2] = w[j] * gc2;
    }
    float gprev = __shfl_up_sync(0xffffffffu, gdel[3], 1, LPR);   // g_delta of the sample just before this lane's first
    if (sl == 0) gprev = 0.f;
    if (L.valid && rv) {
      *reinterpret_cast<float4*>(g_sigma + ray * S + k0) = make_float4(gs[0], gs[1], gs[2], gs[3]);
      float* gr = g_rgbs + (ray * S + k0) * 3;
      *reinterpret_cast<float4*>(gr) = make_float4(gcol[0][0], gcol[0][1], gcol[0][2], gcol[1][0]);
      *reinterpret_cast<float4*>(gr + 4) = make_float4(gcol[1][1], gcol[1][2], gcol[2][0], gcol[2][1]);
      *reinterpret_cast<float4*>(gr + 8) = make_float4(gcol[2][2], gcol[3][0], gcol[3][1], gcol[3][2]);
      if (g_z != nullptr) {
        // g_z_k = w_k gD + g_delta_{k-1} - g_delta_k
        *reinterpret_cast<float4*>(g_z + ray * S + k0) =
            make_float4(w[0] * gD + gprev - gdel[0], w[1] * gD + gdel[0] - gdel[1], w[2] * gD + gdel[1] - gdel[2],
                        w[3] * gD + gdel[2] - gdel[3]);
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
  const bool fast = n_samples % 4 == 0 && n_samples <= 128 && n_samples >= 4 &&
                    ((((uintptr_t)sigma | (uintptr_t)rgb | (uintptr_t)z) & 15) == 0);
  if (fast) {
    const int lpr = n_samples <= 32 ? 8 : (n_samples <= 64 ? 16 : 32);
    int gridf = composite_grid(ceil_div(n_rays, 32 / lpr), 8);
    SNB_REQUIRE(gridf > 0, "composite_fwd: no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
#define SNB_FWD4(L) composite_fwd4_kernel<L><<<gridf, 256, 0, st>>>(sigma, rgb, z, rays_per_zrow, n_rays, n_samples, flags, out_rgb, out_depth, out_acc)
    if (lpr == 8) SNB_FWD4(8); else if (lpr == 16) SNB_FWD4(16); else SNB_FWD4(32);
#undef SNB_FWD4
    SNB_LAUNCH_CHECK();
    return 0;
  }
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
  const bool fast = n_samples % 4 == 0 && n_samples <= 128 && n_samples >= 4 &&
                    ((((uintptr_t)sigma | (uintptr_t)rgb | (uintptr_t)z | (uintptr_t)g_sigma | (uintptr_t)g_rgbs | (uintptr_t)g_z) & 15) == 0);
  if (fast) {
    const int lpr = n_samples <= 32 ? 8 : (n_samples <= 64 ? 16 : 32);
    int gridf = composite_grid(ceil_div(n_rays, 32 / lpr), 8);
    SNB_REQUIRE(gridf > 0, "composite_bwd: no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
#define SNB_BWD4(L) composite_bwd4_kernel<L><<<gridf, 256, 0, st>>>(sigma, rgb, z, rays_per_zrow, n_rays, n_samples, flags, g_rgb, g_depth, g_acc, g_sigma, g_rgbs, g_z)
    if (lpr == 8) SNB_BWD4(8); else if (lpr == 16) SNB_BWD4(16); else SNB_BWD4(32);
#undef SNB_BWD4
    SNB_LAUNCH_CHECK();
    return 0;
  }
  int grid = composite_grid(n_rays, 8);
  SNB_REQUIRE(grid > 0, "composite_bwd: no CUDA device");
  composite_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sigma, rgb, z, rays_per_zrow, n_rays, n_samples, flags,
                                                               g_rgb, g_depth, g_acc, g_sigma, g_rgbs, g_z);
  SNB_LAUNCH_CHECK();
  return 0;
}
