// Refine-iteration loss (the caller just above the render): the masked photometric loss and the
// exponential occupancy loss of optimizer_nuscenes.py:729-736 (same lines in optimizer_kitti.py / optimizer_waymo.py and
// trainer_unified_nuscenes.py:316-332), forward and backward, one launch each instead of ~25 elementwise/reduce launches.
//   den      = sum|occ| + 1e-9
//   loss_rgb = sum((rgb - tgt)^2 |occ|) / den            (the |occ| (N,1) broadcasts over the 3 channels)
//   loss_occ = sum(exp(-occ (0.5 - acc)) |occ|) / den
//   loss     = loss_rgb + coef * loss_occ
// HBM-bound: reads 32 B/ray forward, reads 32 B + writes 16 B/ray backward.
#include "common.cuh"
#include "../../include/supnerf_b200.h"

namespace snb {

struct LossAcc {      // device scratch, 64 bytes
  double num_rgb, num_occ, den;
  unsigned int ticket, pad;
  double den_final;   // den used by the forward (kept for the backward)
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

__global__ void __launch_bounds__(256) refine_loss_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ acc,
                                                             const float* __restrict__ tgt, const float* __restrict__ occ,
                                                             int64_t n, float coef, const float* __restrict__ den_in,
                                                             LossAcc* __restrict__ a, float* __restrict__ out3) {
  __shared__ double sh[8];
  double s_rgb = 0.0, s_occ = 0.0, s_den = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float o = __ldg(occ + i), ao = fabsf(o);
    const float d0 = __ldg(rgb + 3 * i) - __ldg(tgt + 3 * i), d1 = __ldg(rgb + 3 * i + 1) - __ldg(tgt + 3 * i + 1),
                d2 = __ldg(rgb + 3 * i + 2) - __ldg(tgt + 3 * i + 2);
    s_rgb += (double)((d0 * d0) * ao) + (double)((d1 * d1) * ao) + (double)((d2 * d2) * ao);
    s_occ += (double)(expf(-o * (0.5f - __ldg(acc + i))) * ao);
    s_den += (double)ao;
  }
  s_rgb = block_sum(s_rgb, sh);
  s_occ = block_sum(s_occ, sh);
  s_den = block_sum(s_den, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&a->num_rgb, s_rgb);
    atomicAdd(&a->num_occ, s_occ);
    atomicAdd(&a->den, s_den);
    __threadfence();
    const unsigned int t = atomicAdd(&a->ticket, 1u);
    if (t == gridDim.x - 1) {   // last block: every partial is visible
      __threadfence();
      const double nr = atomicAdd(&a->num_rgb, 0.0), no = atomicAdd(&a->num_occ, 0.0), dn = atomicAdd(&a->den, 0.0);
      // torch: den is an fp32 tensor (sum + 1e-9 in fp32); a caller-supplied den (ray-sharded mode) is used as is
      const float den = den_in ? __ldg(den_in) : (float)dn + 1e-9f;
      a->den_final = (double)den;
      const float lr = (float)(nr / (double)den), lo = (float)(no / (double)den);
      out3[0] = lr + coef * lo;
      out3[1] = lr;
      out3[2] = lo;
    }
  }
}

__global__ void __launch_bounds__(256) refine_loss_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ acc,
                                                             const float* __restrict__ tgt, const float* __restrict__ occ,
                                                             int64_t n, float coef, const LossAcc* __restrict__ a,
                                                             const float* __restrict__ g_loss, float* __restrict__ g_rgb,
                                                             float* __restrict__ g_acc) {
  const float g = g_loss ? __ldg(g_loss) : 1.f;
  const float inv_den = (float)(1.0 / a->den_final);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float o = __ldg(occ + i), ao = fabsf(o);
    const float s = 2.f * ao * inv_den * g;
    g_rgb[3 * i] = (__ldg(rgb + 3 * i) - __ldg(tgt + 3 * i)) * s;
    g_rgb[3 * i + 1] = (__ldg(rgb + 3 * i + 1) - __ldg(tgt + 3 * i + 1)) * s;
    g_rgb[3 * i + 2] = (__ldg(rgb + 3 * i + 2) - __ldg(tgt + 3 * i + 2)) * s;
    g_acc[i] = g * coef * expf(-o * (0.5f - __ldg(acc + i))) * o * ao * inv_den;
  }
}

}  // namespace snb

using namespace snb;

extern "C" size_t snb_refine_loss_scratch_bytes(void) { return sizeof(LossAcc); }

extern "C" int snb_refine_loss_fwd(const float* rgb, const float* acc, const float* tgt, const float* occ, int64_t n_rays,
                                   float occ_coef, const float* den, float* out3, void* scratch, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && out3 && scratch, "refine_loss_fwd: bad arguments");
  SNB_REQUIRE(n_rays == 0 || (rgb && acc && tgt && occ), "refine_loss_fwd: null pointer");
  const int sms = sm_count();
  SNB_REQUIRE(sms > 0, "refine_loss_fwd: no CUDA device (there is no CPU fallback)");
  cudaStream_t st = (cudaStream_t)stream;
  SNB_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(LossAcc), st));
  int64_t blocks = ceil_div(n_rays > 0 ? n_rays : 1, 256);
  const int grid = (int)(blocks < 2 * sms ? blocks : 2 * sms);
  refine_loss_fwd_kernel<<<grid, 256, 0, st>>>(rgb, acc, tgt, occ, n_rays, occ_coef, den, (LossAcc*)scratch, out3);
  SNB_LAUNCH_CHECK();
  return 0;
}

extern "C" int snb_refine_loss_bwd(const float* rgb, const float* acc, const float* tgt, const float* occ, int64_t n_rays,
                                   float occ_coef, const void* scratch, const float* g_loss, float* g_rgb, float* g_acc,
                                   void* stream) {
  SNB_REQUIRE(n_rays >= 0 && scratch, "refine_loss_bwd: bad arguments");
  if (n_rays == 0) return 0;
  SNB_REQUIRE(rgb && acc && tgt && occ && g_rgb && g_acc, "refine_loss_bwd: null pointer");
  const int sms = sm_count();
  SNB_REQUIRE(sms > 0, "refine_loss_bwd: no CUDA device (there is no CPU fallback)");
  int64_t blocks = ceil_div(n_rays, 256);
  const int grid = (int)(blocks < 8 * sms ? blocks : 8 * sms);
  refine_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, acc, tgt, occ, n_rays, occ_coef, (const LossAcc*)scratch,
                                                                g_loss, g_rgb, g_acc);
  SNB_LAUNCH_CHECK();
  return 0;
}
