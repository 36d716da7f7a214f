// Multi-object scene compositor, merge step: scripts/demo.py:560-567 of the reference --
//     z_sort  = torch.sort(z_vals, 1).values                  z_vals (R, K = Nb * S): every object's samples along one ray
//     z_args  = torch.searchsorted(z_sort, z_vals)            (left insertion index = number of strictly smaller depths)
//     rgbs_sort   = zeros.scatter_(1, z_args, rgbs)           ties (e.g. the z = -1 of rays that miss an object) collide:
//     sigmas_sort = zeros.scatter_(1, z_args, sigmas)         the LAST element in index order wins, the other slots stay 0
// followed by volume_rendering3(sigmas_sort, rgbs_sort, z_sort, white_bkgd=True) (the existing compositing kernel).
// One block per ray, one thread per sample: ranks by direct counting against the ray's depths in shared memory (K <= 1024, so
// K^2 <= 10^6 compares per ray, all shared-memory broadcasts).  Integer outputs (z_args) are bit-exact by construction.
#include "common.cuh"
#include "../../include/supnerf_b200.h"

namespace snb {

__global__ void __launch_bounds__(1024) merge_sort_samples_kernel(const float* __restrict__ z, const float* __restrict__ sigma,
                                                                 const float* __restrict__ rgb, int64_t n_rays, int K,
                                                                 float* __restrict__ z_sort, float* __restrict__ sigma_sort,
                                                                 float* __restrict__ rgb_sort, int64_t* __restrict__ z_args) {
  extern __shared__ float zs[];
  const int i = threadIdx.x;
  for (int64_t ray = blockIdx.x; ray < n_rays; ray += gridDim.x) {
    const int64_t base = ray * K;
    __syncthreads();
    if (i < K) {
      zs[i] = z[base + i];
      sigma_sort[base + i] = 0.f;
      rgb_sort[3 * (base + i)] = 0.f; rgb_sort[3 * (base + i) + 1] = 0.f; rgb_sort[3 * (base + i) + 2] = 0.f;
    }
    __syncthreads();
    if (i < K) {
      const float zi = zs[i];
      int less = 0, eq_before = 0, eq_after = 0;
      for (int j = 0; j < K; ++j) {
        const float zj = zs[j];
        less += zj < zi;
        const bool eq = zj == zi;
        eq_before += eq && j < i;
        eq_after += eq && j > i;
      }
      z_sort[base + less + eq_before] = zi;
      if (z_args) z_args[base + i] = less;
      if (eq_after == 0) {   // last of its tie group in index order: its values land in the group's first slot
        sigma_sort[base + less] = sigma[base + i];
        rgb_sort[3 * (base + less)] = rgb[3 * (base + i)];
        rgb_sort[3 * (base + less) + 1] = rgb[3 * (base + i) + 1];
        rgb_sort[3 * (base + less) + 2] = rgb[3 * (base + i) + 2];
      }
    }
  }
}

}  // namespace snb

using namespace snb;

extern "C" int snb_merge_sort_samples(const float* z, const float* sigma, const float* rgb, int64_t n_rays, int32_t n_per_ray,
                                      float* z_sort, float* sigma_sort, float* rgb_sort, int64_t* z_args, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_per_ray >= 1 && n_per_ray <= 1024, "merge_sort_samples: samples per ray must be in [1, 1024]");
  if (n_rays == 0) return 0;
  SNB_REQUIRE(z && sigma && rgb && z_sort && sigma_sort && rgb_sort, "merge_sort_samples: null pointer");
  const int sms = sm_count();
  SNB_REQUIRE(sms > 0, "merge_sort_samples: no CUDA device (there is no CPU fallback)");
  const int threads = (n_per_ray + 31) / 32 * 32;
  const int per_sm = 2048 / threads > 0 ? 2048 / threads : 1;
  const int64_t cap = (int64_t)sms * per_sm;
  const int grid = (int)(n_rays < cap ? n_rays : cap);
  merge_sort_samples_kernel<<<grid, threads, (size_t)n_per_ray * sizeof(float), (cudaStream_t)stream>>>(
      z, sigma, rgb, n_rays, n_per_ray, z_sort, sigma_sort, rgb_sort, z_args);
  SNB_LAUNCH_CHECK();
  return 0;
}
