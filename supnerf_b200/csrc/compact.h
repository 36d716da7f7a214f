// Miss-ray compaction helpers (compact.cu) and the internal tensor-core decoder entry points the fused render calls directly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "handle.h"
#include "rb_rows.cuh"

namespace snb {

int compact_plan(const uint8_t* hit, int64_t n_rays, int S, int32_t* order, int32_t* pos, int64_t* counts, cudaStream_t st);
int compact_gather(const float* xyz, const float* vrep, const int32_t* order, const int64_t* counts, int64_t n_rays, int S,
                   float* xyz_c, float* vrep_c, cudaStream_t st);
int compact_expand(const float* sigma_c, const float* rgb_c, const uint8_t* hit, const int32_t* pos, const int64_t* counts,
                   int64_t n_rays, int S, float* sigma, float* rgb, cudaStream_t st);
int compact_reduce(const float* g_sigma, const float* g_rgb, const int32_t* order, const int64_t* counts, int64_t n_rays, int S,
                   float* g_sigma_c, float* g_rgb_c, cudaStream_t st);
int compact_scatter(const float* g_xyz_c, const float* g_vrep_c, const uint8_t* hit, const int32_t* pos, const int64_t* counts,
                    int64_t n_rays, int S, float* g_xyz, float* g_vrep, cudaStream_t st);

// sampler.cu: box-sampler backward reading the decoder's input gradients in compact row order (no scatter to dense)
int sample_box_bwd_compact(const float* rays_o, const float* viewdir, const float* z_steps, const float* jitter, int64_t n_rays,
                           int32_t n_samples, float half_diag, const float* h, const float* g_xyz_c, const float* g_vrep_c,
                           const float* g_z_vals, const int32_t* pos, const int64_t* counts, float* g_rays_o, float* g_viewdir,
                           cudaStream_t st);

// mlp_tc.cu: m_dev (optional) = device-side number of rows actually present (a multiple of 128, <= M);
// tile_start (optional, B + 1 device ints, ascending, even): object b owns the 128-row tiles [tile_start[b], tile_start[b+1]) -- the
// batched render's objects own different numbers of rows (render_batch.cu)
bool tc_two_tile_active(const snb_handle_s* h);
bool tc_one_tile_supported(const snb_handle_s* h);   // the architecture the one-tile kernels (bf16 and split-precision) cover
int tc_forward(const snb_handle_s* h, const float* xyz, const float* viewdir, int64_t M, int64_t B,
               const float* shape_latent, const float* texture_latent, float* sigma, float* rgb, void* ws, cudaStream_t st, bool train,
               const int64_t* m_dev, const int32_t* tile_start = nullptr, const rb::RowSrc* rs = nullptr, bool split = false);
int tc_backward(const snb_handle_s* h, const float* xyz, const float* viewdir, int64_t M, int64_t B,
                const float* shape_latent, const float* texture_latent, const float* sigma, const float* g_sigma,
                const float* g_rgb, const void* ws, void* scratch, float* g_xyz, float* g_viewdir, float* g_shape_latent,
                float* g_texture_latent, float* const* g_weights, cudaStream_t st, bool train, const int64_t* m_dev,
                const int32_t* tile_start = nullptr, bool split = false);

}  // namespace snb
