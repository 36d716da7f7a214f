// Per-object latent layers of the CodeNeRF-family decoder (model_codenerf.py:51,59: `*_latent_layer_j`), hoisted
// out of the per-sample path (exact restructuring, SURVEY 8(a')3).  One launch computes, for every object and
// every latent slot,  z = ReLU(W_lat latent + b_lat)  and the EFFECTIVE BIAS  W_layer z + b_layer  of the layer
// that consumes z (the latent add folded through that layer, used by the tcgen05 back end).  One launch does
// the backward (ReLU mask, then d latent = sum_slots W_lat^T d_pre).
#include "common.cuh"
#include "handle.h"

namespace snb {

constexpr int kMaxSlots = 32;
struct LatentLayers {
  int n_shape, n_total, W, D;
  const float* wl[kMaxSlots]; const float* bl[kMaxSlots];   // latent layers (W, D), (W)
  const float* wc[kMaxSlots]; const float* bc[kMaxSlots];   // consuming layers (W, W), (W)
};

// grid (slots, B), 256 threads.  zlat / ebias: [slot][b][W].
__global__ void __launch_bounds__(256) latent_fwd_kernel(const __grid_constant__ LatentLayers L, int64_t B,
                                                        const float* __restrict__ shape_latent,
                                                        const float* __restrict__ texture_latent,
                                                        float* __restrict__ zlat, float* __restrict__ ebias) {
  extern __shared__ float sh[];   // [D] latent, [W] z
  float* lat = sh;
  float* z = sh + L.D;
  const int slot = blockIdx.x;
  const int64_t b = blockIdx.y;
  const float* src = (slot < L.n_shape ? shape_latent : texture_latent) + b * L.D;
  for (int i = threadIdx.x; i < L.D; i += blockDim.x) lat[i] = src[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = warp; o < L.W; o += nw) {
    const float* w = L.wl[slot] + (size_t)o * L.D;
    float acc = 0.f;
    for (int k = lane; k < L.D; k += 32) acc = fmaf(__ldg(w + k), lat[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = fmaxf(acc + __ldg(L.bl[slot] + o), 0.f);
      z[o] = v;
      zlat[((size_t)slot * B + b) * L.W + o] = v;
    }
  }
  if (ebias == nullptr) return;
  __syncthreads();
  for (int o = warp; o < L.W; o += nw) {
    const float* w = L.wc[slot] + (size_t)o * L.W;
    float acc = 0.f;
    for (int k = lane; k < L.W; k += 32) acc = fmaf(__ldg(w + k), z[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) ebias[((size_t)slot * B + b) * L.W + o] = acc + __ldg(L.bc[slot] + o);
  }
}

// grid (2, B): type 0 = shape slots, 1 = texture slots.  dz [slot][b][W] = d loss / d z (post-ReLU).
__global__ void __launch_bounds__(256) latent_bwd_kernel(const __grid_constant__ LatentLayers L, int64_t B,
                                                        const float* __restrict__ zlat, const float* __restrict__ dz,
                                                        float* __restrict__ g_shape, float* __restrict__ g_texture) {
  extern __shared__ float sh[];   // [W] masked gradient of the current slot
  const int type = blockIdx.x;
  const int64_t b = blockIdx.y;
  float* out = type == 0 ? g_shape : g_texture;
  if (out == nullptr) return;
  const int s0 = type == 0 ? 0 : L.n_shape, s1 = type == 0 ? L.n_shape : L.n_total;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // thread t owns latent units t, t+256, ... (D <= 1024)
  for (int slot = s0; slot < s1; ++slot) {
    __syncthreads();
    for (int o = threadIdx.x; o < L.W; o += blockDim.x) {
      const size_t i = ((size_t)slot * B + b) * L.W + o;
      sh[o] = zlat[i] > 0.f ? dz[i] : 0.f;
    }
    __syncthreads();
    for (int o = 0; o < L.W; ++o) {
      const float g = sh[o];
      const float* w = L.wl[slot] + (size_t)o * L.D;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = threadIdx.x + u * 256;
        if (k < L.D) acc[u] = fmaf(g, __ldg(w + k), acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int k = threadIdx.x + u * 256;
    if (k < L.D) out[b * L.D + k] = acc[u];
  }
}

static int fill_layers(const snb_handle_s* h, LatentLayers& L) {
  L.n_shape = h->arch.shape_blocks;
  L.n_total = h->arch.shape_blocks + h->arch.texture_blocks;
  L.W = h->arch.W; L.D = h->arch.latent_dim;
  SNB_REQUIRE(L.n_total <= kMaxSlots && L.D <= 1024, "latent layers: too many blocks or latent_dim > 1024");
  for (int j = 1; j <= L.n_total; ++j) {
    const bool s = j <= L.n_shape;
    const int ll = s ? h->iSL(j) : h->iTL(j - L.n_shape), lc = s ? h->iS(j) : h->iT(j - L.n_shape);
    L.wl[j - 1] = h->layers[ll].w; L.bl[j - 1] = h->layers[ll].b;
    L.wc[j - 1] = h->layers[lc].w; L.bc[j - 1] = h->layers[lc].b;
  }
  return 0;
}

int latent_forward_fused(const snb_handle_s* h, int64_t B, const float* shape_latent, const float* texture_latent,
                         float* zlat, float* ebias, cudaStream_t st) {
  LatentLayers L;
  if (fill_layers(h, L)) return 2;
  dim3 grid(L.n_total, (unsigned)B);
  latent_fwd_kernel<<<grid, 256, (L.D + L.W) * sizeof(float), st>>>(L, B, shape_latent, texture_latent, zlat, ebias);
  SNB_LAUNCH_CHECK();
  return 0;
}

int latent_backward_fused(const snb_handle_s* h, int64_t B, const float* zlat, const float* dz, float* g_shape_latent,
                          float* g_texture_latent, cudaStream_t st) {
  LatentLayers L;
  if (fill_layers(h, L)) return 2;
  dim3 grid(2, (unsigned)B);
  latent_bwd_kernel<<<grid, 256, L.W * sizeof(float), st>>>(L, B, zlat, dz, g_shape_latent, g_texture_latent);
  SNB_LAUNCH_CHECK();
  return 0;
}

}  // namespace snb
