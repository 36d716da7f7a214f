// fp32 SIMT GEMM building block of the SNB_PREC_FP32 parity mode (FFMA, 128x128x16 tiles, 8x8 per
// thread).  C(i,j) = epi( sum_k A'(i,k) * B'(k,j) ), with strided operands so the same kernel serves
//   forward      Y  = (X [+ rowadd]) W^T     (A k-contiguous, B k-contiguous)
//   data grad    dX = dY W                   (A k-contiguous, B j-contiguous)
//   weight grad  dW = dY^T (X [+ rowadd])    (A i-contiguous, B j-contiguous, split over the row dim, atomics)
#pragma once
#include "common.cuh"

namespace snb {

struct GemmArgs {
  const float* A; int64_t sAi, sAk;   // A'(i,k) = A[i*sAi + k*sAk]
  const float* B; int64_t sBk, sBj;   // B'(k,j) = B[k*sBk + j*sBj]
  float* C; int64_t ldc;              // C(i,j)  = C[i*ldc + j]
  int64_t M; int N; int64_t K;        // i < M, j < N, k < K
  const float* a_add; int64_t a_add_ld; int64_t a_rows_per_obj;  // A'(i,k) += a_add[(i / rpo)*ld + k]   (forward)
  const float* b_add; int64_t b_add_ld; int64_t b_rows_per_obj;  // B'(k,j) += b_add[(k / rpo)*ld + j]   (wgrad: k is the row)
  const float* bias;                  // + bias[j]
  int accumulate;                     // C += (non-atomic)
  int act;                            // 0 none, 1 relu
  int64_t k_split;                    // >0: each blockIdx.z handles k_split of K and atomically adds into C
};

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  constexpr int BM = 128, BN = 128, BK = 16;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * BM;   // M tiles on grid.x (up to 2^31-1: M reaches 10^7 rows), N tiles on grid.y
  const int j0 = blockIdx.y * BN;
  int64_t kbeg = 0, kend = g.K;
  if (g.k_split > 0) {
    kbeg = (int64_t)blockIdx.z * g.k_split;
    kend = kbeg + g.k_split < g.K ? kbeg + g.k_split : g.K;
    if (kbeg >= kend) return;
  }
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile: 128 x 16
    if (A_KC) {
      const int i = t >> 1, kk0 = (t & 1) * 8;
      const int64_t gi = i0 + i;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t gk = k0 + kk0 + u;
        float v = 0.f;
        if (gi < g.M && gk < kend) {
          v = __ldg(g.A + gi * g.sAi + gk * g.sAk);
          if (g.a_add) v += __ldg(g.a_add + (gi / g.a_rows_per_obj) * g.a_add_ld + gk);
        }
        As[kk0 + u][i] = v;
      }
    } else {
      const int kk = t >> 4, ii0 = (t & 15) * 8;
      const int64_t gk = k0 + kk;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t gi = i0 + ii0 + u;
        float v = 0.f;
        if (gi < g.M && gk < kend) v = __ldg(g.A + gi * g.sAi + gk * g.sAk);
        As[kk][ii0 + u] = v;
      }
    }
    // ---- load B tile: 16 x 128
    if (B_KC) {
      const int j = t >> 1, kk0 = (t & 1) * 8;
      const int gj = j0 + j;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t gk = k0 + kk0 + u;
        float v = 0.f;
        if (gj < g.N && gk < kend) v = __ldg(g.B + gk * g.sBk + (int64_t)gj * g.sBj);
        Bs[kk0 + u][j] = v;
      }
    } else {
      const int kk = t >> 4, jj0 = (t & 15) * 8;
      const int64_t gk = k0 + kk;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int gj = j0 + jj0 + u;
        float v = 0.f;
        if (gj < g.N && gk < kend) {
          v = __ldg(g.B + gk * g.sBk + (int64_t)gj * g.sBj);
          if (g.b_add) v += __ldg(g.b_add + (gk / g.b_rows_per_obj) * g.b_add_ld + gj);
        }
        Bs[kk][jj0 + u] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t gi = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
    if (gi >= g.M) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int gj = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
      if (gj >= g.N) continue;
      float v = acc[a][b];
      float* c = g.C + gi * g.ldc + gj;
      if (g.k_split > 0) { atomicAdd(c, v); continue; }
      if (g.bias) v += __ldg(g.bias + gj);
      if (g.accumulate) v += *c;
      if (g.act == 1) v = fmaxf(v, 0.f);
      *c = v;
    }
  }
}

inline int launch_sgemm(const GemmArgs& g, bool a_kc, bool b_kc, cudaStream_t st) {
  if (g.M == 0 || g.N == 0) return 0;
  int64_t splits = 1;
  if (g.k_split > 0) splits = ceil_div(g.K, g.k_split);
  dim3 grid((unsigned)ceil_div(g.M, 128), (unsigned)ceil_div(g.N, 128), (unsigned)splits);
  if (a_kc && b_kc) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (a_kc && !b_kc) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (!a_kc && !b_kc) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  else sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
  SNB_LAUNCH_CHECK();
  return 0;
}

}  // namespace snb
