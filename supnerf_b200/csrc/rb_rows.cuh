// Compact-row bookkeeping of the batched render (render_batch.cu), shared with the tcgen05 decoder (mlp_tc2.cu), whose forward
// computes its rows' sample coordinates itself (K1: no sampler kernel, no coordinates in HBM on the forward path).
#pragma once
#include "common.cuh"

namespace snb {
namespace rb {

struct ObjCounts { int64_t n_hit, n_miss, rows, row_start; };

// what the decoder's forward needs to find a row's ray and sample (all device pointers; rays8 == nullptr: coordinates are read instead)
struct RowSrc {
  const float* rays8;        // (B*N, 8) {o / (diag/2), d, near, far}
  const float* box;          // (B, 4), half diagonal first
  const float* z_steps;      // (S)
  const float* jitter;       // (B*N, S)
  const int32_t* order;      // (B*N) per-object local ray ids, hit rays first
  const ObjCounts* counts;   // (B)
  int64_t N;
  int32_t S;
};

// object of a global compact row (row_start is ascending; every object owns at least one row)
__device__ __forceinline__ int obj_of_row(const ObjCounts* __restrict__ counts, int B, int64_t row) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (counts[mid].row_start <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// (ray, sample) a compact row stands for; pad rows (behind the object's last row, up to the 256-row boundary) replay its first row
__device__ __forceinline__ void row_source(const ObjCounts& c, const int32_t* __restrict__ order, int64_t lr, int S, int64_t* ray, int* k) {
  if (lr >= c.rows) lr = 0;
  const int64_t hit_rows = c.n_hit * S;
  if (lr < hit_rows) {   // an object's rows fit 32 bits (checked on the host): 32-bit division
    const uint32_t l = (uint32_t)lr, q = l / (uint32_t)S;
    *ray = __ldg(order + q);
    *k = (int)(l - q * (uint32_t)S);
  } else {               // a miss ray's samples are one point: the last one carries the weight
    *ray = __ldg(order + c.n_hit + (lr - hit_rows));
    *k = S - 1;
  }
}

// the stratified sample (renderer.py:27-41 + :111-114), in the reference's op order
struct SamplePt { float x[3], d[3], zv; };
__device__ __forceinline__ SamplePt sample_point(const float* __restrict__ r8, float zstep, float jit, float fstep, float half_diag) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(r8)), b = __ldg(reinterpret_cast<const float4*>(r8) + 1);
  const float o[3] = {a.x, a.y, a.z};
  SamplePt s;
  s.d[0] = a.w; s.d[1] = b.x; s.d[2] = b.y;
  const float near = b.z, far = b.w;
  const float zs = __fadd_rn(zstep, __fmul_rn(jit, fstep));
  const float zc = __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, zs)), __fmul_rn(far, zs));
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s.x[i] = __fadd_rn(o[i], __fmul_rn(zc, s.d[i]));
    const float m = __fmul_rn(__fsub_rn(s.x[i], o[i]), half_diag);
    q = __fadd_rn(q, __fmul_rn(m, m));
  }
  s.zv = sqrtf(q);
  return s;
}


// z_vals of one sample given the ray's registers (the tail of sample_point)
__device__ __forceinline__ float sample_zv(const float o[3], const float d[3], float near, float far, float zstep, float jit, float fstep,
                                           float half_diag) {
  const float zs = __fadd_rn(zstep, __fmul_rn(jit, fstep));
  const float zc = __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, zs)), __fmul_rn(far, zs));
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float x = __fadd_rn(o[i], __fmul_rn(zc, d[i]));
    const float m = __fmul_rn(__fsub_rn(x, o[i]), half_diag);
    q = __fadd_rn(q, __fmul_rn(m, m));
  }
  return sqrtf(q);
}

}  // namespace rb
}  // namespace snb
