// PTX wrappers and small device helpers shared by the tcgen05 decoder kernels (mlp_tc.cu: one tile per CTA;
// mlp_tc2.cu: two tiles in flight per CTA).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

namespace snb {
namespace tc {

constexpr int kTileM = 128;
constexpr uint32_t kChunkBytes = 16384;   // [128 m][64 k] bf16, 128-byte swizzled, K-major

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// issue a 32-column TMEM load of this thread's lane (no wait)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// issue a 16-column TMEM load of this thread's lane (no wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// legacy warp-level tensor-core ops, used for the column sums of a shared-memory operand tile (ones^T x tile)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void mma_ones_bf16(float (&d)[4], uint32_t b0, uint32_t b1) {
  const uint32_t one2 = 0x3F803F80u;   // bf16x2 (1.0, 1.0)
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(one2), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  tmem_ld32_issue(taddr, r);
  tmem_ld_wait();
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows 128 B apart, 8-row atoms 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t umma_idesc(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// the same with A/B fp16 (format field 0): the split-precision mode's operands
__device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t M, uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

// two fp32 values -> fp16x2 words of their leading parts and of the remainders: x = hi + lo up to 2^-22 |x| (lo normal) / 2^-25 (subnormal)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// relu + round-to-nearest pack of two fp32 into one bf16x2 word (first argument -> low half = lower column)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// ReLU mask words hold element i of a 32-column group at bit (31 - i): the forward builds them with one funnel shift per
// element from the sign bit of the pre-activation (1 = active).
__device__ __forceinline__ bool mask_bit(uint32_t mw, int i) { return (mw >> (31 - i)) & 1u; }

// bf16x2 AND-mask of the element pair (e, e + 1), e even, of a ReLU mask word: 0xFFFF in the low half iff element e is active,
// in the high half iff element e + 1 is.  The bit of element e sits at 31 - e = the MSB of byte 3 - (e >> 3) of (mw << (e & 7)):
// ONE prmt in sign-replicate mode (selector nibble 8 | byte) expands both bits, against a bit test + select per element.
// sh[t] = mw << t, t = 0 .. 7 (shared by the 16 pairs of a mask word).
template <int E>
__device__ __forceinline__ uint32_t mask_pair(const uint32_t (&sh)[8]) {
  static_assert(E % 2 == 0 && E >= 0 && E < 32, "even element index inside a 32-column mask word");
  constexpr uint32_t j = 3u - (uint32_t)(E >> 3);
  constexpr uint32_t sel = (8u | j) | ((8u | j) << 4) | ((8u | (4u + j)) << 8) | ((8u | (4u + j)) << 12);
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(sh[E & 7]), "r"(sh[(E & 7) + 1]), "n"(sel));
  return d;
}
__device__ __forceinline__ void mask_shifts(uint32_t mw, uint32_t (&sh)[8]) {
#pragma unroll
  for (int t = 0; t < 8; ++t) sh[t] = mw << t;
}

// byte offset of 16-byte unit `unit` (0..7) of row `row` inside a 128B-swizzled [rows][64] bf16 chunk
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t unit) { return row * 128u + ((unit ^ (row & 7u)) << 4); }

// sin/cos(2^f x), f < DEG, by the double-angle recurrence from one accurate sincosf (error ~2^f * 1e-7: far below bf16), or --
// ACCURATE, the split-precision (fp32-grade) mode -- one sincosf per frequency (2^f x is exact in fp32)
template <int DEG, bool ACCURATE = false>
__device__ __forceinline__ void trig_ladder(const float x[3], float s[DEG][3], float c[DEG][3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float sv, cv;
    sincosf(x[a], &sv, &cv);
    s[0][a] = sv; c[0][a] = cv;
#pragma unroll
    for (int f = 1; f < DEG; ++f) {
      if (ACCURATE) {
        sincosf(x[a] * (float)(1 << f), &sv, &cv);
        s[f][a] = sv; c[f][a] = cv;
      } else {
        s[f][a] = 2.f * s[f - 1][a] * c[f - 1][a];
        c[f][a] = 1.f - 2.f * s[f - 1][a] * s[f - 1][a];
      }
    }
  }
}

// PE(x) (model_codenerf.py:4-10 column order) as one bf16 row of 64 columns; this thread owns the 4 units of its column half
// `hh` (columns 32*hh .. 32*hh+31); columns >= 3+6*DEG are zero.  pe_half computes them (16 packed words), store_pe_half writes
// them into a 128B-swizzled chunk -- split so that the arithmetic can run before the chunk is free.
template <int DEG, bool ACCURATE = false>
__device__ __forceinline__ void pe_row_f32(const float x[3], float (&v)[64]) {
  float s[DEG][3], c[DEG][3];
  trig_ladder<DEG, ACCURATE>(x, s, c);
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a] = x[a];
#pragma unroll
  for (int f = 0; f < DEG; ++f)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      v[3 + 3 * f + a] = s[f][a];
      v[3 + 3 * DEG + 3 * f + a] = c[f][a];
    }
}
template <int DEG>
__device__ __forceinline__ void pe_half(const float x[3], uint32_t hh, uint32_t (&pk)[16]) {
  float v[64];
  pe_row_f32<DEG>(x, v);
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = hh ? pack_bf16(v[32 + 2 * i], v[33 + 2 * i]) : pack_bf16(v[2 * i], v[2 * i + 1]);
}
// the fp32 values of this thread's column half (split-precision mode packs them itself)
template <int DEG>
__device__ __forceinline__ void pe_half_f32(const float x[3], uint32_t hh, float (&o)[32]) {
  float v[64];
  pe_row_f32<DEG, true>(x, v);
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = hh ? v[32 + i] : v[i];
}
__device__ __forceinline__ void store_pe_half(uint8_t* aux, uint32_t row, uint32_t hh, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
    *reinterpret_cast<uint4*>(aux + swz(row, hh * 4u + (uint32_t)u)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
}
template <int DEG>
__device__ __forceinline__ void write_pe_row(uint8_t* aux, uint32_t row, uint32_t hh, const float x[3]) {
  uint32_t pk[16];
  pe_half<DEG>(x, hh, pk);
  store_pe_half(aux, row, hh, pk);
}

// One 32-byte row of a bias stage: both 16-byte units hold {hi, mid, lo, 0, 0, 0, 0, 0}, the bf16 split of bias / 2, so the
// row is invariant under the 32-byte swizzle and  ones(16) . row = bias  to fp32 accuracy.
__device__ __forceinline__ void bias_stage_row(uint8_t* row, float bias) {
  const float x = 0.5f * bias;
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  uint4 q;
  q.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
  q.y = (uint32_t)__bfloat16_as_ushort(lo);
  q.z = 0u; q.w = 0u;
  reinterpret_cast<uint4*>(row)[0] = q;
  reinterpret_cast<uint4*>(row)[1] = q;
}

// sum over the 32 lanes (= 32 rows) of each of the 32 per-lane values: lane l ends with column l's sum in v[0]
__device__ __forceinline__ float warp_colsum32(float (&v)[32], uint32_t lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = upper ? v[i + o] : v[i];
      const float send = upper ? v[i] : v[i + o];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

}  // namespace tc
}  // namespace snb
