// K2 / K2b, SNB_PREC_BF16 back end: the CodeNeRF-family decoder as ONE persistent, warp-specialised
// tcgen05 kernel per direction.  A 128-sample tile enters as fp32 xyz / viewdir; the positional
// encoding is built straight into shared memory as bf16 (never to HBM); every layer is a
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM, M=128, N<=256) whose A operand is the previous layer's
// epilogue output kept in shared memory and whose B operand (pre-tiled, pre-swizzled bf16 weight
// images) is streamed from L2 by the bulk-copy engine (cp.async.bulk + mbarrier) through a 4-stage
// ring.  Accumulators ping-pong between the two 256-column halves of TMEM so that the epilogue of
// layer l (bias, ReLU, latent add, mask bits, bf16 pack) overlaps the MMAs of layer l+1 chunk by
// chunk.  The 256->1 (sigma) and 128->3 (rgb) heads run on CUDA cores inside the epilogues.
// Backward: same machinery with transposed weight images; ReLU masks come back as 1 bit/unit;
// latent gradients are per-object column sums reduced by warp butterflies; d(PE) is folded to
// d xyz / d viewdir in the epilogue.
//
// Warp roles (320 threads): warps 0-7 epilogue (warp w owns TMEM lanes 32*(w%4).. and the column
// half w/4 of every 64-column chunk), warp 8 weight producer (+ TMEM alloc), warp 9 MMA issuer.
#include "common.cuh"
#include "handle.h"
#include "tc_ptx.cuh"
#include "rb_rows.cuh"
#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <initializer_list>
#include <utility>
#include <vector>

namespace snb {
namespace tc {

constexpr int kStages = 3;
constexpr uint32_t kStageBytes = 32768;   // [256 n][64 k] bf16
constexpr int kMaxSteps = 16;
constexpr int kMaxFwdSteps = 12;
constexpr int kThreads = 320;
constexpr int kMaxLatentSlots = 8;

// shared-memory map (bytes from the 1024-aligned base)
constexpr uint32_t SM_W = 5 * kChunkBytes;                       // A chunks 0..3, AUX chunk 4, then the weight ring
constexpr uint32_t SM_TAB = SM_W + kStages * kStageBytes;        // fp32 tables
constexpr uint32_t TAB_BIAS = 0;                                 // fwd: [kMaxFwdSteps][256]   bwd: colsum [8][256] at the same place
constexpr uint32_t TAB_Z = TAB_BIAS + kMaxFwdSteps * 256 * 4;    // fwd: [8][256] latent vectors of the current object
constexpr uint32_t TAB_WSIG = TAB_Z + kMaxLatentSlots * 256 * 4; // [256]
constexpr uint32_t TAB_W2 = TAB_WSIG + 256 * 4;                  // [3][128]
constexpr uint32_t TAB_PART = TAB_W2 + 384 * 4;                  // fwd: sig_part [2][128] + rgb_part [2][128][3]; bwd: xyz_part [2][128][3]
constexpr uint32_t TAB_BYTES = TAB_PART + (256 + 768) * 4;
constexpr uint32_t SM_BARS = SM_TAB + TAB_BYTES;
constexpr uint32_t SM_TOTAL = SM_BARS + 32 * 8 + 16;
constexpr uint32_t SM_ALLOC = SM_TOTAL + 1024;                   // + alignment slack
static_assert(SM_ALLOC <= 232448, "shared memory budget");

enum Epi : int { EPI_RELU = 0, EPI_LINEAR_SIGMA = 1, EPI_RGB_HEAD = 2, EPI_B_MASK = 3, EPI_B_EV = 4, EPI_B_XYZ = 5 };

struct Step {
  uint32_t w_off, w2_off;      // byte offsets of the first weight chunk of MMA group 1 / 2 in the packed buffer
  uint16_t n_chunks, n_out;    // K chunks (64 wide), N of group 1
  uint16_t n2_out;             // N of group 2 (0: none); group 2 re-reads the same A chunks
  uint8_t a_chunk[6];          // shared-memory chunk index per K chunk (0..3 = A, 4 = AUX)
  int8_t epi, mask_slot, latent_slot, produce_a;
  const float* bias;           // fp32 bias of the layer (forward)
};

struct Program {
  int n_steps, n_mask_slots;
  Step s[kMaxSteps];
};

struct Params {
  const float* xyz; const float* viewdir;
  int64_t M, rows_per_obj, B;
  const uint8_t* packed;
  const float* zlat;            // fwd: per-object effective biases b + W z, [(Bs+Bt)][B][256]; bwd: unused
  const float* wsig; const float* bsig; const float* w2; const float* b2;
  uint32_t* masks;              // [tile][slot][8 words][128 rows]
  // forward
  float* sigma; float* rgb; float* dbg;
  // backward
  const float* sigma_in; const float* g_sigma; const float* g_rgb;
  float* g_xyz; float* g_viewdir; float* g_zlat;   // g_zlat [(Bs+Bt)][B][256], accumulated with atomics
  int r0_mask_slot, n_latent, ev_step;
  Program prog;
};

struct Smem {
  uint8_t* base;
  uint32_t base_u32;
  __device__ uint8_t* chunk(int c) const { return base + (uint32_t)c * kChunkBytes; }
  __device__ uint32_t chunk_u32(int c) const { return base_u32 + (uint32_t)c * kChunkBytes; }
  __device__ uint32_t stage_u32(int s) const { return base_u32 + SM_W + (uint32_t)s * kStageBytes; }
  __device__ float* tab(uint32_t off) const { return reinterpret_cast<float*>(base + SM_TAB + off); }
  __device__ uint32_t bar(int i) const { return base_u32 + SM_BARS + 8u * i; }
};
// barrier indices
constexpr int BAR_WFULL = 0, BAR_WEMPTY = 4, BAR_AREADY = 8, BAR_ACC = 13;

// per-thread view of the epilogue role
struct EpiCtx {
  uint32_t lane, hh, row, lane_field, tmem_base;
  int64_t grow;      // global sample row of this thread
  bool valid;
};

// ------------------------------------------------------------------------------------------ roles shared by fwd / bwd
__device__ __forceinline__ void producer_loop(const Params& p, const Smem& sm, int64_t n_tiles, uint32_t lane) {
  uint32_t it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int si = 0; si < p.prog.n_steps; ++si) {
      const Step& st = p.prog.s[si];
      const int groups = st.n2_out ? 2 : 1;
      for (int g = 0; g < groups; ++g) {
        const uint32_t bytes = (uint32_t)(g == 0 ? st.n_out : st.n2_out) * 128u;
        const uint32_t off0 = g == 0 ? st.w_off : st.w2_off;
        for (int kc = 0; kc < st.n_chunks; ++kc, ++it) {
          const uint32_t stage = it % kStages, ph = (it / kStages) & 1u;
          if (lane == 0) {
            mbar_wait(sm.bar(BAR_WEMPTY + stage), ph ^ 1u);
            mbar_expect_tx(sm.bar(BAR_WFULL + stage), bytes);
            bulk_g2s(sm.stage_u32(stage), p.packed + off0 + (size_t)kc * bytes, bytes, sm.bar(BAR_WFULL + stage));
          }
          __syncwarp();
        }
      }
    }
  }
}

__device__ __forceinline__ void mma_loop(const Params& p, const Smem& sm, int64_t n_tiles, uint32_t lane, uint32_t tmem_base) {
  uint32_t it = 0, gstep = 0;
  uint32_t a_phase = 0;  // bit c = parity of the next completion of a_ready[c] to wait for
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int si = 0; si < p.prog.n_steps; ++si, ++gstep) {
      const Step& st = p.prog.s[si];
      const uint32_t half = gstep & 1u;
      const int groups = st.n2_out ? 2 : 1;
      for (int g = 0; g < groups; ++g) {
        const uint32_t n = g == 0 ? st.n_out : st.n2_out;
        const uint32_t d_tmem = tmem_base + (g == 0 ? half : (half ^ 1u)) * 256u;
        const uint32_t idesc = umma_idesc(128, n);
        for (int kc = 0; kc < st.n_chunks; ++kc, ++it) {
          const int ac = st.a_chunk[kc];
          const uint32_t stage = it % kStages, ph = (it / kStages) & 1u;
          if (lane == 0) {
            if (g == 0) mbar_wait(sm.bar(BAR_AREADY + ac), (a_phase >> ac) & 1u);
            mbar_wait(sm.bar(BAR_WFULL + stage), ph);
            tc_fence_after();
            const uint32_t a0 = sm.chunk_u32(ac), b0 = sm.stage_u32(stage);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(d_tmem, umma_desc(a0 + kk * 32), umma_desc(b0 + kk * 32), idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            umma_commit(sm.bar(BAR_WEMPTY + stage));
          }
          if (g == 0) a_phase ^= 1u << ac;
          __syncwarp();
        }
      }
      if (lane == 0) umma_commit(sm.bar(BAR_ACC + half));
      __syncwarp();
    }
  }
}

// after this thread's generic-proxy stores into an A chunk: publish to the async proxy and signal the MMA warp
__device__ __forceinline__ void publish_chunk(const Smem& sm, int c, uint32_t lane) {
  tc_fence_before();
  fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(sm.bar(BAR_AREADY + c));
}

__device__ __forceinline__ void store_row32(uint8_t* chunk, uint32_t row, uint32_t hh, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint4 q = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
    *reinterpret_cast<uint4*>(chunk + swz(row, hh * 4 + u)) = q;
  }
}

__device__ __forceinline__ void kernel_prologue(Smem& sm, uint8_t* smem_raw, uint32_t tid, uint32_t warp, uint32_t& tmem_base) {
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  sm.base = smem_raw + pad;
  sm.base_u32 = raw + pad;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm.base + SM_BARS + 32 * 8);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(sm.bar(BAR_WFULL + i), 1); mbar_init(sm.bar(BAR_WEMPTY + i), 1); }
    for (int i = 0; i < 5; ++i) mbar_init(sm.bar(BAR_AREADY + i), 8);
    for (int i = 0; i < 2; ++i) mbar_init(sm.bar(BAR_ACC + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_base = *tmem_slot;
}

// ------------------------------------------------------------------------------------------ forward epilogues
// One layer's epilogue for this thread's row: TMEM -> (+bias, ReLU, mask bits, +latent) -> bf16 -> shared-memory A operand.
// The TMEM load of chunk c+1 is in flight while chunk c is processed.
template <int EPI, bool LAT, bool DBG>
__device__ __forceinline__ void fwd_epilogue(const Params& p, const Smem& sm, const Step& st, int si, uint32_t half, const EpiCtx& e,
                                             uint32_t* mask_tile, float& sig_acc, float (&rgb_acc)[3]) {
  constexpr int NC = (EPI == EPI_RGB_HEAD) ? 2 : 4;
  // latent-conditioned layers read their per-object EFFECTIVE bias  b + W z_obj  (the latent add folded through the layer)
  const float* bias_s = LAT ? sm.tab(TAB_Z) + st.latent_slot * 256 : sm.tab(TAB_BIAS) + si * 256;
  const float* wsig_s = sm.tab(TAB_WSIG);
  const float* w2_s = sm.tab(TAB_W2);
  const uint32_t t0 = e.tmem_base + half * 256u + e.hh * 32u + e.lane_field;
  uint32_t ra[32], rb[32];
  tmem_ld32_issue(t0, ra);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    uint32_t (&r)[32] = (c & 1) ? rb : ra;
    tmem_ld_wait();
    if (c + 1 < NC) tmem_ld32_issue(t0 + (uint32_t)(c + 1) * 64u, (c & 1) ? ra : rb);
    const int col0 = c * 64 + (int)e.hh * 32;
    uint32_t pk[16], nmask = 0;   // nmask collects SIGN bits (1 = negative pre-activation); stored inverted
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 bb = *reinterpret_cast<const float4*>(bias_s + col0 + 4 * i4);
      float v[4] = {__uint_as_float(r[4 * i4]) + bb.x, __uint_as_float(r[4 * i4 + 1]) + bb.y,
                    __uint_as_float(r[4 * i4 + 2]) + bb.z, __uint_as_float(r[4 * i4 + 3]) + bb.w};
      if (EPI != EPI_LINEAR_SIGMA) {
#pragma unroll
        for (int u = 0; u < 4; ++u) nmask = __funnelshift_l(__float_as_uint(v[u]), nmask, 1);
      }
      if (EPI == EPI_LINEAR_SIGMA) {
        const float4 ws = *reinterpret_cast<const float4*>(wsig_s + col0 + 4 * i4);
        sig_acc += v[0] * ws.x + v[1] * ws.y + v[2] * ws.z + v[3] * ws.w;
      }
      if (EPI == EPI_RGB_HEAD) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 ww = *reinterpret_cast<const float4*>(w2_s + k * 128 + col0 + 4 * i4);
          rgb_acc[k] += fmaxf(v[0], 0.f) * ww.x + fmaxf(v[1], 0.f) * ww.y + fmaxf(v[2], 0.f) * ww.z + fmaxf(v[3], 0.f) * ww.w;
        }
      }
      if (DBG) {
        if (e.valid) {
          float* d = p.dbg + ((size_t)si * p.M + e.grow) * 256 + col0 + 4 * i4;
#pragma unroll
          for (int u = 0; u < 4; ++u) d[u] = (EPI == EPI_LINEAR_SIGMA) ? v[u] : fmaxf(v[u], 0.f);
        }
      }
      if (EPI == EPI_LINEAR_SIGMA) {
        pk[2 * i4] = pack_bf16(v[0], v[1]);
        pk[2 * i4 + 1] = pack_bf16(v[2], v[3]);
      } else if (EPI == EPI_RELU) {
        pk[2 * i4] = pack_bf16_relu(v[0], v[1]);
        pk[2 * i4 + 1] = pack_bf16_relu(v[2], v[3]);
      }
    }
    const uint32_t mask = ~nmask;
    if (EPI != EPI_LINEAR_SIGMA) mask_tile[((size_t)st.mask_slot * 8 + c * 2 + e.hh) * 128 + e.row] = mask;
    if (EPI != EPI_RGB_HEAD) {
      store_row32(sm.chunk(c), e.row, e.hh, pk);
      publish_chunk(sm, c, e.lane);
    }
  }
}

template <bool DBG>
__device__ __forceinline__ void fwd_epilogue_dispatch(const Params& p, const Smem& sm, const Step& st, int si, uint32_t half,
                                                      const EpiCtx& e, uint32_t* mask_tile, float& sig_acc, float (&rgb_acc)[3]) {
  if (st.epi == EPI_RELU) {
    if (st.latent_slot >= 0) fwd_epilogue<EPI_RELU, true, DBG>(p, sm, st, si, half, e, mask_tile, sig_acc, rgb_acc);
    else fwd_epilogue<EPI_RELU, false, DBG>(p, sm, st, si, half, e, mask_tile, sig_acc, rgb_acc);
  } else if (st.epi == EPI_LINEAR_SIGMA) {
    fwd_epilogue<EPI_LINEAR_SIGMA, false, DBG>(p, sm, st, si, half, e, mask_tile, sig_acc, rgb_acc);
  } else {
    fwd_epilogue<EPI_RGB_HEAD, false, DBG>(p, sm, st, si, half, e, mask_tile, sig_acc, rgb_acc);
  }
}

// ------------------------------------------------------------------------------------------ forward kernel
__global__ void __launch_bounds__(kThreads, 1) tc_fwd_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  Smem sm;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t tmem_base;
  {  // static tables: every layer's bias, the sigma / rgb head weights
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* b = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    float* bias_s = reinterpret_cast<float*>(b + SM_TAB + TAB_BIAS);
    for (int i = tid; i < p.prog.n_steps * 256; i += kThreads) {
      const int si = i >> 8, c = i & 255;
      bias_s[i] = c < p.prog.s[si].n_out ? __ldg(p.prog.s[si].bias + c) : 0.f;
    }
    float* wsig_s = reinterpret_cast<float*>(b + SM_TAB + TAB_WSIG);
    for (int i = tid; i < 256; i += kThreads) wsig_s[i] = __ldg(p.wsig + i);
    float* w2_s = reinterpret_cast<float*>(b + SM_TAB + TAB_W2);
    for (int i = tid; i < 384; i += kThreads) w2_s[i] = __ldg(p.w2 + i);
  }
  kernel_prologue(sm, smem_raw, tid, warp, tmem_base);
  const int64_t n_tiles = (p.M + kTileM - 1) / kTileM;

  if (warp == 8) {
    producer_loop(p, sm, n_tiles, lane);
  } else if (warp == 9) {
    mma_loop(p, sm, n_tiles, lane, tmem_base);
  } else {
    EpiCtx e;
    e.lane = lane; e.hh = warp >> 2; e.row = (warp & 3u) * 32u + lane; e.lane_field = ((warp & 3u) * 32u) << 16;
    e.tmem_base = tmem_base;
    float* sig_part = sm.tab(TAB_PART);          // [2][128]
    float* rgb_part = sm.tab(TAB_PART) + 256;    // [2][128][3]
    float* z_s = sm.tab(TAB_Z);
    uint32_t gstep = 0, acc_cnt[2] = {0, 0};
    const int nslots = p.prog.n_mask_slots;
    int64_t cur_obj = -1;
    // PE(xyz) of the first tile; later tiles are encoded one tile ahead (during the encoding_viewdir epilogue)
    if ((int64_t)blockIdx.x < n_tiles) {
      int64_t r0 = (int64_t)blockIdx.x * kTileM + e.row;
      if (r0 > p.M - 1) r0 = p.M - 1;
      const float x[3] = {__ldg(p.xyz + 3 * r0), __ldg(p.xyz + 3 * r0 + 1), __ldg(p.xyz + 3 * r0 + 2)};
      write_pe_row<10>(sm.chunk(4), e.row, e.hh, x);
      publish_chunk(sm, 4, lane);
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      e.grow = tile * kTileM + e.row;
      e.valid = e.grow < p.M;
      const int64_t crow = e.valid ? e.grow : p.M - 1;
      const int64_t next_tile = tile + gridDim.x;
      float xn[3] = {0.f, 0.f, 0.f};
      if (next_tile < n_tiles) {  // prefetch the next tile's coordinates: consumed ~5 layers from now
        int64_t rn = next_tile * kTileM + e.row;
        if (rn > p.M - 1) rn = p.M - 1;
        xn[0] = __ldg(p.xyz + 3 * rn); xn[1] = __ldg(p.xyz + 3 * rn + 1); xn[2] = __ldg(p.xyz + 3 * rn + 2);
      }
      const float dir[3] = {__ldg(p.viewdir + 3 * crow), __ldg(p.viewdir + 3 * crow + 1), __ldg(p.viewdir + 3 * crow + 2)};
      const int64_t obj = (tile * kTileM) / p.rows_per_obj;   // tiles never straddle objects (checked on the host)
      if (obj != cur_obj) {  // (re)load the per-object latent vectors; all epilogue warps are between tiles here
        cur_obj = obj;
        for (int i = tid; i < p.n_latent * 256; i += 256)
          z_s[i] = __ldg(p.zlat + ((size_t)(i >> 8) * p.B + obj) * 256 + (i & 255));
        epi_bar_sync();
      }
      uint32_t* mask_tile = p.masks + (size_t)tile * nslots * 8 * 128;
      float sig_acc = 0.f, rgb_acc[3] = {0.f, 0.f, 0.f};
      for (int si = 0; si < p.prog.n_steps; ++si, ++gstep) {
        const Step& st = p.prog.s[si];
        const uint32_t half = gstep & 1u;
        mbar_wait(sm.bar(BAR_ACC + half), acc_cnt[half] & 1u);
        acc_cnt[half]++;
        tc_fence_after();
        if (si == 0) {  // layer 0's MMAs are done with PE(xyz): the AUX chunk now takes PE(viewdir)
          write_pe_row<4>(sm.chunk(4), e.row, e.hh, dir);
          publish_chunk(sm, 4, lane);
        } else if (si == p.ev_step && next_tile < n_tiles) {  // encoding_viewdir's MMAs are done with PE(viewdir)
          write_pe_row<10>(sm.chunk(4), e.row, e.hh, xn);
          publish_chunk(sm, 4, lane);
        }
        if (p.dbg != nullptr) fwd_epilogue_dispatch<true>(p, sm, st, si, half, e, mask_tile, sig_acc, rgb_acc);
        else fwd_epilogue_dispatch<false>(p, sm, st, si, half, e, mask_tile, sig_acc, rgb_acc);
        tc_fence_before();
      }
      sig_part[e.hh * 128 + e.row] = sig_acc;
#pragma unroll
      for (int k = 0; k < 3; ++k) rgb_part[(e.hh * 128 + e.row) * 3 + k] = rgb_acc[k];
      epi_bar_sync();
      if (e.hh == 0 && e.valid) {
        const float sp = sig_part[e.row] + sig_part[128 + e.row] + __ldg(p.bsig);
        p.sigma[e.grow] = sp > 20.f ? sp : log1pf(expf(sp));   // nn.Softplus(): beta 1, threshold 20
#pragma unroll
        for (int k = 0; k < 3; ++k)
          p.rgb[3 * e.grow + k] = rgb_part[e.row * 3 + k] + rgb_part[(128 + e.row) * 3 + k] + __ldg(p.b2 + k);
      }
      epi_bar_sync();   // sig_part / rgb_part are rewritten at the end of the next tile only, but z_s may be reloaded next
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ backward epilogues
__device__ __forceinline__ uint32_t mask_word(const uint32_t* mask_tile, int slot, int word, uint32_t row) {
  return __ldg(mask_tile + ((size_t)slot * 8 + word) * 128 + row);
}

// acc = gradient w.r.t. the layer's input.  Optional: per-object column sums of it (latent gradient); then mask by the
// producing layer's ReLU bits and hand on as the next A operand.  EV: add the sigma-head gradient first, no mask.
template <bool EV, bool COLSUM, bool MASK, bool PRODUCE>
__device__ __forceinline__ void bwd_epilogue(const Smem& sm, const Step& st, uint32_t half, const EpiCtx& e,
                                             const uint32_t* mask_tile, float gsp) {
  const float* wsig_s = sm.tab(TAB_WSIG);
  float* colsum = sm.tab(TAB_BIAS);
  const uint32_t t0 = e.tmem_base + half * 256u + e.hh * 32u + e.lane_field;
  uint32_t ra[32], rb[32];
  tmem_ld32_issue(t0, ra);
  uint32_t mw_next = 0xffffffffu;
  if (MASK) mw_next = mask_word(mask_tile, st.mask_slot, (int)e.hh, e.row);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t (&r)[32] = (c & 1) ? rb : ra;
    tmem_ld_wait();
    if (c + 1 < 4) tmem_ld32_issue(t0 + (uint32_t)(c + 1) * 64u, (c & 1) ? ra : rb);
    const uint32_t mw = mw_next;
    if (MASK && c + 1 < 4) mw_next = mask_word(mask_tile, st.mask_slot, (c + 1) * 2 + (int)e.hh, e.row);
    const int col0 = c * 64 + (int)e.hh * 32;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (EV) {
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        const float4 ws = *reinterpret_cast<const float4*>(wsig_s + col0 + 4 * i4);
        v[4 * i4] += gsp * ws.x; v[4 * i4 + 1] += gsp * ws.y; v[4 * i4 + 2] += gsp * ws.z; v[4 * i4 + 3] += gsp * ws.w;
      }
    }
    if (PRODUCE) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = (!MASK || mask_bit(mw, 2 * i)) ? v[2 * i] : 0.f;
        const float b = (!MASK || mask_bit(mw, 2 * i + 1)) ? v[2 * i + 1] : 0.f;
        pk[i] = pack_bf16(a, b);
      }
      store_row32(sm.chunk(c), e.row, e.hh, pk);
      publish_chunk(sm, c, e.lane);   // the MMA warp can start the next layer on this chunk while we reduce below
    }
    if (COLSUM) {
      const float cs = warp_colsum32(v, e.lane);
      atomicAdd(colsum + st.latent_slot * 256 + col0 + e.lane, cs);
    }
  }
}

__device__ __forceinline__ void bwd_epilogue_dispatch(const Smem& sm, const Step& st, uint32_t half, const EpiCtx& e,
                                                      const uint32_t* mask_tile, float gsp) {
  if (st.epi == EPI_B_EV) {
    bwd_epilogue<true, false, false, true>(sm, st, half, e, mask_tile, gsp);
  } else if (st.latent_slot >= 0) {
    if (st.produce_a) bwd_epilogue<false, true, true, true>(sm, st, half, e, mask_tile, gsp);
    else bwd_epilogue<false, true, false, false>(sm, st, half, e, mask_tile, gsp);
  } else {
    bwd_epilogue<false, false, true, true>(sm, st, half, e, mask_tile, gsp);
  }
}

// ------------------------------------------------------------------------------------------ backward kernel
__global__ void __launch_bounds__(kThreads, 1) tc_bwd_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  Smem sm;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t tmem_base;
  {
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* b = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    float* colsum0 = reinterpret_cast<float*>(b + SM_TAB + TAB_BIAS);
    for (uint32_t i = tid; i < kMaxLatentSlots * 256; i += kThreads) colsum0[i] = 0.f;
    float* wsig_s = reinterpret_cast<float*>(b + SM_TAB + TAB_WSIG);
    for (int i = tid; i < 256; i += kThreads) wsig_s[i] = __ldg(p.wsig + i);
    float* w2_s = reinterpret_cast<float*>(b + SM_TAB + TAB_W2);
    for (int i = tid; i < 384; i += kThreads) w2_s[i] = __ldg(p.w2 + i);
  }
  kernel_prologue(sm, smem_raw, tid, warp, tmem_base);
  const int64_t n_tiles = (p.M + kTileM - 1) / kTileM;

  if (warp == 8) {
    producer_loop(p, sm, n_tiles, lane);
  } else if (warp == 9) {
    mma_loop(p, sm, n_tiles, lane, tmem_base);
  } else {
    EpiCtx e;
    e.lane = lane; e.hh = warp >> 2; e.row = (warp & 3u) * 32u + lane; e.lane_field = ((warp & 3u) * 32u) << 16;
    e.tmem_base = tmem_base;
    float* xyz_part = sm.tab(TAB_PART);          // [2][128][3]
    float* colsum = sm.tab(TAB_BIAS);            // [slots][256]
    const float* w2_s = sm.tab(TAB_W2);
    uint32_t gstep = 0, acc_cnt[2] = {0, 0};
    const int nslots = p.prog.n_mask_slots;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      e.grow = tile * kTileM + e.row;
      e.valid = e.grow < p.M;
      const int64_t crow = e.valid ? e.grow : p.M - 1;
      const int64_t obj = (tile * kTileM) / p.rows_per_obj;  // tiles never straddle objects (checked on the host)
      const uint32_t* mask_tile = p.masks + (size_t)tile * nslots * 8 * 128;
      const float gsg = e.valid ? __ldg(p.g_sigma + e.grow) : 0.f;
      const float gsp = gsg * (-expm1f(-__ldg(p.sigma_in + crow)));   // d softplus = 1 - exp(-softplus)
      // ---- prologue: d pre-activation of rgb.0 = (g_rgb W2) * mask -> A chunks 0,1 (128 columns)
      {
        float g3[3] = {0.f, 0.f, 0.f};
        if (e.valid) { g3[0] = __ldg(p.g_rgb + 3 * e.grow); g3[1] = __ldg(p.g_rgb + 3 * e.grow + 1); g3[2] = __ldg(p.g_rgb + 3 * e.grow + 2); }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = c * 64 + (int)e.hh * 32;
          const uint32_t mw = mask_word(mask_tile, p.r0_mask_slot, c * 2 + e.hh, e.row);
          uint32_t pk[16];
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 w0 = *reinterpret_cast<const float4*>(w2_s + col0 + 4 * i4);
            const float4 w1 = *reinterpret_cast<const float4*>(w2_s + 128 + col0 + 4 * i4);
            const float4 w2 = *reinterpret_cast<const float4*>(w2_s + 256 + col0 + 4 * i4);
            float v[4] = {g3[0] * w0.x + g3[1] * w1.x + g3[2] * w2.x, g3[0] * w0.y + g3[1] * w1.y + g3[2] * w2.y,
                          g3[0] * w0.z + g3[1] * w1.z + g3[2] * w2.z, g3[0] * w0.w + g3[1] * w1.w + g3[2] * w2.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) if (!mask_bit(mw, 4 * i4 + u)) v[u] = 0.f;
            pk[2 * i4] = pack_bf16(v[0], v[1]);
            pk[2 * i4 + 1] = pack_bf16(v[2], v[3]);
          }
          store_row32(sm.chunk(c), e.row, e.hh, pk);
          publish_chunk(sm, c, lane);
        }
      }
      for (int si = 0; si < p.prog.n_steps; ++si, ++gstep) {
        const Step& st = p.prog.s[si];
        const uint32_t half = gstep & 1u;
        mbar_wait(sm.bar(BAR_ACC + half), acc_cnt[half] & 1u);
        acc_cnt[half]++;
        tc_fence_after();
        if (st.epi == EPI_B_XYZ) {
          // acc = d PE(xyz) (64 columns): fold to d xyz.  g_x = g_0 + sum_f 2^f (g_sin,f cos_f - g_cos,f sin_f)
          uint32_t r[32];
          tmem_ld32(e.tmem_base + half * 256u + e.hh * 32u + e.lane_field, r);
          const float x[3] = {__ldg(p.xyz + 3 * crow), __ldg(p.xyz + 3 * crow + 1), __ldg(p.xyz + 3 * crow + 2)};
          float s[10][3], c[10][3];
          trig_ladder<10>(x, s, c);
          float g[3] = {0.f, 0.f, 0.f};
          if (e.hh == 0) {  // columns 0..31: x (0-2), sin f=0..8 (3-29), sin f=9 a=0,1 (30,31)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float gv = __uint_as_float(r[i]);
              if (i < 3) g[i] += gv;
              else { const int f = (i - 3) / 3, a = (i - 3) % 3; g[a] += gv * (float)(1 << f) * c[f][a]; }
            }
          } else {        // columns 32..63: sin f=9 a=2 (32), cos f=0..9 (33-62), pad (63)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float gv = __uint_as_float(r[i]);
              const int col = 32 + i;
              if (col == 32) g[2] += gv * 512.f * c[9][2];
              else if (col < 63) { const int f = (col - 33) / 3, a = (col - 33) % 3; g[a] -= gv * (float)(1 << f) * s[f][a]; }
            }
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) xyz_part[(e.hh * 128 + e.row) * 3 + a] = g[a];
          tc_fence_before();
          continue;
        }
        if (st.epi == EPI_B_EV && st.n2_out && e.hh == 0) {
          // group 2 accumulators (other TMEM half, columns 0..31) = d PE(viewdir): fold to d viewdir (deg 4: 27 columns)
          uint32_t r[32];
          tmem_ld32(e.tmem_base + (half ^ 1u) * 256u + e.lane_field, r);
          const float d[3] = {__ldg(p.viewdir + 3 * crow), __ldg(p.viewdir + 3 * crow + 1), __ldg(p.viewdir + 3 * crow + 2)};
          float s[4][3], c[4][3];
          trig_ladder<4>(d, s, c);
          float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 27; ++i) {
            const float gv = __uint_as_float(r[i]);
            if (i < 3) g[i] += gv;
            else if (i < 15) { const int f = (i - 3) / 3, a = (i - 3) % 3; g[a] += gv * (float)(1 << f) * c[f][a]; }
            else { const int f = (i - 15) / 3, a = (i - 15) % 3; g[a] -= gv * (float)(1 << f) * s[f][a]; }
          }
          if (e.valid && p.g_viewdir) { p.g_viewdir[3 * e.grow] = g[0]; p.g_viewdir[3 * e.grow + 1] = g[1]; p.g_viewdir[3 * e.grow + 2] = g[2]; }
        }
        bwd_epilogue_dispatch(sm, st, half, e, mask_tile, gsp);
        tc_fence_before();
      }
      // ---- tile end: d xyz, and flush the latent column sums when the next tile belongs to another object
      const int64_t next = tile + gridDim.x;
      const bool flush = next >= n_tiles || (next * kTileM) / p.rows_per_obj != obj;
      epi_bar_sync();
      if (p.g_xyz != nullptr && e.hh == 0 && e.valid) {
#pragma unroll
        for (int a = 0; a < 3; ++a) p.g_xyz[3 * e.grow + a] = xyz_part[e.row * 3 + a] + xyz_part[(128 + e.row) * 3 + a];
      }
      if (flush) {
        for (int sl = 0; sl < p.n_latent; ++sl) {
          const float v = colsum[sl * 256 + tid];
          if (v != 0.f) atomicAdd(p.g_zlat + ((size_t)sl * p.B + obj) * 256 + tid, v);
          colsum[sl * 256 + tid] = 0.f;
        }
      }
      epi_bar_sync();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ weight packing
struct PackJob {
  const float* src; int ld; int transposed; int n_valid; int n_pad; int k0; int k_limit; uint32_t dst_off;
};
constexpr int kJobsPerLaunch = 64;
struct PackJobs { int n; PackJob j[kJobsPerLaunch]; };

// one block per job: dst[n][k] (128B-swizzled rows of 64 bf16) = src'(n, k0 + k), zero outside the valid range
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackJobs jobs, uint8_t* __restrict__ packed) {
  const PackJob& jb = jobs.j[blockIdx.x];
  for (int e = threadIdx.x; e < jb.n_pad * 64; e += blockDim.x) {
    int n, k;
    if (jb.transposed) { k = e / jb.n_pad; n = e % jb.n_pad; }  // consecutive threads walk the contiguous source dimension
    else { n = e / 64; k = e % 64; }
    const int kg = jb.k0 + k;
    float v = 0.f;
    if (n < jb.n_valid && kg < jb.k_limit) v = jb.transposed ? jb.src[(size_t)kg * jb.ld + n] : jb.src[(size_t)n * jb.ld + kg];
    const uint32_t off = jb.dst_off + (uint32_t)n * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(packed + off) = __float2bfloat16_rn(v);
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------ host side
using namespace tc;

static bool tc_supported(const snb_handle_s* h, const char** why) {
  const snb_arch& a = h->arch;
  if (a.arch != SNB_ARCH_CODENERF) { *why = "bf16 mode covers the CodeNeRF/AutoRFMix/SUPNeRF decoder only"; return false; }
  if (a.W != 256) { *why = "bf16 mode needs W == 256"; return false; }
  if (a.num_xyz_freq != 10 || a.num_dir_freq != 4) { *why = "bf16 mode needs num_xyz_freq == 10 and num_dir_freq == 4"; return false; }
  if (a.shape_blocks + a.texture_blocks > kMaxLatentSlots) { *why = "bf16 mode needs shape_blocks + texture_blocks <= 8"; return false; }
  return true;
}

struct TcPlan {
  Program fwd, bwd_full, bwd_noxyz;
  std::vector<PackJob> jobs;
  uint32_t total_bytes = 0;
  int r0_slot = 0;
};

static void add_chunks(TcPlan& pl, const float* src, int ld, bool transposed, int n_valid, int n_pad, int k_limit, int n_chunks,
                       uint32_t* first_off) {
  *first_off = pl.total_bytes;
  for (int c = 0; c < n_chunks; ++c) {
    PackJob j;
    j.src = src; j.ld = ld; j.transposed = transposed ? 1 : 0; j.n_valid = n_valid; j.n_pad = n_pad; j.k0 = c * 64;
    j.k_limit = k_limit; j.dst_off = pl.total_bytes;
    pl.jobs.push_back(j);
    pl.total_bytes += (uint32_t)n_pad * 128u;
  }
}

static Step make_step(int epi, int n_out, int n_chunks, std::initializer_list<int> chunks, int mask_slot, int latent_slot,
                      int produce_a, const float* bias) {
  Step s{};
  s.epi = (int8_t)epi; s.n_out = (uint16_t)n_out; s.n_chunks = (uint16_t)n_chunks; s.mask_slot = (int8_t)mask_slot;
  s.latent_slot = (int8_t)latent_slot; s.produce_a = (int8_t)produce_a; s.bias = bias; s.n2_out = 0;
  int i = 0;
  for (int c : chunks) s.a_chunk[i++] = (uint8_t)c;
  return s;
}

// Builds the forward / backward step programs and the weight-image packing jobs for the handle's current pointers.
static TcPlan build_plan(const snb_handle_s* h) {
  TcPlan pl;
  const int Bs = h->arch.shape_blocks, Bt = h->arch.texture_blocks, W = 256, dv = h->d_dir(), dx = h->d_xyz();
  const auto& ly = h->layers;
  const int slot_vv = Bs + 1, slot_r = Bs + Bt + 2;
  pl.r0_slot = slot_r;
  Program& f = pl.fwd;
  f.n_steps = 0; f.n_mask_slots = Bs + Bt + 3;
  auto push = [](Program& pr, const Step& s) { pr.s[pr.n_steps++] = s; };
  {  // ---------------- forward
    Step s = make_step(EPI_RELU, 256, 1, {4}, 0, -1, 1, ly[h->iX].b);
    add_chunks(pl, ly[h->iX].w, dx, false, 256, 256, dx, 1, &s.w_off);
    push(f, s);
    for (int j = 1; j <= Bs; ++j) {
      s = make_step(EPI_RELU, 256, 4, {0, 1, 2, 3}, j, j - 1, 1, ly[h->iS(j)].b);   // effective bias slot j-1 = b + W zs_j
      add_chunks(pl, ly[h->iS(j)].w, W, false, 256, 256, W, 4, &s.w_off);
      push(f, s);
    }
    s = make_step(EPI_LINEAR_SIGMA, 256, 4, {0, 1, 2, 3}, -1, -1, 1, ly[h->iES].b);
    add_chunks(pl, ly[h->iES].w, W, false, 256, 256, W, 4, &s.w_off);
    push(f, s);
    s = make_step(EPI_RELU, 256, 5, {4, 0, 1, 2, 3}, slot_vv, -1, 1, ly[h->iEV].b);
    {  // chunk 0 = the PE(viewdir) columns [W, W+dv) of encoding_viewdir, chunks 1..4 = its first W columns
      add_chunks(pl, ly[h->iEV].w + W, W + dv, false, 256, 256, dv, 1, &s.w_off);
      uint32_t dummy;
      add_chunks(pl, ly[h->iEV].w, W + dv, false, 256, 256, W, 4, &dummy);
    }
    push(f, s);
    for (int j = 1; j <= Bt; ++j) {
      s = make_step(EPI_RELU, 256, 4, {0, 1, 2, 3}, slot_vv + j, Bs + j - 1, 1, ly[h->iT(j)].b);
      add_chunks(pl, ly[h->iT(j)].w, W, false, 256, 256, W, 4, &s.w_off);
      push(f, s);
    }
    s = make_step(EPI_RGB_HEAD, 128, 4, {0, 1, 2, 3}, slot_r, -1, 0, ly[h->iR0].b);
    add_chunks(pl, ly[h->iR0].w, W, false, 128, 128, W, 4, &s.w_off);
    push(f, s);
  }
  {  // ---------------- backward (B operand = W^T: n = input unit, k = output unit)
    Program& b = pl.bwd_full;
    b.n_steps = 0; b.n_mask_slots = f.n_mask_slots;
    Step s = make_step(EPI_B_MASK, 256, 2, {0, 1}, slot_vv + Bt, -1, 1, nullptr);          // through rgb.0 -> d T_Bt, mask of T_Bt
    add_chunks(pl, ly[h->iR0].w, W, true, 256, 256, 128, 2, &s.w_off);
    push(b, s);
    for (int j = Bt; j >= 1; --j) {                                                        // through texture_layer_j
      s = make_step(EPI_B_MASK, 256, 4, {0, 1, 2, 3}, slot_vv + j - 1, Bs + j - 1, 1, nullptr);
      add_chunks(pl, ly[h->iT(j)].w, W, true, 256, 256, W, 4, &s.w_off);
      push(b, s);
    }
    s = make_step(EPI_B_EV, 256, 4, {0, 1, 2, 3}, -1, -1, 1, nullptr);                     // through encoding_viewdir
    add_chunks(pl, ly[h->iEV].w, W + dv, true, 256, 256, W, 4, &s.w_off);
    add_chunks(pl, ly[h->iEV].w + W, W + dv, true, dv, 64, W, 4, &s.w2_off);
    s.n2_out = 64;
    push(b, s);
    s = make_step(EPI_B_MASK, 256, 4, {0, 1, 2, 3}, Bs, -1, 1, nullptr);                   // through encoding_shape, mask of H_Bs
    add_chunks(pl, ly[h->iES].w, W, true, 256, 256, W, 4, &s.w_off);
    push(b, s);
    for (int j = Bs; j >= 1; --j) {                                                        // through shape_layer_j
      s = make_step(EPI_B_MASK, 256, 4, {0, 1, 2, 3}, j - 1, j - 1, 1, nullptr);
      add_chunks(pl, ly[h->iS(j)].w, W, true, 256, 256, W, 4, &s.w_off);
      push(b, s);
    }
    s = make_step(EPI_B_XYZ, 64, 4, {0, 1, 2, 3}, -1, -1, 0, nullptr);                     // through encoding_xyz -> d PE(xyz)
    add_chunks(pl, ly[h->iX].w, dx, true, dx, 64, W, 4, &s.w_off);
    push(b, s);
    // variant without pose gradients: drop the last step and the d PE(viewdir) group; the new last step feeds nobody
    pl.bwd_noxyz = b;
    Program& n = pl.bwd_noxyz;
    n.n_steps = b.n_steps - 1;
    n.s[n.n_steps - 1].produce_a = 0;
    for (int i = 0; i < n.n_steps; ++i) if (n.s[i].epi == EPI_B_EV) n.s[i].n2_out = 0;
  }
  return pl;
}

// second-generation kernels (mlp_tc2.cu): two tiles in flight per CTA
bool tc2_supported(const snb_handle_s* h);
bool tc_two_tile_active(const snb_handle_s* h);
size_t tc2_packed_bytes(const snb_handle_s* h);
int tc2_pack_weights(const snb_handle_s* h, void* packed, cudaStream_t st);
int tc2_launch_fwd(const snb_handle_s* h, const void* packed2, const float* xyz, const float* viewdir, int64_t M, int64_t B,
                   const uint8_t* eimg, uint32_t* masks, float* sigma, float* rgb, float* dbg, uint8_t* save, cudaStream_t st,
                   const int64_t* m_dev, const int32_t* tile_start, const rb::RowSrc* rs);
int tc2_launch_bwd(const snb_handle_s* h, const void* packed2, const float* xyz, const float* viewdir, int64_t M, int64_t B,
                   const uint32_t* masks, const float* sigma, const float* g_sigma, const float* g_rgb, float* g_xyz,
                   float* g_viewdir, float* g_zlat, uint8_t* save, cudaStream_t st, const int64_t* m_dev, const int32_t* tile_start);
size_t tc2_fwd_save_bytes(const snb_handle_s* h, int64_t M);
size_t tc2_bwd_save_bytes(const snb_handle_s* h, int64_t M);
int tc2_launch_wgrad(const snb_handle_s* h, int64_t M, int64_t B, const uint8_t* fsave, const uint8_t* bsave, const float* sigma,
                     const float* g_sigma, const float* g_rgb, const float* s_lat, const float* zlat, float* const* gw,
                     cudaStream_t st, const int64_t* m_dev);
int tc2_launch_latent_wgrad(const snb_handle_s* h, int64_t B, const float* zlat, const float* dz, const float* shape_latent,
                            const float* texture_latent, float* const* gw, cudaStream_t st);

static size_t v1_packed_bytes(const snb_handle_s* h) {
  if (h->v1_packed_bytes_cache == 0) h->v1_packed_bytes_cache = ((size_t)build_plan(h).total_bytes + 1024 + 1023) & ~size_t(1023);
  return h->v1_packed_bytes_cache;
}
static bool use_v2(const snb_handle_s* h) {
  static const bool force_v1 = [] { const char* e = getenv("SNB_TC_V1"); return e && atoi(e) != 0; }();
  return !force_v1 && tc2_supported(h);
}

bool tc_two_tile_active(const snb_handle_s* h) { const char* why; return tc_supported(h, &why) && use_v2(h); }

size_t tc_packed_bytes(const snb_handle_s* h) {
  const char* why;
  if (!tc_supported(h, &why)) return 16;
  return v1_packed_bytes(h) + tc2_packed_bytes(h);
}

int tc_pack_weights(snb_handle_s* h, void* packed, cudaStream_t st) {
  const char* why = "";
  SNB_REQUIRE(tc_supported(h, &why), "snb_pack_weights: %s", why);
  SNB_REQUIRE(((uintptr_t)packed & 15) == 0, "snb_pack_weights: buffer must be 16-byte aligned");
  TcPlan pl = build_plan(h);
  for (size_t i = 0; i < pl.jobs.size(); i += kJobsPerLaunch) {
    PackJobs jb;
    jb.n = (int)std::min<size_t>(kJobsPerLaunch, pl.jobs.size() - i);
    for (int k = 0; k < jb.n; ++k) jb.j[k] = pl.jobs[i + k];
    pack_kernel<<<jb.n, 256, 0, st>>>(jb, (uint8_t*)packed);
    SNB_LAUNCH_CHECK();
  }
  if (tc2_supported(h) && tc2_pack_weights(h, (uint8_t*)packed + v1_packed_bytes(h), st)) return 1;
  h->packed = packed;
  return 0;
}

static inline int64_t tiles_of(int64_t M) { return (M + kTileM - 1) / kTileM; }

// workspace: [zlat (Bs+Bt)*B*256 fp32][masks tiles*slots*8*128 u32]
static size_t ws_zlat_bytes(const snb_handle_s* h, int64_t B) {
  return (size_t)(h->arch.shape_blocks + h->arch.texture_blocks) * B * 256 * sizeof(float);
}
// workspace: [zlat][effective biases fp32][effective-bias stage images, 32 B per unit][masks]
static size_t ws_eimg_off(const snb_handle_s* h, int64_t B) { return (2 * ws_zlat_bytes(h, B) + 255) & ~size_t(255); }
static size_t ws_masks_off(const snb_handle_s* h, int64_t B) { return (ws_eimg_off(h, B) + 8 * ws_zlat_bytes(h, B) + 255) & ~size_t(255); }
size_t tc_workspace_bytes(const snb_handle_s* h, int64_t M, int64_t B) {
  const int slots = h->arch.shape_blocks + h->arch.texture_blocks + 3;
  return ws_masks_off(h, B) + (size_t)tiles_of(M) * slots * 8 * 128 * 4 + 256;
}
size_t tc_bwd_scratch_bytes(const snb_handle_s* h, int64_t, int64_t B) { return ((2 * ws_zlat_bytes(h, B) + 255) & ~size_t(255)) + 256; }   // column sums + their W^T fold
// training mode (SNB_PREC_BF16_TRAIN): the operand tiles kept for the weight-gradient kernels live behind the normal
// workspace (forward) / scratch (backward: + a d xyz / d viewdir dummy, since training runs the full backward program)
size_t tc_train_workspace_extra(const snb_handle_s* h, int64_t M) { return tc2_supported(h) ? tc2_fwd_save_bytes(h, M) + 1024 : 0; }
size_t tc_train_scratch_extra(const snb_handle_s* h, int64_t M) {
  return tc2_supported(h) ? tc2_bwd_save_bytes(h, M) + (size_t)M * 24 + 1024 : 0;
}
static inline uint8_t* align1k(uint8_t* p) { return (uint8_t*)(((uintptr_t)p + 1023) & ~uintptr_t(1023)); }

static int tc_common_checks(const snb_handle_s* h, int64_t M, int64_t B, const char* who, bool ragged = false) {
  const char* why = "";
  SNB_REQUIRE(tc_supported(h, &why), "%s: %s", who, why);
  SNB_REQUIRE(h->packed != nullptr, "%s: weights not packed (call snb_pack_weights)", who);
  SNB_REQUIRE(ragged || (M / B) % kTileM == 0, "%s: bf16 mode needs samples-per-object (%lld) to be a multiple of %d; use fp32 mode",
              who, (long long)(M / B), kTileM);
  return 0;
}

static void fill_common(Params& p, const snb_handle_s* h, const float* xyz, const float* viewdir, int64_t M, int64_t B,
                        const float* zlat, uint32_t* masks) {
  p = Params{};
  p.xyz = xyz; p.viewdir = viewdir; p.M = M; p.B = B; p.rows_per_obj = M / B;
  p.packed = (const uint8_t*)h->packed; p.zlat = zlat; p.masks = masks;
  p.wsig = h->layers[h->iSG].w; p.bsig = h->layers[h->iSG].b;
  p.n_latent = h->arch.shape_blocks + h->arch.texture_blocks;
  p.ev_step = h->arch.shape_blocks + 2;  // forward step index of encoding_viewdir
  p.w2 = h->layers[h->iR2].w; p.b2 = h->layers[h->iR2].b;
}

static int tc_grid(int64_t M) {
  const int sms = sm_count();
  const int64_t t = tiles_of(M);
  return (int)(t < sms ? t : sms);
}

// optional kernel-only timing (bench.py roofline): CUDA events recorded on the launch stream around the tcgen05 kernels
void tc_timing_enable(snb_handle_s* h, int on) {
  h->timing_on = on != 0;
  for (auto* v : {&h->ev_fwd, &h->ev_bwd}) {
    for (auto& e : *v) { cudaEventDestroy((cudaEvent_t)e.first); cudaEventDestroy((cudaEvent_t)e.second); }
    v->clear();
  }
}
int tc_timing_read(snb_handle_s* h, int which, float* ms, int max_n) {   // call after a stream/device synchronize
  auto& v = which == 0 ? h->ev_fwd : h->ev_bwd;
  int n = 0;
  for (auto& e : v) {
    if (n >= max_n) break;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, (cudaEvent_t)e.first, (cudaEvent_t)e.second) == cudaSuccess) ms[n++] = t;
  }
  return n;
}
struct ScopedKernelTimer {
  cudaStream_t st; std::pair<cudaEvent_t, cudaEvent_t> ev; bool on;
  ScopedKernelTimer(cudaStream_t s, bool enable) : st(s), on(enable) {
    if (on) { cudaEventCreate(&ev.first); cudaEventCreate(&ev.second); cudaEventRecord(ev.first, st); }
  }
  void stop(std::vector<std::pair<void*, void*>>& into) { if (on) { cudaEventRecord(ev.second, st); into.push_back({(void*)ev.first, (void*)ev.second}); } }
};


int tc_forward(const snb_handle_s* h, const float* xyz, const float* viewdir, int64_t M, int64_t B,
               const float* shape_latent, const float* texture_latent, float* sigma, float* rgb, void* ws, cudaStream_t st,
               bool train, const int64_t* m_dev, const int32_t* tile_start, const rb::RowSrc* rs) {
  if (tc_common_checks(h, M, B, "mlp_fwd(bf16)", tile_start != nullptr)) return 2;
  SNB_REQUIRE(rs == nullptr || (tile_start != nullptr && rs->rays8 && rs->box && rs->z_steps && rs->jitter && rs->order && rs->counts),
              "mlp_fwd(bf16): a row source needs per-object tile offsets and all of its pointers");
  SNB_REQUIRE(rs != nullptr || (xyz != nullptr && viewdir != nullptr), "mlp_fwd(bf16): null coordinates");
  SNB_REQUIRE(m_dev == nullptr || (use_v2(h) && (B == 1 || tile_start != nullptr)),
              "mlp_fwd(bf16): a device-side row count needs the two-tile kernels and one object (or per-object tile offsets)");
  SNB_REQUIRE(tile_start == nullptr || (use_v2(h) && m_dev != nullptr && !train), "mlp_fwd(bf16): per-object tile offsets need the two-tile kernels, frozen weights and a device-side row count");
  SNB_REQUIRE(!train || use_v2(h), "mlp_fwd(bf16, training): weight gradients need the two-tile tcgen05 kernels (W = 256, "
                                    "shape_blocks + texture_blocks <= 4); use precision='fp32' for this architecture");
  uint8_t* fsave = train ? align1k((uint8_t*)ws + tc_workspace_bytes(h, M, B)) : nullptr;
  float* zlat = (float*)ws;
  float* ebias = (float*)((uint8_t*)ws + ws_zlat_bytes(h, B));
  uint32_t* masks = (uint32_t*)((uint8_t*)ws + ws_masks_off(h, B));
  uint8_t* eimg = (uint8_t*)ws + ws_eimg_off(h, B);
  if (latent_forward_fused(h, B, shape_latent, texture_latent, zlat, ebias, st, use_v2(h) ? eimg : nullptr)) return 1;
  if (use_v2(h)) {
    ScopedKernelTimer tm2(st, h->timing_on);
    if (tc2_launch_fwd(h, (const uint8_t*)h->packed + v1_packed_bytes(h), xyz, viewdir, M, B, eimg, masks, sigma, rgb,
                       h->dbg_acts, fsave, st, m_dev, tile_start, rs)) return 1;
    tm2.stop(const_cast<snb_handle_s*>(h)->ev_fwd);
    SNB_LAUNCH_CHECK();
    return 0;
  }
  TcPlan pl = build_plan(h);
  Params p;
  fill_common(p, h, xyz, viewdir, M, B, ebias, masks);
  p.sigma = sigma; p.rgb = rgb; p.dbg = h->dbg_acts;
  p.prog = pl.fwd;
  SNB_CHECK_CUDA(cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC));
  ScopedKernelTimer tm(st, h->timing_on);
  tc_fwd_kernel<<<tc_grid(M), kThreads, SM_ALLOC, st>>>(p);
  tm.stop(const_cast<snb_handle_s*>(h)->ev_fwd);
  SNB_LAUNCH_CHECK();
  return 0;
}

int tc_backward(const snb_handle_s* h, const float* xyz, const float* viewdir, int64_t M, int64_t B,
                const float* shape_latent, const float* texture_latent, const float* sigma, const float* g_sigma,
                const float* g_rgb, const void* ws, void* scratch, float* g_xyz, float* g_viewdir, float* g_shape_latent,
                float* g_texture_latent, float* const* g_weights, cudaStream_t st, bool train, const int64_t* m_dev,
                const int32_t* tile_start) {
  if (tc_common_checks(h, M, B, "mlp_bwd(bf16)", tile_start != nullptr)) return 2;
  SNB_REQUIRE(m_dev == nullptr || (use_v2(h) && (B == 1 || tile_start != nullptr)),
              "mlp_bwd(bf16): a device-side row count needs the two-tile kernels and one object (or per-object tile offsets)");
  SNB_REQUIRE(tile_start == nullptr || (use_v2(h) && m_dev != nullptr && !train), "mlp_bwd(bf16): per-object tile offsets need the two-tile kernels, frozen weights and a device-side row count");
  SNB_REQUIRE(g_weights == nullptr || (train && use_v2(h)),
              "mlp_bwd(bf16): weight gradients need the forward to have run in training mode (SNB_PREC_BF16_TRAIN: the python "
              "modules select it when a weight requires grad) on an architecture the two-tile kernels cover; otherwise freeze "
              "the weights (requires_grad_(False)) or use precision='fp32'");
  SNB_REQUIRE((g_xyz == nullptr) == (g_viewdir == nullptr), "mlp_bwd(bf16): request both g_xyz and g_viewdir or neither");
  const float* zlat = (const float*)ws;
  uint32_t* masks = (uint32_t*)((uint8_t*)ws + ws_masks_off(h, B));
  float* g_zlat = (float*)scratch;
  SNB_CHECK_CUDA(cudaMemsetAsync(g_zlat, 0, ws_zlat_bytes(h, B), st));
  if (use_v2(h)) {
    float* fold_tmp = (float*)((uint8_t*)scratch + ws_zlat_bytes(h, B));
    uint8_t* bsave = nullptr;
    const bool want_w = train && g_weights != nullptr;
    if (want_w) {   // training: keep every step's d pre-activation tile, run the full program (d xyz into a dummy if unwanted)
      bsave = align1k((uint8_t*)scratch + tc_bwd_scratch_bytes(h, M, B));
      float* dummy = (float*)(bsave + tc2_bwd_save_bytes(h, M));
      if (g_xyz == nullptr) { g_xyz = dummy; g_viewdir = dummy + 3 * M; }
      for (size_t i = 0; i < h->layers.size(); ++i) {   // every weight gradient is accumulated (atomics / +=): zero first
        SNB_CHECK_CUDA(cudaMemsetAsync(g_weights[2 * i], 0, sizeof(float) * h->layers[i].out * h->layers[i].in, st));
        SNB_CHECK_CUDA(cudaMemsetAsync(g_weights[2 * i + 1], 0, sizeof(float) * h->layers[i].out, st));
      }
    }
    ScopedKernelTimer tm2(st, h->timing_on);
    if (tc2_launch_bwd(h, (const uint8_t*)h->packed + v1_packed_bytes(h), xyz, viewdir, M, B, masks, sigma, g_sigma, g_rgb, g_xyz,
                       g_viewdir, g_zlat, bsave, st, m_dev, tile_start)) return 1;
    tm2.stop(const_cast<snb_handle_s*>(h)->ev_bwd);
    SNB_LAUNCH_CHECK();
    if (!want_w) return latent_backward_fused(h, B, zlat, g_zlat, g_shape_latent, g_texture_latent, st, fold_tmp);
    const uint8_t* fsave = align1k((uint8_t*)ws + tc_workspace_bytes(h, M, B));
    if (tc2_launch_wgrad(h, M, B, fsave, bsave, sigma, g_sigma, g_rgb, g_zlat, zlat, g_weights, st, m_dev)) return 1;
    // latent layers (per object): fold the column sums through W_layer^T (= d loss / d z), d latent, then their weight gradients
    if (latent_backward_fused(h, B, zlat, g_zlat, g_shape_latent, g_texture_latent, st, fold_tmp)) return 1;
    return tc2_launch_latent_wgrad(h, B, zlat, fold_tmp, shape_latent, texture_latent, g_weights, st);
  }
  TcPlan pl = build_plan(h);
  Params p;
  fill_common(p, h, xyz, viewdir, M, B, zlat, masks);
  p.sigma_in = sigma; p.g_sigma = g_sigma; p.g_rgb = g_rgb; p.g_xyz = g_xyz; p.g_viewdir = g_viewdir; p.g_zlat = g_zlat;
  p.r0_mask_slot = pl.r0_slot;
  p.prog = g_xyz ? pl.bwd_full : pl.bwd_noxyz;
  SNB_CHECK_CUDA(cudaFuncSetAttribute(tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_ALLOC));
  ScopedKernelTimer tm(st, h->timing_on);
  tc_bwd_kernel<<<tc_grid(M), kThreads, SM_ALLOC, st>>>(p);
  tm.stop(const_cast<snb_handle_s*>(h)->ev_bwd);
  SNB_LAUNCH_CHECK();
  return latent_backward_fused(h, B, zlat, g_zlat, g_shape_latent, g_texture_latent, st);
}

}  // namespace snb
