// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16, SS mode, 128B swizzle) on B200 for a few shapes.
// Every CTA (one per SM) issues `iters` x 4 MMAs (K=16 each) over operands resident in shared memory, commits, waits.
// Prints cycles per MMA instruction and the implied MAC/cycle/SM.   nvcc -arch=sm_100a -O3 -cudart shared mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) { uint64_t d = (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)(1024u >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d; }
__device__ __forceinline__ uint32_t idesc(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }
template <int CG>
__global__ void __launch_bounds__(128, 1) k(int M, int N, int iters, int distinct_b, long long* out) {
  extern __shared__ uint8_t raw[];
  uint32_t base = smem_u32(raw); uint32_t pad = (1024 - (base & 1023)) & 1023; base += pad;
  __shared__ uint64_t bar; __shared__ uint32_t tslot;
  uint32_t* w = (uint32_t*)(raw + pad);
  for (int i = threadIdx.x; i < (16384 + 4 * 32768) / 4; i += blockDim.x) w[i] = 0x3c003c00u + (i & 7);  // finite bf16 values
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    if (CG == 1) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    else { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;"); } else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    uint32_t id = idesc(M, N);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t a0 = base, b0 = base + 16384 + (distinct_b ? (it & 3) * 32768 : 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = desc(a0 + kk * 32), db = desc(b0 + kk * 32);
        uint32_t acc = (it | kk) ? 1u : 0u;
        if (CG == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
        else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
      }
    }
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;"); } else __syncthreads();
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
  }
}
template <int CG> void run(int M, int N, int distinct_b, const char* tag) {
  int iters = 2048, grid = 148;
  long long* out; cudaMalloc(&out, grid * 8); cudaMemset(out, 0, grid * 8);
  size_t smem = 16384 + 4 * 32768 + 1024;
  cudaFuncSetAttribute(k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)CG, 1, 1}; cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, k<CG>, M, N, iters, distinct_b, out);
    cudaEventRecord(e1); cudaError_t e2 = cudaDeviceSynchronize();
    if (err != cudaSuccess || e2 != cudaSuccess) { printf("%s: launch error %s / %s\n", tag, cudaGetErrorString(err), cudaGetErrorString(e2)); return; }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  double cyc = 0; int n = 0; for (int i = 0; i < grid; ++i) if (h[i] > 0) { cyc += h[i]; ++n; }
  cyc /= n; double per = cyc / (iters * 4.0);
  double macs_per_sm = (double)M * N * 16 / CG;   // per SM per instruction
  double tflops = 2.0 * M * N * 16 * iters * 4.0 * n / (ms * 1e-3) / 1e12;
  printf("%-34s cg%d M%3d N%3d: %.1f cyc/MMA  %.0f MAC/cyc/SM  kernel %.3f ms -> %.0f TFLOP/s chip\n", tag, CG, M, N, per, macs_per_sm / per, ms, tflops);
}
int main() {
  run<1>(128, 256, 0, "cg1 same B"); run<1>(128, 256, 1, "cg1 4 distinct B tiles");
  run<1>(128, 128, 0, "cg1 N128"); run<1>(128, 64, 0, "cg1 N64"); run<1>(64, 256, 0, "cg1 M64");
  run<2>(256, 256, 0, "cg2 M256 (128/CTA) N256"); run<2>(256, 128, 0, "cg2 M256 N128");
  return 0;
}
