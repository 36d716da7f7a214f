// Micro-benchmark: serial latency of a tcgen05.mma batch on B200: issue `nb` MMAs (M128 N256 K16, bf16, SS), commit,
// wait for the mbarrier, repeat.  mode 0: the issuing thread waits itself.  mode 1: a second warp waits for the commit,
// does the fences an epilogue would (tcgen05.fence, fence.proxy.async) and arrives on a second mbarrier the issuer waits on
// (the hand-shake of a fused MLP layer chain).  Prints cycles per round trip and the overhead over nb x 128 cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared mma_latency.cu -o mma_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) { uint64_t d = (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)(1024u >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d; }
__device__ __forceinline__ uint32_t idesc(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t ph) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(ph) : "memory");
}
__global__ void __launch_bounds__(64, 1) k(int nb, int iters, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint32_t base = smem_u32(raw); uint32_t pad = (1024 - (base & 1023)) & 1023; base += pad;
  __shared__ uint64_t bar[2]; __shared__ uint32_t tslot;
  uint32_t* w = (uint32_t*)(raw + pad);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) w[i] = 0x3c003c00u + (i & 7);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tslot;
  const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
  if (threadIdx.x == 0) {
    uint32_t id = idesc(128, 256);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < nb; ++j) {
        uint64_t da = desc(base + (j & 3) * 32), db = desc(base + 16384 + (j & 3) * 32);
        uint32_t acc = j ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm + (it & 1) * 256), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0) : "memory");
      if (mode == 0) { mwait(b0, it & 1); asm volatile("tcgen05.fence::after_thread_sync;"); }
      else { mwait(b1, it & 1); asm volatile("tcgen05.fence::after_thread_sync;"); }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (threadIdx.x == 32 && mode == 1) {
    for (int it = 0; it < iters; ++it) {
      mwait(b0, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      asm volatile("tcgen05.fence::before_thread_sync;");
      asm volatile("fence.proxy.async.shared::cta;");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  size_t smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int iters = 512;
  for (int mode = 0; mode < 2; ++mode)
    for (int nb : {1, 4, 16, 32}) {
      for (int grid : {1, 148}) {
        cudaMemset(out, 0, 148 * 8);
        k<<<grid, 64, smem>>>(nb, iters, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < grid; ++i) c += h[i]; c /= grid;
        printf("mode %d nb %2d grid %3d: %.0f cyc per round trip, overhead over nb*128 = %.0f cyc\n", mode, nb, grid, c / iters, c / iters - nb * 128.0);
      }
    }
  return 0;
}
