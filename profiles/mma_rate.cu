// Microbenchmark: issue rate of tcgen05.mma for kind::i8 / f8f6f4 / f16 (bf16) / tf32, M=128, N=256, cta_group::1,
// operands in smem (content irrelevant), one CTA per SM.  Prints ns per MMA and TMAC/s per chip.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)(1024u >> 4) << 32; d |= 1ull << 46; d |= 2ull << 61; return d;
}
template <int KIND> __device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 1) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 2) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 3) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND, int BN>
__global__ void __launch_bounds__(128, 1) rate_kernel(uint32_t idesc, int iters, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot; __shared__ __align__(8) uint64_t bar;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + BN * 128) / 4; i += blockDim.x) ((volatile uint32_t*)(raw + (base - smem_u32(raw))))[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    uint64_t a = make_desc(base), b = make_desc(base + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma<KIND>(tmem + (i & 1) * BN, a + 2 * k, b + 2 * k, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int KIND, int BN> void run(const char* name, uint32_t idesc, int kelems, int nblocks) {
  long long* dc; cudaMalloc(&dc, 8);
  size_t smem = 16384 + BN * 128 + 1024;
  cudaFuncSetAttribute(rate_kernel<KIND, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<KIND, BN><<<nblocks, 128, smem>>>(idesc, 1000, dc); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  rate_kernel<KIND, BN><<<nblocks, 128, smem>>>(idesc, iters, dc);
  cudaEventRecord(e1); cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  double n_mma = (double)iters * 4;
  double macs = n_mma * 128.0 * BN * kelems * nblocks;
  printf("%-10s N=%3d blocks=%3d: %s  %.1f ns/MMA  %.1f cycles/MMA  %.1f TMAC/s chip (%.2f Pop/s)\n", name, BN, nblocks,
         cudaGetErrorString(err), ms * 1e6 / n_mma, (double)cyc / n_mma, macs / (ms * 1e-3) / 1e12, 2 * macs / (ms * 1e-3) / 1e15);
  cudaFree(dc);
}
int main() {
  int sms = 148;
  auto idesc = [](int cfmt, int afmt, int bfmt, int n) { return (uint32_t)((cfmt << 4) | (afmt << 7) | (bfmt << 10) | ((n >> 3) << 17) | ((128 >> 4) << 24)); };
  for (int nb : {1, sms}) {
    run<0, 256>("i8", idesc(2, 1, 1, 256), 32, nb);
    run<1, 256>("fp8 e4m3", idesc(1, 0, 0, 256), 32, nb);
    run<2, 256>("bf16", idesc(1, 1, 1, 256), 16, nb);
    run<3, 256>("tf32", idesc(1, 2, 2, 256), 8, nb);
    run<0, 128>("i8", idesc(2, 1, 1, 128), 32, nb);
    run<1, 128>("fp8 e4m3", idesc(1, 0, 0, 128), 32, nb);
  }
  return 0;
}
