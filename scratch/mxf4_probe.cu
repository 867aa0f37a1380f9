// Probe: tcgen05.mma kind::mxf4.block_scale with unit (ue8m0 = 0x7F) scale factors written to TMEM by tcgen05.st.
//  (1) exactness: A[r,:] = va(r), B[n,:] = vb(n) constant rows of e2m1 nibbles -> D[r,n] must equal K * va(r) * vb(n)
//  (2) issue rate at N = 240 / 256 (zero operands), 1 and 148 CTAs, next to kind::i8.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)(1024u >> 4) << 32; d |= 1ull << 46; d |= 2ull << 61; return d;
}
__device__ __forceinline__ void mma_mxf4(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MODE 0: mxf4 exactness (writes D[128,BN] floats); MODE 1: mxf4 rate; MODE 2: i8 rate
template <int MODE, int BN>
__global__ void __launch_bounds__(128, 1) probe_kernel(uint32_t idesc, int iters, long long* cycles, float* dout, const uint8_t* codes_a,
                                                      const uint8_t* codes_b, int sf_col) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __shared__ uint32_t slot; __shared__ __align__(8) uint64_t bar;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A tile: 128 rows x 128 B, B tile: BN rows x 128 B (one 128B-swizzle K block = 256 fp4 = 4 MMAs of K=64)
  for (int i = threadIdx.x; i < (16384 + BN * 128) / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (MODE == 0) {
      int byte = i * 4, row = (byte < 16384) ? byte / 128 : (byte - 16384) / 128;
      uint8_t c = (byte < 16384) ? codes_a[row] : codes_b[row];
      v = 0x01010101u * (uint32_t)(uint8_t)(c | (c << 4));
    }
    ((volatile uint32_t*)gen)[i] = v;
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem = slot;
  // unit scale factors: 16 columns x 128 lanes of 0x7F bytes (ue8m0 2^0)
  tmem_st16(tmem + (uint32_t)sf_col + ((uint32_t)(warp * 32) << 16), 0x7F7F7F7Fu);
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t sfa = tmem + sf_col, sfb = tmem + sf_col + 4;
  if (threadIdx.x == 0) {
    uint64_t a = make_desc(base), b = make_desc(base + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == 2) mma_i8(tmem + (MODE ? (i & 1) * BN : 0), a + 2 * k, b + 2 * k, idesc, 1);
        else mma_mxf4(tmem + (MODE ? (i & 1) * BN : 0), a + 2 * k, b + 2 * k, idesc, (MODE == 0 && i == 0 && k == 0) ? 0u : 1u, sfa, sfb);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  if (MODE == 0) {
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + c0 + ((uint32_t)(warp * 32) << 16), r);
      for (int j = 0; j < 16; ++j) dout[(warp * 32 + lane) * BN + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float e2m1(uint8_t c) {
  static const float mag[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
  float v = mag[c & 7];
  return (c & 8) ? -v : v;
}
static uint32_t idesc_mxf4(int n, int sfa_id = 0, int sfb_id = 0) {
  return (uint32_t)((sfb_id << 4) | (1 << 7) | (1 << 10) | ((n >> 3) << 17) | (1 << 23) | ((128 >> 4) << 24) | (sfa_id << 29));
}
template <int BN> void exact(int iters, int sf_col) {
  uint8_t ha[128], hb[256];
  const uint8_t pal[6] = {2, 10, 0, 2, 10, 4};   // +1, -1, 0, +1, -1, +2
  for (int i = 0; i < 128; ++i) ha[i] = pal[(i * 7 + i / 32) % 6];
  for (int i = 0; i < 256; ++i) hb[i] = pal[(i * 5 + 1 + i / 64) % 5];
  uint8_t *da, *db; float* dd; long long* dc;
  cudaMalloc(&da, 128); cudaMalloc(&db, 256); cudaMalloc(&dd, 128 * BN * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(da, ha, 128, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, 256, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xFF, 128 * BN * 4);
  size_t smem = 16384 + BN * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel<0, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<0, BN><<<1, 128, smem>>>(idesc_mxf4(BN), iters, dc, dd, da, db, sf_col);
  cudaError_t err = cudaDeviceSynchronize();
  float* hd = (float*)malloc(128 * BN * 4);
  cudaMemcpy(hd, dd, 128 * BN * 4, cudaMemcpyDeviceToHost);
  long bad = 0; float K = 256.f * iters;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < BN; ++n) {
      float ref = K * e2m1(ha[r]) * e2m1(hb[n]);
      if (hd[r * BN + n] != ref) { if (bad < 5) printf("   mismatch r=%d n=%d got %g want %g\n", r, n, hd[r * BN + n], ref); ++bad; }
    }
  printf("exact mxf4 N=%d K=%d sf_col=%d: %s, %ld mismatches of %d\n", BN, (int)K, sf_col, cudaGetErrorString(err), bad, 128 * BN);
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc); free(hd);
}
template <int MODE, int BN> void rate(const char* name, uint32_t idesc, int kelems, int nblocks) {
  long long* dc; cudaMalloc(&dc, 8);
  size_t smem = 16384 + BN * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe_kernel<MODE, BN><<<nblocks, 128, smem>>>(idesc, 1000, dc, nullptr, nullptr, nullptr, 496); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe_kernel<MODE, BN><<<nblocks, 128, smem>>>(idesc, iters, dc, nullptr, nullptr, nullptr, 496);
  cudaEventRecord(e1); cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  double n_mma = (double)iters * 4;
  double macs = n_mma * 128.0 * BN * kelems * nblocks;
  printf("%-6s N=%3d blocks=%3d: %s  %.1f ns/MMA  %.1f cycles/MMA  %.1f TMAC/s chip (%.2f Pop/s)\n", name, BN, nblocks,
         cudaGetErrorString(err), ms * 1e6 / n_mma, (double)cyc / n_mma, macs / (ms * 1e-3) / 1e12, 2 * macs / (ms * 1e-3) / 1e15);
  cudaFree(dc);
}
int main() {
  exact<240>(1, 480); exact<240>(16, 480); exact<256>(16, 256); exact<128>(3, 496); exact<64>(2, 128);
  for (int nb : {1, 148}) {
    rate<1, 240>("mxf4", idesc_mxf4(240), 64, nb);
    rate<1, 128>("mxf4", idesc_mxf4(128), 64, nb);
    rate<2, 240>("i8", (2u << 4) | (1u << 7) | (1u << 10) | ((240u >> 3) << 17) | ((128u >> 4) << 24), 32, nb);
  }
  return 0;
}
