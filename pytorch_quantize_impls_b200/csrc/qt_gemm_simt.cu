// CUDA-core contractions: XNOR+popcount (1-bit), signed popcount (ternary), dp4a (8-bit codes),
// fp32 / bf16-plane FMA.  One register-tiled kernel template over 32-bit "K words":
//   128x128 output tile per 256-thread CTA, 8x8 accumulators per thread, K staged through shared
//   memory in 8-word slabs stored K-major-transposed ([k][row]) so that a thread's 8 row operands and
//   8 column operands are two LDS.128 each (broadcast across the 16 threads that share them).
// These are the any-shape routes (ragged K, tiny M, first layers); the large aligned shapes go to the
// tcgen05 kernels in qt_gemm_tc.cu.
#include <type_traits>
#include <cuda_fp16.h>
#include "qt_common.cuh"

namespace qt {

constexpr int BM = 128, BN = 128, KT = 8, TM = 8, TN = 8, NTHREADS = 256;

struct OpXnor {      // acc = sum popc(a ^ w);  result = K - 2 acc
  using Acc = int;
  static constexpr bool kHasAux = false;
  __device__ static __forceinline__ void mac(Acc& acc, uint32_t a, uint32_t w, uint32_t) { acc += __popc(a ^ w); }
};
struct OpTernary {   // acc = sum popc(nz & (a ^ s));  result = popc(nz) - 2 acc
  using Acc = int;
  static constexpr bool kHasAux = true;
  __device__ static __forceinline__ void mac(Acc& acc, uint32_t a, uint32_t s, uint32_t nz) { acc += __popc(nz & (a ^ s)); }
};
template <bool AS, bool WS>
struct OpDp4a {
  using Acc = int;
  static constexpr bool kHasAux = false;
  __device__ static __forceinline__ void mac(Acc& acc, uint32_t a, uint32_t w, uint32_t) {
    if (AS && WS) asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(w));
    else if (AS && !WS) asm("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(w));
    else if (!AS && WS) asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(w));
    else asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(w));
  }
};
struct OpF32 {
  using Acc = float;
  static constexpr bool kHasAux = false;
  __device__ static __forceinline__ void mac(Acc& acc, uint32_t a, uint32_t w, uint32_t) {
    acc = fmaf(__uint_as_float(a), __uint_as_float(w), acc);
  }
};
struct OpBf16x2 {
  using Acc = float;
  static constexpr bool kHasAux = false;
  __device__ static __forceinline__ void mac(Acc& acc, uint32_t a, uint32_t w, uint32_t) {
    acc = fmaf(__uint_as_float(a << 16), __uint_as_float(w << 16), acc);
    acc = fmaf(__uint_as_float(a & 0xffff0000u), __uint_as_float(w & 0xffff0000u), acc);
  }
};

struct OpFp16x2 {
  using Acc = float;
  static constexpr bool kHasAux = false;
  __device__ static __forceinline__ void mac(Acc& acc, uint32_t a, uint32_t w, uint32_t) {
    float2 fa = __half22float2(*reinterpret_cast<__half2*>(&a)), fw = __half22float2(*reinterpret_cast<__half2*>(&w));
    acc = fmaf(fa.x, fw.x, acc);
    acc = fmaf(fa.y, fw.y, acc);
  }
};

struct SimtArgs {
  const uint32_t* a;     // [M, lda] words
  const uint32_t* w;     // [N, ldw] words
  const uint32_t* aux;   // ternary nz plane [N, ldw]
  int64_t lda, ldw, M, N;
  int64_t kwords;        // words per row actually contracted
  int64_t K;             // logical K (popcount kernels)
  int npass;             // bf16 planes: passes
  int64_t a_plane, w_plane;  // plane strides in words
  int pa[8], pw[8];
  int finish;            // 0: plain, 1: K - 2 acc, 2: nzcount - 2 acc
  Epi ep;
};

template <typename Op>
__global__ void __launch_bounds__(NTHREADS) simt_gemm_kernel(SimtArgs g) {
  using Acc = typename Op::Acc;
  __shared__ __align__(16) uint32_t As[KT][BM + 4];
  __shared__ __align__(16) uint32_t Ws[KT][BN + 4];
  __shared__ __align__(16) uint32_t Xs[Op::kHasAux ? KT : 1][Op::kHasAux ? BN + 4 : 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

  Acc acc[TM][TN];
  int nzc[TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = Acc(0);
#pragma unroll
  for (int j = 0; j < TN; ++j) nzc[j] = 0;

  // global -> smem mapping: thread loads 4 consecutive K words of one row (one 16 B load when aligned)
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const bool a_vec = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.a) & 15) == 0);
  const bool w_vec = (g.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.w) & 15) == 0) &&
                     (!Op::kHasAux || (reinterpret_cast<uintptr_t>(g.aux) & 15) == 0);

  for (int pass = 0; pass < g.npass; ++pass) {
    const uint32_t* A = g.a + (int64_t)g.pa[pass] * g.a_plane;
    const uint32_t* W = g.w + (int64_t)g.pw[pass] * g.w_plane;
    for (int64_t k0 = 0; k0 < g.kwords; k0 += KT) {
      uint32_t av[4] = {0, 0, 0, 0}, wv[4] = {0, 0, 0, 0}, xv[4] = {0, 0, 0, 0};
      const int64_t kk = k0 + lk;
      {
        const int64_t r = m0 + lrow;
        if (r < g.M) {
          const uint32_t* p = A + r * g.lda + kk;
          if (a_vec && kk + 4 <= g.kwords) {
            uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
            av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (kk + j < g.kwords) av[j] = __ldg(p + j);
          }
        }
      }
      {
        const int64_t r = n0 + lrow;
        if (r < g.N) {
          const uint32_t* p = W + r * g.ldw + kk;
          if (w_vec && kk + 4 <= g.kwords) {
            uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
            wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
            if (Op::kHasAux) {
              uint4 u = __ldg(reinterpret_cast<const uint4*>(g.aux + r * g.ldw + kk));
              xv[0] = u.x; xv[1] = u.y; xv[2] = u.z; xv[3] = u.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (kk + j < g.kwords) {
              wv[j] = __ldg(p + j);
              if (Op::kHasAux) xv[j] = __ldg(g.aux + r * g.ldw + kk + j);
            }
          }
        }
      }
      __syncthreads();   // previous slab fully consumed
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[lk + j][lrow] = av[j];
        Ws[lk + j][lrow] = wv[j];
        if (Op::kHasAux) Xs[lk + j][lrow] = xv[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        uint32_t ar[TM], wr[TN], xr[TN];
        *reinterpret_cast<uint4*>(&ar[0]) = *reinterpret_cast<const uint4*>(&As[k][ty * TM]);
        *reinterpret_cast<uint4*>(&ar[4]) = *reinterpret_cast<const uint4*>(&As[k][ty * TM + 4]);
        *reinterpret_cast<uint4*>(&wr[0]) = *reinterpret_cast<const uint4*>(&Ws[k][tx * TN]);
        *reinterpret_cast<uint4*>(&wr[4]) = *reinterpret_cast<const uint4*>(&Ws[k][tx * TN + 4]);
        if (Op::kHasAux) {
          *reinterpret_cast<uint4*>(&xr[0]) = *reinterpret_cast<const uint4*>(&Xs[k][tx * TN]);
          *reinterpret_cast<uint4*>(&xr[4]) = *reinterpret_cast<const uint4*>(&Xs[k][tx * TN + 4]);
#pragma unroll
          for (int j = 0; j < TN; ++j) nzc[j] += __popc(xr[j]);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) Op::mac(acc[i][j], ar[i], wr[j], Op::kHasAux ? xr[j] : 0u);
      }
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      if constexpr (sizeof(Acc) == 4 && std::is_same<Acc, int>::value) {
        int v = acc[i][j];
        if (g.finish == 1) v = (int)g.K - 2 * v;
        else if (g.finish == 2) v = nzc[j] - 2 * v;
        if (g.ep.acc_out) g.ep.acc_out[m * g.N + n] = v;
        if (g.ep.out) g.ep.out[epi_addr(g.ep, m, n)] = epi_int(g.ep, m, n, v);
      } else {
        if (g.ep.out) g.ep.out[epi_addr(g.ep, m, n)] = epi_f32(g.ep, m, n, acc[i][j]);
      }
    }
  }
}

template <typename Op>
static int launch_simt(SimtArgs& g, cudaStream_t stream) {
  if (g.M == 0 || g.N == 0) return QT_OK;
  if (epi_needs_tc(g.ep)) {
    set_error("fused requant / partial-sum row operands are implemented by the tcgen05 kernels only");
    return QT_EUNSUPPORTED;
  }
  dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM));
  QT_REQUIRE(grid.y <= 65535, "simt gemm: M too large for grid.y");
  simt_gemm_kernel<Op><<<grid, NTHREADS, 0, stream>>>(g);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

int simt_gemm_i8(const void* a, int a_signed, int64_t lda, const void* w, int w_signed, int64_t ldw,
                 int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, cudaStream_t stream) {
  QT_REQUIRE(lda % 4 == 0 && ldw % 4 == 0, "qt_gemm_i8: lda/ldw must be multiples of 4 bytes");
  QT_REQUIRE(lda >= ((K + 3) / 4) * 4 && ldw >= ((K + 3) / 4) * 4, "qt_gemm_i8: rows must be padded (zero-filled) to 4 bytes");
  SimtArgs g{};
  g.a = (const uint32_t*)a; g.w = (const uint32_t*)w; g.aux = nullptr;
  g.lda = lda / 4; g.ldw = ldw / 4; g.M = M; g.N = N; g.kwords = (K + 3) / 4; g.K = K;
  g.npass = 1; g.finish = 0; g.ep = make_epi(ep, M, N);
  if (a_signed && w_signed) return launch_simt<OpDp4a<true, true>>(g, stream);
  if (a_signed && !w_signed) return launch_simt<OpDp4a<true, false>>(g, stream);
  if (!a_signed && w_signed) return launch_simt<OpDp4a<false, true>>(g, stream);
  return launch_simt<OpDp4a<false, false>>(g, stream);
}

int simt_gemm_f16(const void* a, int64_t lda, int64_t a_plane_stride, const void* w, int64_t ldw,
                  int64_t w_plane_stride, int fmt, int npass, const int* pa, const int* pw,
                  int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, cudaStream_t stream) {
  QT_REQUIRE(lda % 2 == 0 && ldw % 2 == 0 && a_plane_stride % 2 == 0 && w_plane_stride % 2 == 0,
             "qt_gemm_f16: strides must be even");
  QT_REQUIRE(lda >= ((K + 1) / 2) * 2 && ldw >= ((K + 1) / 2) * 2, "qt_gemm_f16: rows must be zero-padded to 2 elements");
  SimtArgs g{};
  g.a = (const uint32_t*)a; g.w = (const uint32_t*)w;
  g.lda = lda / 2; g.ldw = ldw / 2; g.M = M; g.N = N; g.kwords = (K + 1) / 2; g.K = K;
  g.npass = npass; g.a_plane = a_plane_stride / 2; g.w_plane = w_plane_stride / 2;
  for (int i = 0; i < npass; ++i) { g.pa[i] = pa[i]; g.pw[i] = pw[i]; }
  g.finish = 0; g.ep = make_epi(ep, M, N);
  if (fmt == 1) return launch_simt<OpFp16x2>(g, stream);
  return launch_simt<OpBf16x2>(g, stream);
}

}  // namespace qt

using namespace qt;

extern "C" int qt_gemm_b1b1(const uint32_t* a_bits, int64_t lda_words, const uint32_t* w_bits, int64_t ldw_words,
                            int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, void* stream) {
  QT_REQUIRE(a_bits && w_bits, "qt_gemm_b1b1: null operand");
  QT_REQUIRE(M >= 0 && N >= 0 && K > 0, "qt_gemm_b1b1: bad shape");
  QT_REQUIRE(lda_words * 32 >= K && ldw_words * 32 >= K, "qt_gemm_b1b1: ld*32 < K");
  if (int rc = check_epi(ep, M, N)) return rc;
  SimtArgs g{};
  g.a = a_bits; g.w = w_bits; g.lda = lda_words; g.ldw = ldw_words; g.M = M; g.N = N;
  g.kwords = (K + 31) / 32; g.K = K; g.npass = 1; g.finish = 1; g.ep = make_epi(ep, M, N);
  return launch_simt<OpXnor>(g, (cudaStream_t)stream);
}

extern "C" int qt_gemm_b1t2(const uint32_t* a_bits, int64_t lda_words, const uint32_t* w_nz, const uint32_t* w_sign,
                            int64_t ldw_words, int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, void* stream) {
  QT_REQUIRE(a_bits && w_nz && w_sign, "qt_gemm_b1t2: null operand");
  QT_REQUIRE(M >= 0 && N >= 0 && K > 0, "qt_gemm_b1t2: bad shape");
  QT_REQUIRE(lda_words * 32 >= K && ldw_words * 32 >= K, "qt_gemm_b1t2: ld*32 < K");
  if (int rc = check_epi(ep, M, N)) return rc;
  SimtArgs g{};
  g.a = a_bits; g.w = w_sign; g.aux = w_nz; g.lda = lda_words; g.ldw = ldw_words; g.M = M; g.N = N;
  g.kwords = (K + 31) / 32; g.K = K; g.npass = 1; g.finish = 2; g.ep = make_epi(ep, M, N);
  return launch_simt<OpTernary>(g, (cudaStream_t)stream);
}

extern "C" int qt_gemm_f32(const float* a, int64_t lda, const float* w, int64_t ldw,
                           int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, void* stream) {
  QT_REQUIRE(a && w, "qt_gemm_f32: null operand");
  QT_REQUIRE(M >= 0 && N >= 0 && K > 0 && lda >= K && ldw >= K, "qt_gemm_f32: bad shape");
  if (int rc = check_epi(ep, M, N)) return rc;
  QT_REQUIRE(ep->out, "qt_gemm_f32: needs ep->out");
  SimtArgs g{};
  g.a = (const uint32_t*)a; g.w = (const uint32_t*)w; g.lda = lda; g.ldw = ldw; g.M = M; g.N = N;
  g.kwords = K; g.K = K; g.npass = 1; g.finish = 0; g.ep = make_epi(ep, M, N);
  return launch_simt<OpF32>(g, (cudaStream_t)stream);
}
