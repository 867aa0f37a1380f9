// tcgen05 contraction core for sm_100a:  D[M,N] = sum_passes A_p[M,K] . W_p[N,K]^T  with a fused epilogue.
//
//   * operands are K-major byte matrices in HBM/L2 (int8/uint8 codes, or bf16 planes), fetched by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage shared-memory ring guarded by
//     full/empty mbarriers;
//   * one elected thread issues tcgen05.mma (kind::i8 -> s32, kind::f16/bf16 -> f32, kind::mxf4 -> f32), 128 x BN x 32 B
//     per instruction, accumulators live in TMEM (2 x BN columns: double buffered across output tiles);
//   * kind::mxf4.block_scale (e2m1 operands, 64 K-elements per 32 B: twice the MAC rate of kind::i8) is run with UNIT
//     scale factors: the ue8m0 value 0x7F (2^0) is written once into 16 spare TMEM columns with tcgen05.st and every MMA
//     points its SFA / SFB operands there, so +-1 / {-1,0,1} / 2-bit codes are multiplied exactly and the fp32
//     accumulators hold exact integers (tile width 240 instead of 256 leaves room for those columns);
//   * eight epilogue warps drain TMEM with tcgen05.ld (32x32b.x32), apply the integer-exact affine +
//     scale/bias epilogue and write fp32 through 128B-swizzled smem tiles + cp.async.bulk.tensor stores
//     (row-major output) or coalesced per-column stores (NCHW output), overlapping the next tile's main loop;
//   * persistent: grid = min(#tiles, #SMs), static round-robin tile schedule with M fastest so that
//     concurrently running CTAs share the same weight tile in L2.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "qt_common.cuh"

namespace qt {

constexpr int TC_BM = 128;
constexpr int TC_BK_BYTES = 128;          // one 128B swizzle row of K per stage
constexpr int TC_THREADS = 320;           // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..9 epilogue
constexpr uint32_t TC_SF_COL = 496;       // kind::mxf4: TMEM columns 496..511 hold the unit scale factors

struct TcArgs {
  int64_t M, N;
  int num_kblocks;      // per pass
  int npass;
  int pa[8], pw[8];
  int64_t a_plane_rows, w_plane_rows;
  uint32_t idesc;
  int is_int;
  Epi ep;
  int tiles_m, tiles_n;
  int tma_store;        // 1: row-major fp32 output goes through swizzled smem + cp.async.bulk.tensor store
  int epi_fast;         // 1: the epilogue is the plain "fp32 tile -> TMA store" form (tight code path, see epilogue_f32_tma)
  int epi_debug;        // profiling aid (QTB200_EPI_DEBUG): 1 = release the accumulator without reading it, 2 = read TMEM, store nothing
  // implicit-GEMM conv (IM2COL kernels): A rows are output pixels (b, oh, ow), K runs over (kh, kw, c-blocks)
  int cv_OW, cv_OHW;            // output width, output pixels per image
  int cv_sh, cv_sw, cv_ph, cv_pw, cv_dh, cv_dw, cv_kw;
  int cv_cblocks;               // channel blocks (of BKB bytes) per filter tap
  int cv_c0;                    // first channel of the conv group
  // division by run-time constants as multiply-high + shift (a 32-bit integer division is ~40 dependent instructions; the
  // single TMA-issuing lane executed two per k-block and four per tile, which bounded the conv kernels -- see DESIGN.md 3.2)
  uint32_t fd_band[2], fd_ohw[2], fd_ow[2];
};

// n / d for 0 <= n < 2^31: q = umulhi(n, mul) >> shr, mul == 0 encodes d == 1
static inline void fastdiv_make(uint32_t d, uint32_t out[2]) {
  if (d <= 1) { out[0] = 0; out[1] = 0; return; }
  uint32_t lg = 0;
  while ((1ull << lg) < d) ++lg;
  const uint32_t p = 31 + lg;
  out[0] = (uint32_t)(((1ull << p) + d - 1) / d);
  out[1] = p - 32;
}
__device__ __forceinline__ int fastdiv(int n, const uint32_t fd[2]) {
  return fd[0] ? (int)(__umulhi((uint32_t)n, fd[0]) >> fd[1]) : n;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One lane of a converged warp (elect.sync): the TMA / MMA issuing warps run their loops with all 32 lanes so that stage
// indices, shared-memory addresses and descriptors stay warp-uniform (uniform registers feed UTMALDG / UTCxMMA directly);
// under `if (lane == 0)` the same values count as divergent and every issue costs an ELECT + R2UR broadcast loop
// (~15 instructions per MMA: the single issuing thread, not the tensor pipe, bounded the kernels -- DESIGN.md 3.2).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA im2col load: 128 output pixels x BKB channel bytes of filter tap (off_w, off_h), starting at base pixel (w, h) of
// image n (coordinates in the padded "bounding box" space); out-of-image taps are zero-filled by the hardware.
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// cta_group::2 variants: the destination is this CTA's shared memory, the mbarrier may be the peer (leader) CTA's.
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the variable at shared::cta address `addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// accumulator hand-off to the pair leader: the TMEM reads were completed by tcgen05.wait::ld and ordered by
// tcgen05.fence::before_thread_sync, no global / shared data rides on this arrive
// .relaxed: a releasing arrive is a MEMBAR, which waits for every global store the warp still has in flight (the requant codes
// go out as plain stores) -- per tile, on the critical path of the accumulator ring.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit of cta_group::2 MMAs: arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_cg2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Tile rasterisation: M is cut into bands of TC_BAND_M tiles; inside a band the schedule is M-fastest, so the CTAs that
// run concurrently share a handful of weight tiles AND a band of activation rows small enough to stay L2-resident
// while the fp32 output streams through (without bands the A operand is re-fetched from HBM once per ~2 weight tiles).
constexpr int TC_BAND_M = 16;
template <int CG = 1>
__device__ __forceinline__ void tile_coords(const TcArgs& g, int tile, int& tm, int& tn) {
  constexpr int BAND = TC_BAND_M / CG;                            // the band is 2048 rows either way
  const int band_tiles = BAND * g.tiles_n;
  const int band = fastdiv(tile, g.fd_band);
  const int t = tile - band * band_tiles;
  const int bm = min(BAND, g.tiles_m - band * BAND);             // last band may be shorter
  tn = (bm == BAND) ? t / BAND : t / bm;                         // BAND is a power of two
  tm = band * BAND + (t - tn * bm);
}

template <int KIND, int CG = 1>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                       uint32_t sf_tmem) {
  if (CG == 2) {
    // one instruction drives the tensor cores of both SMs of the pair: M = 256 (128 rows per CTA), each CTA's shared memory
    // holds its own A rows and half of the B rows at the same offsets
    if (KIND == 2) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sf_tmem), "r"(sf_tmem + 4u)
          : "memory");
    } else if (KIND == 0) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
    }
    return;
  }
  if (KIND == 2) {
    // SFA at sf_tmem (4 columns), SFB at sf_tmem + 4 (<= 8 columns for N <= 256): all bytes there are 0x7F (ue8m0 1.0)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sf_tmem), "r"(sf_tmem + 4u)
        : "memory");
  } else if (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// K-major, 128B-swizzled operand tile: 8-row atoms of 1024 B (SBO), LBO unused, descriptor version 1 (sm_100).
template <int BKB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  // rows of BKB bytes, BKB-byte swizzle (128 / 64 / 32), 8-row atoms -> SBO = 8 * BKB
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((8u * BKB) >> 4) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)(BKB == 128 ? 2 : (BKB == 64 ? 4 : 6)) << 61;   // SWIZZLE_128B / 64B / 32B
  return d;
}

// tcgen05.ld of one 32-column chunk of this warp's 32 TMEM lanes (thread = lane = accumulator row) + the wait for it, as ONE
// asm statement: the destination registers are written asynchronously until tcgen05.wait::ld retires, and nothing the
// compiler could schedule between two separate statements (a register move, a spill) may touch them.  (A split issue / wait
// that keeps the next chunk's read in flight during the stores was measured: the TMEM read is not what the epilogue waits
// for -- DESIGN.md 3.2 -- so the simple form costs nothing.)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 per-column epilogue parameters (col_scale / bias) of columns n0..n0+31 into registers.  Every lane reads the SAME
// addresses (a lane owns one output row and all 32 columns of the chunk): 8 uniform 16-byte loads, one L1 wavefront each.
__device__ __forceinline__ void load_col32(const float* __restrict__ p, int n0, int n_lim, bool vec, float fill, float (&v)[32]) {
  if (vec && n0 + 32 <= n_lim) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + n0) + q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (n0 + j < n_lim) ? __ldg(p + n0 + j) : fill;
  }
}

// ---- fused re-quantisation of one 32-column chunk of an output row (QtRequant) ------------------------------------
// Semantics are those of act_quant_kernel (qt_quantize.cu): safeSign / ternary thresholds / rint(n y) / torch.sign.
template <int MODE>
__device__ __forceinline__ float rq_quant_t(const Epi& e, float v) {
  if (MODE == QT_Q_SIGN) return (v < 0.f) ? -1.f : 1.f;
  if (MODE == QT_Q_TERNARY) {
    const float s = (v < 0.f) ? -1.f : 1.f;
    const float t = v - 0.5f * s;
    return (s + ((t < 0.f) ? -1.f : 1.f)) * 0.5f;
  }
  if (MODE == QT_Q_DOREFA) return rintf(e.rq_n * v);
  return (float)((v > 0.f) - (v < 0.f));   // QT_Q_XNOR_ROW
}
__device__ __forceinline__ float rq_quant(const Epi& e, float v) {
  switch (e.rq_mode) {
    case QT_Q_SIGN: return (v < 0.f) ? -1.f : 1.f;
    case QT_Q_TERNARY: {
      const float s = (v < 0.f) ? -1.f : 1.f;
      const float t = v - 0.5f * s;
      return (s + ((t < 0.f) ? -1.f : 1.f)) * 0.5f;
    }
    case QT_Q_DOREFA: return rintf(e.rq_n * v);
    default: return (float)((v > 0.f) - (v < 0.f));   // QT_Q_XNOR_ROW
  }
}
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_global_v2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// y[0..31]: fp32 outputs of row m, columns n0..n0+31 (valid while n < n_lim).  Writes the codes of columns n0..n0+ncols-1
// (zeros past n_lim), accumulates the row partial sums.  Works through the chunk 8 columns at a time so that only the packed
// words, not 32 more floats and 32 more integers, are live beside the caller's registers.
template <int MODE, int KIND>
__device__ __forceinline__ void rq_store_chunk_t(const Epi& e, const float (&y)[32], int64_t m, int n0, int n_lim, int ncols,
                                                 float& psum, int& isum, bool& ovf) {
  constexpr int kind = KIND;
  const bool f16 = (kind == 3 || kind == 5);
  const float lo = (kind == 7) ? -4.f : ((kind == 1) ? -128.f : 0.f), hi = (kind == 7) ? 4.f : ((kind == 1) ? 127.f : 255.f);
  uint8_t* const base = reinterpret_cast<uint8_t*>(e.rq_codes);
  uint32_t w[4];      // words waiting for their 16-byte (int8) / 8-byte (e2m1) store
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = y[8 * q + j];
      if (e.rq_clamp) v = fminf(fmaxf(v, e.rq_lo), e.rq_hi);
      const bool ok = n0 + 8 * q + j < n_lim;
      const float qv = rq_quant_t<MODE>(e, v);
      if (MODE == QT_Q_XNOR_ROW && ok) psum += v;
      c[j] = ok ? qv : 0.f;
    }
    if (f16) {                           // bf16 / fp16 lanes: 16 bytes per 8 columns
      uint32_t h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kind == 5) {
          __half2 t = __floats2half2_rn(c[2 * j], c[2 * j + 1]);
          h[j] = *reinterpret_cast<uint32_t*>(&t);
        } else {
          __nv_bfloat162 t = __floats2bfloat162_rn(c[2 * j], c[2 * j + 1]);
          h[j] = *reinterpret_cast<uint32_t*>(&t);
        }
      }
      if (8 * q < ncols && n0 + 8 * q < e.rq_cover) st_global_v4(base + (m * e.rq_ld + n0) * 2 + 16 * q, h[0], h[1], h[2], h[3]);
      continue;
    }
    int k[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = c[j];
      if (!(v >= lo && v <= hi)) { ovf = true; v = (v != v) ? 0.f : fminf(fmaxf(v, lo), hi); }
      k[j] = (int)v;
      isum += k[j];
    }
    if (kind == 7) {                     // e2m1 nibbles: 4 bytes per 8 columns, element 2j in the low nibble of byte j
      uint32_t acc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int mag = k[j] < 0 ? -k[j] : k[j];
        acc |= (((0x65420u >> (4 * mag)) & 0xFu) | (k[j] < 0 ? 8u : 0u)) << (4 * j);
      }
      w[q & 1] = acc;
      if (q & 1) {
        uint8_t* dst = base + ((m * e.rq_ld + n0) >> 1) + 4 * (q - 1);
        if ((q == 1 && n0 < e.rq_cover) || (q == 3 && 16 < ncols && n0 + 16 < e.rq_cover)) st_global_v2(dst, w[0], w[1]);
      }
    } else {                             // int8 / uint8 lanes: 8 bytes per 8 columns
      w[2 * (q & 1)] = (uint32_t)(k[0] & 0xff) | ((uint32_t)(k[1] & 0xff) << 8) | ((uint32_t)(k[2] & 0xff) << 16) |
                       ((uint32_t)(k[3] & 0xff) << 24);
      w[2 * (q & 1) + 1] = (uint32_t)(k[4] & 0xff) | ((uint32_t)(k[5] & 0xff) << 8) | ((uint32_t)(k[6] & 0xff) << 16) |
                           ((uint32_t)(k[7] & 0xff) << 24);
      if (q & 1) {
        uint8_t* dst = base + (m * e.rq_ld + n0) + 8 * (q - 1);
        if ((q == 1 && n0 < e.rq_cover) || (q == 3 && 16 < ncols && n0 + 16 < e.rq_cover)) st_global_v4(dst, w[0], w[1], w[2], w[3]);
      }
    }
  }
}

// The quantizer and the lane format are compile-time constants inside the chunk code (no per-element selects); the switch
// costs one warp-uniform branch per 32 x 32 chunk.
__device__ __forceinline__ void rq_store_chunk(const Epi& e, const float (&y)[32], int64_t m, int n0, int n_lim, int ncols,
                                               float& psum, int& isum, bool& ovf) {
#define QT_RQ(M_, K_) case (M_) * 8 + (K_): rq_store_chunk_t<M_, K_>(e, y, m, n0, n_lim, ncols, psum, isum, ovf); break;
  switch (e.rq_mode * 8 + e.rq_codes_kind) {
    QT_RQ(QT_Q_SIGN, 1) QT_RQ(QT_Q_SIGN, 7) QT_RQ(QT_Q_TERNARY, 1) QT_RQ(QT_Q_TERNARY, 7)
    QT_RQ(QT_Q_DOREFA, 1) QT_RQ(QT_Q_DOREFA, 2) QT_RQ(QT_Q_DOREFA, 7)
    QT_RQ(QT_Q_XNOR_ROW, 3) QT_RQ(QT_Q_XNOR_ROW, 5)
    default: break;
  }
#undef QT_RQ
}

// ---------------------------------------------------------------------------------------------
// Epilogue, tight form: fp32 row-major output through swizzled staging tiles + bulk tensor stores, optional row / column
// scales, bias and clamp -- no requant, residual, raw-accumulator or NCHW output (those take the general path below).
// A product with a short K loop (e2m1 operands: 4x the MACs per byte of fp16) is paced by this code, not by the MMAs: the
// column parameters are requested before the accumulator chunk is read, the fp32 tile goes out through one or two staging
// tiles per warp, and nothing but the 32 accumulators and 32 parameters is live.
// ---------------------------------------------------------------------------------------------
template <int BN, int KIND, int CG, int EPB>
__device__ __forceinline__ void epilogue_f32_tma(const CUtensorMap& map_out, const TcArgs& g, uint32_t tmem_base, uint32_t epi_base,
                                                 uint32_t tfull0, uint32_t tempty0, int worker, int num_workers, uint32_t cta_rank,
                                                 int warp, int lane) {
  const int ew = warp - 2, lane_grp = warp & 3, half = ew >> 2;
  constexpr int CHUNKS = (BN + 31) / 32;
  constexpr int CH_PER_WARP = (CHUNKS + 1) / 2;
  constexpr bool INT_ACC = (KIND == 0 || KIND == 2);
  const Epi& e = g.ep;
  const int num_tiles = g.tiles_m * g.tiles_n;
  const uint32_t my_buf = epi_base + (uint32_t)ew * (4096u * EPB);
  uint32_t buf_sel = 0;
  const int N32 = (int)g.N;
  const bool plain_acc = (KIND == 2) && e.acc_mul == 1 && e.row_sum == nullptr;
  const bool has_cs = e.col_scale != nullptr, has_b = e.bias != nullptr;
  int as = 0;
  uint32_t aphase = 0;
  for (int tile = worker; tile < num_tiles; tile += num_workers) {
    int tm, tn;
    tile_coords<CG>(g, tile, tm, tn);
    const int row0 = (tm * CG + (int)cta_rank) * TC_BM + lane_grp * 32;
    const int64_t m = (int64_t)row0 + lane;
    const int n_tile = tn * BN;
    const int n_lim = min(N32, n_tile + BN);
    const int c_begin = half * CH_PER_WARP;
    int c_end = min(CHUNKS, c_begin + CH_PER_WARP);
    while (c_end > c_begin && n_tile + (c_end - 1) * 32 >= n_lim) --c_end;
    const bool row_ok = m < g.M;
    float mul = e.scale;
    int32_t rsum = 0;
    if (row_ok) {
      if (e.row_scale) {
        if (e.row_scale_parts > 0) {
          float rs = 0.f;
          for (int p = 0; p < e.row_scale_parts; ++p) rs += __ldg(e.row_scale + (int64_t)p * g.M + m);
          mul *= rs * e.row_scale_mul;
        } else {
          mul *= __ldg(e.row_scale + m);
        }
      }
      if (e.row_sum) {
        int32_t rs = 0;
        if (e.row_sum_parts > 0) {
          for (int p = 0; p < e.row_sum_parts; ++p) rs += __ldg(e.row_sum + (int64_t)p * g.M + m);
        } else {
          rs = __ldg(e.row_sum + m);
        }
        rsum = e.rs_mul * rs;
      }
    }
    mbar_wait(tfull0 + 8u * as, aphase);
    tc_fence_after();
    uint32_t r[32];
    const uint32_t t_row = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(lane_grp * 32) << 16);
    if (g.epi_debug == 1) c_end = c_begin;
#pragma unroll 1
    for (int cidx = c_begin; cidx < c_end; ++cidx) {
      const int c0 = cidx * 32;
      const int n0 = n_tile + c0;
      const bool full_chunk = (BN % 32 == 0) || (c0 + 32 <= BN);
      // column parameters first: their (L1 / L2) latency runs under the TMEM read
      float pb[32];
      if (has_b) load_col32(e.bias, n0, N32, true, 0.f, pb);
      tmem_ld32(t_row + (uint32_t)c0, r);
      if (g.epi_debug == 2) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a;
        if (!INT_ACC || plain_acc) a = __uint_as_float(r[j]);
        else if (KIND == 2) a = (float)(e.acc_mul * __float2int_rn(__uint_as_float(r[j])) + rsum);
        else a = (float)(e.acc_mul * (int32_t)r[j] + rsum);
        v[j] = a * mul;
      }
      if (has_cs) {
        float cs[32];
        load_col32(e.col_scale, n0, N32, true, 1.f, cs);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= cs[j];
      }
      if (has_b) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += pb[j];
      }
      if (e.out_clamp) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], e.out_lo), e.out_hi);
      }
      if (g.epi_debug == 3) {      // profiling aid: registers -> global directly (each lane owns 128 contiguous bytes of its row)
        if (row_ok && n0 + 32 <= n_lim) {
          float4* o = reinterpret_cast<float4*>(e.out + m * e.ldo + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) __stcs(o + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
      } else if (full_chunk) {
        if (lane == 0) {
          if (EPB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else tma_store_wait_read0();
        }
        __syncwarp();
        const uint32_t tile_buf = my_buf + buf_sel * 4096u;
        const uint32_t rowaddr = tile_buf + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(rowaddr + (uint32_t)((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(&map_out, tile_buf, n0, row0);
        if (EPB == 2) buf_sel ^= 1u;
      } else if (row_ok) {       // the 16-column tail chunk of a 240-wide tile
        float* o = e.out + m * e.ldo + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (n0 + j + 4 <= n_lim) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CG == 2) mbar_arrive_remote(mapa_shared(tempty0 + 8u * as, 0));
      else mbar_arrive_relaxed(tempty0 + 8u * as);
    }
    if (++as == 2) { as = 0; aphase ^= 1u; }
  }
  if (lane == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// Epilogue, tight form for the conv / linear chains of k-bit (DoReFa) activations: the output goes out as 8-bit codes
//   c = rint(n * clamp(y, lo, hi))           y = acc * (scale * row) * col_scale + bias [+ residual]
// (int8 lane for k <= 7, uint8 for k = 8; the clamp keeps n * y inside the lane, so no overflow bookkeeping), optionally
// together with the clamped fp32 value (channels-last, TMA store) -- the residual-block form.  Same arithmetic, in the same
// order, as the general path (bit-identical codes); ~1/3 of its instructions: the quantizer is FMUL + F2I.RN + I2IP per
// element, no per-element selects, no row-sum / overflow tracking.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack4_sat(bool uns, int k0, int k1, int k2, int k3) {
  uint32_t t, w;
  if (uns) {
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(k3), "r"(k2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(k1), "r"(k0), "r"(t));
  } else {
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(k3), "r"(k2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(k1), "r"(k0), "r"(t));
  }
  return w;
}

__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// cpc != 0: shared-memory cache of the tile's per-column scale / bias (2 x BN floats).  The column parameters of an output
// tile are the same for every tile of that N block, but a 16-byte global load per 4 columns in front of every use costs an
// L1 / L2 round trip per chunk on the epilogue's critical path (ncu: a third of the epilogue warps' stall samples in the conv
// kernels); the eight epilogue warps refresh the cache together when the N block changes and read it with broadcast LDS.128.
template <int BN, int KIND, int CG, int EPB>
__device__ __forceinline__ void epilogue_rq8(const CUtensorMap& map_out, const TcArgs& g, uint32_t tmem_base, uint32_t epi_base,
                                             uint32_t tfull0, uint32_t tempty0, int worker, int num_workers, uint32_t cta_rank,
                                             int warp, int lane, uint32_t cpc) {
  const int ew = warp - 2, lane_grp = warp & 3, half = ew >> 2;
  constexpr int CHUNKS = (BN + 31) / 32;
  constexpr int CH_PER_WARP = (CHUNKS + 1) / 2;
  constexpr bool INT_ACC = (KIND == 0 || KIND == 2);
  const Epi& e = g.ep;
  const int num_tiles = g.tiles_m * g.tiles_n;
  const uint32_t my_buf = epi_base + (uint32_t)ew * (4096u * EPB);
  uint32_t buf_sel = 0;
  const int N32 = (int)g.N;
  const bool plain_acc = (KIND == 2) && e.acc_mul == 1 && e.row_sum == nullptr;
  const bool has_cs = e.col_scale != nullptr, has_b = e.bias != nullptr, has_res = e.residual != nullptr, has_out = e.out != nullptr;
  const bool uns = e.rq_codes_kind == 2;
  const float lo = e.rq_lo, hi = e.rq_hi, qn = e.rq_n;
  uint8_t* const cbase = reinterpret_cast<uint8_t*>(e.rq_codes);
  // Folded column parameters (the BatchNorm-folded conv chains): with nothing per-row attached the pre-activation is
  // acc * (scale * col_scale[n]) + bias[n] -- one FFMA per element against the cached products instead of FMUL, FMUL, FADD.
  // `lean` (no fp32 side output, no residual): the quantizer's scale goes into the cache as well and the float clamp becomes
  // an integer clamp on the code (rn(n * clamp(v, lo, hi)) == clamp(rn(n * v), rn(n * lo), rn(n * hi)): rounding is
  // monotonic), of which the saturating pack supplies whichever side coincides with the lane's range.  The fold removes
  // float roundings, it does not add any: codes differ from the three-rounding form only where a pre-activation sits within
  // an ulp of a rounding boundary (the documented <= 1 code on <= 0.1 % tolerance of a folded BatchNorm).
  const bool fold = cpc != 0 && has_cs && e.row_scale == nullptr &&
                    (KIND == 1 || (KIND == 0 && e.acc_mul == 1 && e.row_sum == nullptr));
  const int klo = __float2int_rn(qn * lo), khi = __float2int_rn(qn * hi);
  const bool lean = fold && !has_res && !has_out && klo <= 0 && khi >= 0;      // pad columns must quantize to code 0
  const bool clamp_lo = klo != (uns ? 0 : -128), clamp_hi = khi != (uns ? 255 : 127);
  int as = 0;
  uint32_t aphase = 0;
  int cached_tn = -1;
  for (int tile = worker; tile < num_tiles; tile += num_workers) {
    int tm, tn;
    tile_coords<CG>(g, tile, tm, tn);
    const int row0 = (tm * CG + (int)cta_rank) * TC_BM + lane_grp * 32;
    const int64_t m = (int64_t)row0 + lane;
    const int n_tile = tn * BN;
    const int n_lim = min(N32, n_tile + BN);
    if (cpc != 0 && tn != cached_tn) {          // every epilogue warp walks the same tile sequence: they all arrive here
      asm volatile("bar.sync 1, 256;" ::: "memory");           // readers of the previous block are done
      for (int c = ew * 32 + lane; c < BN; c += 256) {
        const int n = n_tile + c;
        float csv = (has_cs && n < N32) ? __ldg(e.col_scale + n) : 1.f;
        float bv = (has_b && n < N32) ? __ldg(e.bias + n) : 0.f;
        if (fold) csv *= e.scale;
        if (lean) {                                  // columns past N quantize to code 0 without a per-element test
          csv = n < N32 ? qn * csv : 0.f;
          bv = n < N32 ? qn * bv : 0.f;
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cpc + 4u * (uint32_t)c), "f"(csv) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cpc + 4u * (uint32_t)(BN + c)), "f"(bv) : "memory");
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      cached_tn = tn;
    }
    const int n_cover = max(n_lim, (int)min((int64_t)(n_tile + BN), e.rq_cover));       // zero codes for pad channels
    const int c_begin = half * CH_PER_WARP;
    int c_end = min(CHUNKS, c_begin + CH_PER_WARP);
    while (c_end > c_begin && n_tile + (c_end - 1) * 32 >= n_cover) --c_end;
    const bool row_ok = m < g.M;
    float mul = e.scale;
    int32_t rsum = 0;
    if (row_ok) {
      if (e.row_scale) {
        if (e.row_scale_parts > 0) {
          float rs = 0.f;
          for (int p = 0; p < e.row_scale_parts; ++p) rs += __ldg(e.row_scale + (int64_t)p * g.M + m);
          mul *= rs * e.row_scale_mul;
        } else {
          mul *= __ldg(e.row_scale + m);
        }
      }
      if (e.row_sum) {
        int32_t rs = 0;
        if (e.row_sum_parts > 0) {
          for (int p = 0; p < e.row_sum_parts; ++p) rs += __ldg(e.row_sum + (int64_t)p * g.M + m);
        } else {
          rs = __ldg(e.row_sum + m);
        }
        rsum = e.rs_mul * rs;
      }
      if (has_res) {
        for (int c = c_begin; c < c_end; ++c)
          if (n_tile + c * 32 < n_lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.residual + m * e.ld_res + n_tile + c * 32));
      }
    }
    // the residual rows of the warp's FIRST chunk are requested before the accumulator wait: their L2 / HBM latency runs under
    // this tile's MMAs instead of on the epilogue's critical path (ncu: a quarter of the epilogue warps' time in the second
    // convs of the residual blocks was the scoreboard wait of `y += residual`)
    float res[32];
    const bool res_pre = has_res && row_ok && c_end > c_begin && n_tile + c_begin * 32 + 32 <= n_lim;
    if (res_pre) {
      const float4* rp = reinterpret_cast<const float4*>(e.residual + m * e.ld_res + n_tile + c_begin * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = __ldcs(rp + q);
        res[4 * q] = t.x; res[4 * q + 1] = t.y; res[4 * q + 2] = t.z; res[4 * q + 3] = t.w;
      }
    }
    mbar_wait(tfull0 + 8u * as, aphase);
    tc_fence_after();
    uint32_t r[32];
    const uint32_t t_row = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
    for (int cidx = c_begin; cidx < c_end; ++cidx) {
      const int c0 = cidx * 32;
      const int n0 = n_tile + c0;
      const bool whole = n0 + 32 <= n_lim;            // all 32 columns are real outputs (else: tail / pad-channel chunk)
      if (lean) {
        tmem_ld32(t_row + (uint32_t)c0, r);
        uint32_t w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ca = lds_v4(cpc + 4u * (uint32_t)(c0 + 4 * q));
          const float4 cb = lds_v4(cpc + 4u * (uint32_t)(BN + c0 + 4 * q));
          const float a0 = KIND == 1 ? __uint_as_float(r[4 * q]) : (float)(int32_t)r[4 * q];
          const float a1 = KIND == 1 ? __uint_as_float(r[4 * q + 1]) : (float)(int32_t)r[4 * q + 1];
          const float a2 = KIND == 1 ? __uint_as_float(r[4 * q + 2]) : (float)(int32_t)r[4 * q + 2];
          const float a3 = KIND == 1 ? __uint_as_float(r[4 * q + 3]) : (float)(int32_t)r[4 * q + 3];
          int k0 = __float2int_rn(fmaf(a0, ca.x, cb.x)), k1 = __float2int_rn(fmaf(a1, ca.y, cb.y));
          int k2 = __float2int_rn(fmaf(a2, ca.z, cb.z)), k3 = __float2int_rn(fmaf(a3, ca.w, cb.w));
          if (clamp_lo) { k0 = max(k0, klo); k1 = max(k1, klo); k2 = max(k2, klo); k3 = max(k3, klo); }
          if (clamp_hi) { k0 = min(k0, khi); k1 = min(k1, khi); k2 = min(k2, khi); k3 = min(k3, khi); }
          w[q] = pack4_sat(uns, k0, k1, k2, k3);
        }
        if (row_ok) {
          uint8_t* dst = cbase + (m * e.rq_ld + n0);
          if (n0 < e.rq_cover) st_global_v4(dst, w[0], w[1], w[2], w[3]);
          if (n0 + 16 < e.rq_cover && c0 + 16 < BN) st_global_v4(dst + 16, w[4], w[5], w[6], w[7]);
        }
        continue;
      }
      if (has_res && row_ok && whole && !(res_pre && cidx == c_begin)) {
        const float4* rp = reinterpret_cast<const float4*>(e.residual + m * e.ld_res + n0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = __ldcs(rp + q);
          res[4 * q] = t.x; res[4 * q + 1] = t.y; res[4 * q + 2] = t.z; res[4 * q + 3] = t.w;
        }
      }
      tmem_ld32(t_row + (uint32_t)c0, r);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a;
        if (!INT_ACC || plain_acc) a = __uint_as_float(r[j]);
        else if (KIND == 2) a = (float)(e.acc_mul * __float2int_rn(__uint_as_float(r[j])) + rsum);
        else a = (float)(e.acc_mul * (int32_t)r[j] + rsum);
        v[j] = a;
      }
      if (!fold) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= mul;
      }
      if (fold) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ca = lds_v4(cpc + 4u * (uint32_t)(c0 + 4 * q));
          const float4 cb = lds_v4(cpc + 4u * (uint32_t)(BN + c0 + 4 * q));
          v[4 * q] = fmaf(v[4 * q], ca.x, cb.x); v[4 * q + 1] = fmaf(v[4 * q + 1], ca.y, cb.y);
          v[4 * q + 2] = fmaf(v[4 * q + 2], ca.z, cb.z); v[4 * q + 3] = fmaf(v[4 * q + 3], ca.w, cb.w);
        }
      } else if (cpc != 0) {
        if (has_cs) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = lds_v4(cpc + 4u * (uint32_t)(c0 + 4 * q));
            v[4 * q] *= t.x; v[4 * q + 1] *= t.y; v[4 * q + 2] *= t.z; v[4 * q + 3] *= t.w;
          }
        }
        if (has_b) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = lds_v4(cpc + 4u * (uint32_t)(BN + c0 + 4 * q));
            v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
          }
        }
      } else {
        if (has_cs) {
          float cs[32];
          load_col32(e.col_scale, n0, N32, true, 1.f, cs);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= cs[j];
        }
        if (has_b) {
          float pb[32];
          load_col32(e.bias, n0, N32, true, 0.f, pb);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += pb[j];
        }
      }
      if (has_res && row_ok) {
        if (whole) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += res[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < n_lim) v[j] += __ldg(e.residual + m * e.ld_res + n0 + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], lo), hi);
      // codes: 32 bytes per row chunk, two 16-byte stores
      if (row_ok) {
        uint32_t w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          int k0 = __float2int_rn(qn * v[4 * q]), k1 = __float2int_rn(qn * v[4 * q + 1]);
          int k2 = __float2int_rn(qn * v[4 * q + 2]), k3 = __float2int_rn(qn * v[4 * q + 3]);
          if (!whole) {
            if (n0 + 4 * q >= n_lim) k0 = 0;
            if (n0 + 4 * q + 1 >= n_lim) k1 = 0;
            if (n0 + 4 * q + 2 >= n_lim) k2 = 0;
            if (n0 + 4 * q + 3 >= n_lim) k3 = 0;
          }
          w[q] = pack4_sat(uns, k0, k1, k2, k3);
        }
        uint8_t* dst = cbase + (m * e.rq_ld + n0);
        if (n0 < e.rq_cover) st_global_v4(dst, w[0], w[1], w[2], w[3]);
        if (n0 + 16 < e.rq_cover && c0 + 16 < BN) st_global_v4(dst + 16, w[4], w[5], w[6], w[7]);
      }
      if (has_out && n0 < n_lim) {
        // fp32 tile through the staging buffer (whole chunks; the tensor map clips M / N tails)
        if (lane == 0) {
          if (EPB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else tma_store_wait_read0();
        }
        __syncwarp();
        const uint32_t tile_buf = my_buf + buf_sel * 4096u;
        const uint32_t rowaddr = tile_buf + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(rowaddr + (uint32_t)((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(&map_out, tile_buf, n0, row0);
        if (EPB == 2) buf_sel ^= 1u;
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CG == 2) mbar_arrive_remote(mapa_shared(tempty0 + 8u * as, 0));
      else mbar_arrive_relaxed(tempty0 + 8u * as);
    }
    if (++as == 2) { as = 0; aphase ^= 1u; }
  }
  if (has_out && lane == 0) tma_store_wait_all();
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile -- each CTA
// stages its own 128 A rows and HALF of the B rows, the leader CTA (cluster rank 0) issues every MMA for both SMs, so the
// shared-memory traffic per SM and MMA drops from 12 KB to 8 KB (at cta_group::1 the operand reads + TMA fills of a
// 128 x 256 tile ask for ~190 B/clk of the 128 B/clk shared memory: that, not the tensor pipe, was the limit).
template <int BN, int KIND, int STAGES, int BKB, bool IM2COL, int CG, int EPB>
__device__ __forceinline__ void tc_gemm_body(const CUtensorMap& map_a, const CUtensorMap& map_w, const CUtensorMap& map_out,
                                             const TcArgs& g) {
  static_assert(CG == 1 || BN % 16 == 0, "cta_group::2 needs N % 16 == 0");
  constexpr uint32_t A_BYTES = TC_BM * BKB;
  constexpr uint32_t W_BYTES = (BN / CG) * BKB;         // this CTA's share of the B tile
  constexpr uint32_t STAGE_BYTES = A_BYTES + W_BYTES;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // kind::mxf4: 2 x BN accumulator columns (BN <= 240) + 16 columns of unit scale factors at TC_SF_COL
  constexpr uint32_t TMEM_COLS = (KIND == 2 || 2 * BN > 256) ? 512u : (2 * BN > 128 ? 256u : (2 * BN > 64 ? 128u : 2u * BN));
  static_assert(KIND != 2 || 2 * BN <= TC_SF_COL, "mxf4 tiles must leave the scale-factor columns free");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // 8 epilogue warps x EPB staging tiles of 32 rows x 128 B.  EPB = 2: a warp fills one tile while the bulk store of the
  // previous chunk still reads the other (short-K products, where the epilogue and not the MMA loop sets the tile time)
  static_assert(EPB == 1 || EPB == 2, "one or two staging tiles per epilogue warp");
  constexpr uint32_t EPI_BYTES = 8 * 4096 * EPB;
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t cpc_base = IM2COL ? bar_base + 256u : 0u;       // column-parameter cache of the conv kernels (2 x BN floats)
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + EPI_BYTES + 8u * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (g.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8 * CG);     // the epilogue warps of BOTH CTAs release the leader's accumulator stage
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // one warp of EACH CTA of the pair performs the (collective) cta_group::2 allocation
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (KIND == 2) {
    if (warp >= 2 && warp < 6) {   // four warps, one per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31)
      const uint32_t taddr = tmem_base + TC_SF_COL + ((uint32_t)((warp & 3) * 32) << 16);
      const uint32_t one = 0x7F7F7F7Fu;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(one) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // the peer's scale-factor columns are in place before the leader issues MMAs
    tc_fence_after();
  }

  const int num_tiles = g.tiles_m * g.tiles_n;      // tiles of (128 * CG) x BN
  const int iters = g.npass * g.num_kblocks;
  const int worker = (int)blockIdx.x / CG, num_workers = (int)gridDim.x / CG;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      int tm, tn;
      tile_coords<CG>(g, tile, tm, tn);
      for (int pass = 0; pass < g.npass; ++pass) {
        const int a_row = (int)(g.pa[pass] * g.a_plane_rows) + (tm * CG + (int)cta_rank) * TC_BM;
        const int w_row = (int)(g.pw[pass] * g.w_plane_rows) + tn * BN + (int)cta_rank * (BN / CG);
        int cv_n = 0, cv_h = 0, cv_w = 0;
        if (IM2COL) {   // first output pixel of this CTA's 128 rows -> base position of its filter window
          const int m0 = (tm * CG + (int)cta_rank) * TC_BM;
          cv_n = fastdiv(m0, g.fd_ohw);
          const int r = m0 - cv_n * g.cv_OHW;
          const int oh = fastdiv(r, g.fd_ow);
          cv_h = oh * g.cv_sh - g.cv_ph;
          cv_w = (r - oh * g.cv_OW) * g.cv_sw - g.cv_pw;
        }
        if (!IM2COL && g.ep.a_ready != nullptr) {
          // the A rows of this tile are being written by a kernel on another stream: wait for their progress counter
          const int blk = a_row / g.ep.a_ready_rows;
          if (lane == 0) {
            const int32_t* flag = g.ep.a_ready + blk;
            int32_t v;
            unsigned spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
              if (v >= g.ep.a_ready_target) break;
              __nanosleep(200);
              if (++spins > (1u << 24)) asm volatile("trap;");      // ~ seconds: the producer is not running -- fail loudly, do not hang
            } while (true);
            asm volatile("fence.proxy.async.global;" ::: "memory");  // the generic-proxy writes become visible to the TMA reads
          }
          __syncwarp();
        }
        int cb = 0, off_w = 0, off_h = 0, kx = 0;              // k-block -> (filter tap, channel block), kept incrementally
        for (int kb = 0; kb < g.num_kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          if (elect_one()) {
            // cta_group::2: the MMA issuer waits on the LEADER's full barrier, which counts the bytes of both CTAs' loads
            if (leader) mbar_expect_tx(full_bar(stage), CG * STAGE_BYTES);
            if (CG == 2) {
              const uint32_t lbar = mapa_shared(full_bar(stage), 0);
              if (IM2COL) {
                tma_load_im2col_cg2(sa, &map_a, lbar, g.cv_c0 + cb * BKB, cv_w, cv_h, cv_n, (uint16_t)off_w, (uint16_t)off_h);
              } else {
                tma_load_2d_cg2(sa, &map_a, lbar, kb * BKB, a_row);
              }
              tma_load_2d_cg2(sa + A_BYTES, &map_w, lbar, kb * BKB, w_row);
            } else {
              if (IM2COL) {
                tma_load_im2col(sa, &map_a, full_bar(stage), g.cv_c0 + cb * BKB, cv_w, cv_h, cv_n, (uint16_t)off_w, (uint16_t)off_h);
              } else {
                tma_load_2d(sa, &map_a, full_bar(stage), kb * BKB, a_row);
              }
              tma_load_2d(sa + A_BYTES, &map_w, full_bar(stage), kb * BKB, w_row);
            }
          }
          __syncwarp();
          if (IM2COL) {
            if (++cb == g.cv_cblocks) {
              cb = 0; off_w += g.cv_dw;
              if (++kx == g.cv_kw) { kx = 0; off_w = 0; off_h += g.cv_dh; }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_smem_desc<BKB>(sa);
          const uint64_t bdesc = make_smem_desc<BKB>(sa + A_BYTES);
          if (elect_one()) {               // always the same lane: tcgen05.commit tracks the MMAs of the thread that executes it
#pragma unroll
            for (int k = 0; k < BKB / 32; ++k) {
              // +32 B along K inside the swizzle atom == +2 in the (addr >> 4) field
              tc_mma<KIND, CG>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), g.idesc, (it > 0 || k > 0) ? 1u : 0u,
                               tmem_base + TC_SF_COL);
            }
            // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
            if (CG == 2) tc_commit_cg2(empty_bar(stage)); else tc_commit(empty_bar(stage));
            if (it == iters - 1) {         // accumulator complete -> epilogue(s)
              if (CG == 2) tc_commit_cg2(tfull_bar(as)); else tc_commit(tfull_bar(as));
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (g.epi_fast == 1) {
    epilogue_f32_tma<BN, KIND, CG, EPB>(map_out, g, tmem_base, epi_base, tfull_bar(0), tempty_bar(0), worker, num_workers, cta_rank,
                                        warp, lane);
  } else if (g.epi_fast == 2) {
    if constexpr (BN % 32 == 0)
      epilogue_rq8<BN, KIND, CG, EPB>(map_out, g, tmem_base, epi_base, tfull_bar(0), tempty_bar(0), worker, num_workers, cta_rank,
                                      warp, lane, cpc_base);
  } else {
    // ---- epilogue: 8 warps.  Warp w may only touch TMEM lanes 32*(w%4)..+31; the two warps of a lane quadrant split
    // the BN accumulator columns in halves, so each SM sub-partition always has a second warp to hide latencies.
    const int ew = warp - 2;
    const int lane_grp = warp & 3;
    const int half = ew >> 2;
    constexpr int CHUNKS = (BN + 31) / 32;          // BN = 240 (kind::mxf4): the last chunk holds 16 columns
    constexpr int CH_PER_WARP = (CHUNKS + 1) / 2;
    constexpr bool INT_ACC = (KIND == 0 || KIND == 2);   // kind::mxf4 accumulators are exact integers held in fp32
    int as = 0;
    uint32_t aphase = 0;
    const Epi& e = g.ep;
    const bool vec_ok = (e.out != nullptr) && e.out_mode == 0 && (e.ldo % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(e.out) & 15) == 0);
    const uint32_t my_buf = epi_base + (uint32_t)ew * (4096u * EPB);   // this warp's 32 x 128 B staging tile(s)
    uint32_t buf_sel = 0;
    const int N32 = (int)g.N;
    const bool cs_vec = (reinterpret_cast<uintptr_t>(e.col_scale) & 15) == 0;
    const bool b_vec = (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0;
    const bool res_vec = (e.ld_res % 4 == 0) && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0);
    // kind::mxf4 holds exact integers in fp32: with the identity integer affine they are used as they are
    const bool plain_acc = (KIND == 2) && e.acc_mul == 1 && e.row_sum == nullptr && e.acc_out == nullptr;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      int tm, tn;
      tile_coords<CG>(g, tile, tm, tn);
      const int64_t m = (int64_t)(tm * CG + (int)cta_rank) * TC_BM + lane_grp * 32 + lane;
      const int n_tile = tn * BN;
      const int n_lim = min(N32, n_tile + BN);     // first column past this tile
      // chunks of this warp that hold at least one valid column (warp-uniform)
      const int c_begin = half * CH_PER_WARP;
      int c_end = min(CHUNKS, c_begin + CH_PER_WARP);
      // a requant row may be wider than N (channel padding for the consumer's 128-byte K blocks): those chunks are visited
      // too and receive zero codes
      const int n_cover = (e.rq_mode >= 0) ? max(n_lim, (int)min((int64_t)(n_tile + BN), e.rq_cover)) : n_lim;
      while (c_end > c_begin && n_tile + (c_end - 1) * 32 >= n_cover) --c_end;
      const bool row_ok = m < g.M;
      float mul = e.scale;
      int32_t rsum = 0;
      int64_t nchw_base = 0;
      if (row_ok) {          // row operands: fetched before the accumulator is waited for
        if (e.row_scale) {
          if (e.row_scale_parts > 0) {     // partial row sums left by the previous layer's requant epilogue, fixed order
            float rs = 0.f;
            for (int p = 0; p < e.row_scale_parts; ++p) rs += __ldg(e.row_scale + (int64_t)p * g.M + m);
            mul *= rs * e.row_scale_mul;
          } else {
            mul *= __ldg(e.row_scale + m);
          }
        }
        if (e.row_sum) {
          int32_t rs = 0;
          if (e.row_sum_parts > 0) {
            for (int p = 0; p < e.row_sum_parts; ++p) rs += __ldg(e.row_sum + (int64_t)p * g.M + m);
          } else {
            rs = __ldg(e.row_sum + m);
          }
          rsum = e.rs_mul * rs;
        }
        if (e.out_mode == 1) {
          int64_t img = m / e.nchw_inner, r = m - img * e.nchw_inner;
          nchw_base = img * e.ldo * e.nchw_inner + r;
        }
      }
      if (e.residual && row_ok) {     // residual rows come from HBM: pull this warp's 128-byte row pieces into L2 ahead of use
        for (int c = c_begin; c < c_end; ++c)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(e.residual + m * e.ld_res + n_tile + c * 32));
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      float rq_psum = 0.f;
      int rq_isum = 0;
      bool rq_ovf = false;
      uint32_t r[32];
      const uint32_t t_row = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
      for (int cidx = c_begin; cidx < c_end; ++cidx) {
        const int c0 = cidx * 32;
        const int n0 = n_tile + c0;
        const bool full_chunk = (BN % 32 == 0) || (c0 + 32 <= BN);
        const bool add_res = e.residual != nullptr && row_ok;
        tmem_ld32(t_row + (uint32_t)c0, r);
        if (INT_ACC && e.acc_out && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < n_lim) e.acc_out[m * g.N + n0 + j] = (KIND == 2) ? __float2int_rn(__uint_as_float(r[j])) : (int32_t)r[j];
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a;
          if (!INT_ACC || plain_acc) a = __uint_as_float(r[j]);
          else if (KIND == 2) a = (float)(e.acc_mul * __float2int_rn(__uint_as_float(r[j])) + rsum);
          else a = (float)(e.acc_mul * (int32_t)r[j] + rsum);
          // y = acc * (scale * row_scale) * col_scale + bias; with scale = row/col scale = 1 the multiply is exact and
          // float(acc) + bias is one rounding (BinaryNet / Terner bit-exactness)
          v[j] = a * mul;
        }
        if (!e.out && e.rq_mode < 0) continue;
        if (e.col_scale) {
          float cs[32];
          load_col32(e.col_scale, n0, N32, cs_vec, 1.f, cs);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= cs[j];
        }
        if (e.bias) {
          float bb[32];
          load_col32(e.bias, n0, N32, b_vec, 0.f, bb);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bb[j];
        }
        if (add_res) {
          const float* rp = e.residual + m * e.ld_res + n0;
          if (res_vec && n0 + 32 <= n_lim) {
            float4 t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = __ldcs(reinterpret_cast<const float4*>(rp) + q);
#pragma unroll
            for (int q = 0; q < 8; ++q) { v[4 * q] += t[q].x; v[4 * q + 1] += t[q].y; v[4 * q + 2] += t[q].z; v[4 * q + 3] += t[q].w; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < n_lim) v[j] += __ldg(rp + j);
          }
        }
        if (e.out_clamp) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], e.out_lo), e.out_hi);
        }
        if (e.rq_mode >= 0 && row_ok)
          rq_store_chunk(e, v, m, n0, n_lim, full_chunk ? 32 : (BN - c0), rq_psum, rq_isum, rq_ovf);
        if (!e.out) continue;
        if (g.tma_store && full_chunk) {
          // registers (one output row per lane) -> 128B-swizzled smem tile -> one bulk tensor store per 32x32 block:
          // every global write is a full 128-byte line; M/N tails are clipped by the tensor map.
          if (lane == 0) {      // the store that last used this staging tile has finished reading it
            if (EPB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else tma_store_wait_read0();
          }
          __syncwarp();
          const uint32_t tile_buf = my_buf + buf_sel * 4096u;
          const uint32_t rowaddr = tile_buf + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(rowaddr + (uint32_t)((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&map_out, tile_buf, n0, (tm * CG + (int)cta_rank) * TC_BM + lane_grp * 32);
          if (EPB == 2) buf_sel ^= 1u;
        } else if (e.out_mode == 0) {
          if (!row_ok) continue;
          float* o = e.out + m * e.ldo + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (vec_ok && n0 + j + 4 <= n_lim) {
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int jj = j; jj < j + 4; ++jj)
                if (n0 + jj < n_lim) o[jj] = v[jj];
            }
          }
        } else {
          if (!row_ok) continue;
          float* o = e.out + nchw_base + (int64_t)n0 * e.nchw_inner;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < n_lim) o[(int64_t)j * e.nchw_inner] = v[j];   // lanes = consecutive pixels: coalesced per column
        }
      }
      if (e.rq_mode >= 0) {
        // one partial per (N tile, column half): summed in index order by the consumer (deterministic)
        const int64_t part = (int64_t)(tn * 2 + half) * g.M + m;
        if (row_ok && e.rq_row_part) e.rq_row_part[part] = rq_psum;
        if (row_ok && e.rq_row_sum_part) e.rq_row_sum_part[part] = rq_isum;
        if (rq_ovf && e.rq_overflow) atomicOr(e.rq_overflow, 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_remote(mapa_shared(tempty_bar(as), 0));   // the leader issues the MMAs of both CTAs
        else mbar_arrive_relaxed(tempty_bar(as));
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (g.tma_store && lane == 0) tma_store_wait_all();   // all bulk stores of this warp have completed
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();     // neither CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int KIND, int STAGES, int BKB = TC_BK_BYTES, bool IM2COL = false, int EPB = 1>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_out, TcArgs g) {
  tc_gemm_body<BN, KIND, STAGES, BKB, IM2COL, 1, EPB>(map_a, map_w, map_out, g);
}

// CTA-pair variant: 256 x BN tiles, cluster of two CTAs on the two SMs of a TPC.
template <int BN, int KIND, int STAGES, int EPB = 1, int BKB = TC_BK_BYTES, bool IM2COL = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_out, TcArgs g) {
  tc_gemm_body<BN, KIND, STAGES, BKB, IM2COL, 2, EPB>(map_a, map_w, map_out, g);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2-D byte matrix [rows, row_bytes] with row pitch ld_bytes; box = 128 B x box_rows, 128B swizzle, zero OOB fill.
static CUtensorMapSwizzle swizzle_for(int bkb) {
  return bkb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bkb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t row_bytes, uint64_t ld_bytes, uint32_t box_rows,
                    bool f32 = false, int bkb = TC_BK_BYTES) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return QT_ECUDA; }
  cuuint64_t dims[2] = {f32 ? row_bytes / 4 : row_bytes, rows};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {f32 ? 32u : (cuuint32_t)bkb, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_for(bkb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return QT_ECUDA; }
  return QT_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeIm2colFn)p;
  });
  return fn;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// the tight epilogue serves: fp32 row-major output via TMA store, 16-byte aligned column vectors whose length covers whole
// 32-column chunks up to N (N % 4 == 0), nothing else attached
static bool epi_is_plain_f32(const TcArgs& g) {
  const Epi& e = g.ep;
  static int off = -1;
  if (off < 0) { const char* s = getenv("QTB200_EPI_FAST"); off = (s && atoi(s) == 0) ? 1 : 0; }
  if (off) return false;
  if (!g.tma_store || e.rq_mode >= 0 || e.residual || e.acc_out || e.out_mode != 0 || !e.out) return false;
  if (g.N % 4 != 0) return false;
  if (e.col_scale && (reinterpret_cast<uintptr_t>(e.col_scale) & 15)) return false;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return false;
  return true;
}

// the tight DoReFa-codes epilogue: 8-bit code lanes behind a clamp that keeps n * y inside the lane, no code row sums, an fp32
// side output only in the TMA-store form, aligned column vectors / residual rows
static bool epi_is_rq8(const TcArgs& g) {
  const Epi& e = g.ep;
  static int off = -1;
  if (off < 0) { const char* s = getenv("QTB200_EPI_FAST"); off = (s && atoi(s) == 0) ? 1 : 0; }
  if (off) return false;
  if (e.rq_mode != QT_Q_DOREFA || !(e.rq_codes_kind == 1 || e.rq_codes_kind == 2) || !e.rq_clamp) return false;
  const float lane_lo = e.rq_codes_kind == 1 ? -128.f : 0.f, lane_hi = e.rq_codes_kind == 1 ? 127.f : 255.f;
  if (!(rintf(e.rq_n * e.rq_lo) >= lane_lo && rintf(e.rq_n * e.rq_hi) <= lane_hi)) return false;
  if (e.rq_row_sum_part || e.rq_row_part || e.acc_out || g.N % 4 != 0) return false;
  if (e.out) {
    if (!g.tma_store || e.out_mode != 0) return false;
    if (!e.out_clamp || e.out_lo != e.rq_lo || e.out_hi != e.rq_hi) return false;     // one clamp serves both outputs
  }
  if (e.residual && ((e.ld_res % 4) || (reinterpret_cast<uintptr_t>(e.residual) & 15))) return false;
  if (e.col_scale && (reinterpret_cast<uintptr_t>(e.col_scale) & 15)) return false;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return false;
  return true;
}

constexpr size_t TC_SMEM_MAX = 232448;    // 227 KB opt-in limit per CTA

template <int BN, int KIND, int STAGES, int BKB = TC_BK_BYTES, bool IM2COL = false, int EPB = 1>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mw, TcArgs& g, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (TC_BM * BKB + BN * BKB) + (size_t)EPB * 8 * 4096 + 1024 + 256 + (IM2COL ? 8 * BN : 0);
  static_assert(smem <= TC_SMEM_MAX, "tc_gemm_kernel: shared memory budget");
  // output tensor map (row-major fp32): only when every row pitch / base is 16-byte aligned
  CUtensorMap mo = ma;
  g.tma_store = 0;
  if (g.ep.out && g.ep.out_mode == 0 && g.ep.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(g.ep.out) & 15) == 0) {
    if (int rc = make_map(&mo, g.ep.out, (uint64_t)g.M, (uint64_t)g.N * 4, (uint64_t)g.ep.ldo * 4, 32, true)) return rc;
    g.tma_store = 1;
  }
  g.epi_fast = epi_is_plain_f32(g) ? 1 : (epi_is_rq8(g) ? 2 : 0);
  { const char* dbg = getenv("QTB200_EPI_DEBUG"); g.epi_debug = dbg ? atoi(dbg) : 0; }
  // the opt-in shared-memory size is a per-device function attribute: remember it per device (one process may drive several)
  static bool attr_set[64] = {};
  int dev = 0;
  QT_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    QT_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<BN, KIND, STAGES, BKB, IM2COL, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  g.tiles_m = (int)ceil_div(g.M, TC_BM);
  g.tiles_n = (int)ceil_div(g.N, BN);
  fastdiv_make((uint32_t)(TC_BAND_M * g.tiles_n), g.fd_band);
  int grid = std::min(g.tiles_m * g.tiles_n, num_sms());
  tc_gemm_kernel<BN, KIND, STAGES, BKB, IM2COL, EPB><<<grid, TC_THREADS, smem, stream>>>(ma, mw, mo, g);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

// CTA-pair launch: tiles of 256 x BN, grid = 2 x min(#tiles, #SMs / 2); `mw` must have been built with BN / 2 box rows and
// g.idesc with M = 256.
template <int BN, int KIND, int STAGES, int EPB = 1, int BKB = TC_BK_BYTES, bool IM2COL = false>
static int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mw, TcArgs& g, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (TC_BM * BKB + (BN / 2) * BKB) + (size_t)EPB * 8 * 4096 + 1024 + 256 + (IM2COL ? 8 * BN : 0);
  static_assert(smem <= TC_SMEM_MAX, "tc_gemm2_kernel: shared memory budget");
  CUtensorMap mo = ma;
  g.tma_store = 0;
  if (g.ep.out && g.ep.out_mode == 0 && g.ep.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(g.ep.out) & 15) == 0) {
    if (int rc = make_map(&mo, g.ep.out, (uint64_t)g.M, (uint64_t)g.N * 4, (uint64_t)g.ep.ldo * 4, 32, true)) return rc;
    g.tma_store = 1;
  }
  g.epi_fast = epi_is_plain_f32(g) ? 1 : (epi_is_rq8(g) ? 2 : 0);
  { const char* dbg = getenv("QTB200_EPI_DEBUG"); g.epi_debug = dbg ? atoi(dbg) : 0; }
  static bool attr_set[64] = {};
  int dev = 0;
  QT_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    QT_CUDA_OK(cudaFuncSetAttribute(tc_gemm2_kernel<BN, KIND, STAGES, EPB, BKB, IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  g.tiles_m = (int)ceil_div(g.M, 2 * TC_BM);
  g.tiles_n = (int)ceil_div(g.N, BN);
  fastdiv_make((uint32_t)(TC_BAND_M / 2 * g.tiles_n), g.fd_band);
  const int grid = 2 * std::min(g.tiles_m * g.tiles_n, num_sms() / 2);
  tc_gemm2_kernel<BN, KIND, STAGES, EPB, BKB, IM2COL><<<grid, TC_THREADS, smem, stream>>>(ma, mw, mo, g);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

// 0 = auto (CTA pairs for large tiles), 1 = always cta_group::1, 2 = cta_group::2 whenever the tile shape allows
static int g_cta_group = -1;
static bool use_cta_pair(int64_t M, int bn) {
  if (g_cta_group < 0) {
    const char* s = getenv("QTB200_CTA_GROUP");
    const int v = s ? atoi(s) : 0;
    g_cta_group = (v == 1 || v == 2) ? v : 0;
  }
  if (g_cta_group == 1 || bn < 240) return false;
  if (g_cta_group == 2) return true;
  return M >= 1024;      // enough 256-row tiles to keep the 74 pairs busy
}

// Tile / pipeline variants.  Short K loops (a tile's MMAs take less time than draining its accumulator) run with two staging
// tiles per epilogue warp and one operand stage less; long K loops keep the deeper operand ring.
template <int KIND>
static int dispatch_tc(const CUtensorMap& ma, const CUtensorMap& mw, TcArgs& g, int bn, cudaStream_t stream, bool pair = false) {
  const bool short_k = g.npass * g.num_kblocks <= 24;
  if (pair) return short_k ? launch_tc2<256, KIND, 5, 2>(ma, mw, g, stream) : launch_tc2<256, KIND, 6, 1>(ma, mw, g, stream);
  if (bn == 64) return launch_tc<64, KIND, 6, TC_BK_BYTES, false, 2>(ma, mw, g, stream);
  if (bn == 128) return launch_tc<128, KIND, 5, TC_BK_BYTES, false, 2>(ma, mw, g, stream);
  return short_k ? launch_tc<256, KIND, 3, TC_BK_BYTES, false, 2>(ma, mw, g, stream)
                 : launch_tc<256, KIND, 4, TC_BK_BYTES, false, 1>(ma, mw, g, stream);
}

static int pick_bn(int64_t N) { return N <= 64 ? 64 : (N <= 128 ? 128 : 256); }

// kind::mxf4 tiles: 240 columns (two accumulators + the scale-factor columns fit the 512 TMEM columns) unless 128-wide
// tiles waste clearly fewer padded columns.  QTB200_F4_BN=64|128|240 overrides (tuning / tests).
static int g_f4_tile_n = -1;
static int pick_bn_f4(int64_t N) {
  if (g_f4_tile_n < 0) {
    const char* s = getenv("QTB200_F4_BN");
    int v = s ? atoi(s) : 0;
    g_f4_tile_n = (v == 64 || v == 128 || v == 240) ? v : 0;
  }
  if (g_f4_tile_n) return g_f4_tile_n;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const int64_t pad240 = ceil_div(N, 240) * 240, pad128 = ceil_div(N, 128) * 128;
  return (double)pad128 < 0.9 * (double)pad240 ? 128 : 240;
}

static int dispatch_f4(const CUtensorMap& ma, const CUtensorMap& mw, TcArgs& g, int bn, cudaStream_t stream, bool pair = false) {
  // e2m1 operands carry 4x the MACs per byte of fp16: the K loop of a tile is short, the fp32 drain is what has to keep up
  if (pair) return launch_tc2<240, 2, 5, 2>(ma, mw, g, stream);
  if (bn == 64) return launch_tc<64, 2, 6, TC_BK_BYTES, false, 2>(ma, mw, g, stream);
  if (bn == 128) return launch_tc<128, 2, 5, TC_BK_BYTES, false, 2>(ma, mw, g, stream);
  return launch_tc<240, 2, 3, TC_BK_BYTES, false, 2>(ma, mw, g, stream);
}

static bool tc_available() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, maj = 0, min = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
    ok = (maj == 10 && min == 0) ? 1 : 0;
  }
  return ok == 1;
}

int simt_gemm_i8(const void* a, int a_signed, int64_t lda, const void* w, int w_signed, int64_t ldw,
                 int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, cudaStream_t stream);
int simt_gemm_f16(const void* a, int64_t lda, int64_t a_plane_stride, const void* w, int64_t ldw,
                  int64_t w_plane_stride, int fmt, int npass, const int* pa, const int* pw,
                  int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, cudaStream_t stream);

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace qt

using namespace qt;

extern "C" int qt_gemm_i8(const void* a, int a_signed, int64_t lda, const void* w, int w_signed, int64_t ldw,
                          int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, int backend, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(a && w, "qt_gemm_i8: null operand");
  QT_REQUIRE(M >= 0 && N >= 0 && K > 0 && lda >= K && ldw >= K, "qt_gemm_i8: bad shape");
  if (int rc = check_epi(ep, M, N)) return rc;
  if (M == 0 || N == 0) return QT_OK;
  const bool tc_ok = tc_available() && lda % 16 == 0 && ldw % 16 == 0 && al16(a) && al16(w) &&
                     M < (1ll << 31) && N < (1ll << 31);
  if (backend == 1 && !tc_ok) { set_error("qt_gemm_i8: tcgen05 backend needs sm_100 and 16-byte aligned operands/pitches"); return QT_EUNSUPPORTED; }
  const bool use_tc = backend == 1 || (backend == 0 && tc_ok && (double)M * (double)N * (double)K >= 2.0e5);
  if (!use_tc) return simt_gemm_i8(a, a_signed, lda, w, w_signed, ldw, M, N, K, ep, stream);

  const int bn = pick_bn(N);
  const bool pair = use_cta_pair(M, bn);
  const int cg = pair ? 2 : 1;
  CUtensorMap ma, mw;
  if (int rc = make_map(&ma, a, (uint64_t)M, (uint64_t)K, (uint64_t)lda, TC_BM)) return rc;
  if (int rc = make_map(&mw, w, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(bn / cg))) return rc;
  TcArgs g{};
  g.M = M; g.N = N; g.num_kblocks = (int)ceil_div(K, TC_BK_BYTES); g.npass = 1; g.pa[0] = g.pw[0] = 0;
  g.a_plane_rows = g.w_plane_rows = 0; g.is_int = 1; g.ep = make_epi(ep, M, N);
  // instruction descriptor: D = s32 (2 << 4), A/B = s8 (1) or u8 (0) at bits 7 / 10, K-major both, N >> 3 at 17, M >> 4 at 24
  // (M = 256 for a CTA pair)
  g.idesc = (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((w_signed ? 1u : 0u) << 10) | ((uint32_t)(bn >> 3) << 17) |
            ((uint32_t)((TC_BM * cg) >> 4) << 24);
  if (ep->requant) ep->requant->row_parts = 2 * (int)ceil_div(N, bn);
  return dispatch_tc<0>(ma, mw, g, bn, stream, pair);
}

extern "C" int qt_set_option(const char* name, int value) {
  QT_REQUIRE(name != nullptr, "qt_set_option: null name");
  if (strcmp(name, "f4_tile_n") == 0) {
    QT_REQUIRE(value == 0 || value == 64 || value == 128 || value == 240, "qt_set_option: f4_tile_n must be 0, 64, 128 or 240");
    g_f4_tile_n = value;
    return QT_OK;
  }
  if (strcmp(name, "cta_group") == 0) {
    QT_REQUIRE(value >= 0 && value <= 2, "qt_set_option: cta_group must be 0 (auto), 1 or 2");
    g_cta_group = value;
    return QT_OK;
  }
  set_error("qt_set_option: unknown option '%s'", name);
  return QT_EINVAL;
}

extern "C" int qt_gemm_f4(const void* a, int64_t lda, const void* w, int64_t ldw, int64_t M, int64_t N, int64_t K,
                          const QtEpilogue* ep, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(a && w, "qt_gemm_f4: null operand");
  QT_REQUIRE(M >= 0 && N >= 0 && K > 0 && lda >= K && ldw >= K, "qt_gemm_f4: bad shape");
  QT_REQUIRE(lda % 32 == 0 && ldw % 32 == 0 && al16(a) && al16(w),
             "qt_gemm_f4: e2m1 rows must be 16-byte aligned (lda, ldw multiples of 32 elements)");
  QT_REQUIRE(K < (1ll << 22), "qt_gemm_f4: K too large for exact fp32 accumulation of 2-bit codes");
  if (int rc = check_epi(ep, M, N)) return rc;
  if (M == 0 || N == 0) return QT_OK;
  if (!tc_available() || M >= (1ll << 31) || N >= (1ll << 31)) {
    set_error("qt_gemm_f4: needs an sm_100 device (tcgen05 kind::mxf4); there is no CUDA-core route for e2m1 operands");
    return QT_EUNSUPPORTED;
  }
  const int bn = pick_bn_f4(N);
  const bool pair = use_cta_pair(M, bn);
  const int cg = pair ? 2 : 1;
  CUtensorMap ma, mw;
  // byte matrices: two e2m1 codes per byte, K/2 bytes per row (zero OOB fill = +0.0 codes)
  if (int rc = make_map(&ma, a, (uint64_t)M, (uint64_t)((K + 1) / 2), (uint64_t)(lda / 2), TC_BM)) return rc;
  if (int rc = make_map(&mw, w, (uint64_t)N, (uint64_t)((K + 1) / 2), (uint64_t)(ldw / 2), (uint32_t)(bn / cg))) return rc;
  TcArgs g{};
  g.M = M; g.N = N; g.num_kblocks = (int)ceil_div((K + 1) / 2, TC_BK_BYTES); g.npass = 1; g.pa[0] = g.pw[0] = 0;
  g.a_plane_rows = g.w_plane_rows = 0; g.is_int = 1; g.ep = make_epi(ep, M, N);
  // block-scaled instruction descriptor: A/B = e2m1 (1) at bits 7 / 10, K-major, N >> 3 at 17, scale format ue8m0 (1) at 23,
  // M >> 4 at 24, scale-factor ids 0, K = 64 per instruction
  g.idesc = (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | (1u << 23) | ((uint32_t)((TC_BM * cg) >> 4) << 24);
  if (ep->requant) ep->requant->row_parts = 2 * (int)ceil_div(N, bn);
  return dispatch_f4(ma, mw, g, bn, stream, pair);
}

extern "C" int qt_gemm_f16(const void* a, int64_t lda, int64_t a_plane_stride, const void* w, int64_t ldw,
                           int64_t w_plane_stride, int fmt, int npass, const int* pa, const int* pw,
                           int64_t M, int64_t N, int64_t K, const QtEpilogue* ep, int backend, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(a && w && pa && pw, "qt_gemm_f16: null operand");
  QT_REQUIRE(npass >= 1 && npass <= 8, "qt_gemm_f16: npass must be 1..8");
  QT_REQUIRE(fmt == 0 || fmt == 1, "qt_gemm_f16: fmt must be 0 (bf16) or 1 (fp16)");
  QT_REQUIRE(M >= 0 && N >= 0 && K > 0 && lda >= K && ldw >= K, "qt_gemm_f16: bad shape");
  if (int rc = check_epi(ep, M, N)) return rc;
  QT_REQUIRE(ep->out || ep->requant, "qt_gemm_f16: needs ep->out (or a requant output)");
  if (M == 0 || N == 0) return QT_OK;
  int max_pa = 0, max_pw = 0;
  for (int i = 0; i < npass; ++i) { max_pa = std::max(max_pa, pa[i]); max_pw = std::max(max_pw, pw[i]); }
  bool tc_ok = tc_available() && (lda * 2) % 16 == 0 && (ldw * 2) % 16 == 0 && al16(a) && al16(w);
  if (max_pa > 0) tc_ok = tc_ok && a_plane_stride % lda == 0;
  if (max_pw > 0) tc_ok = tc_ok && w_plane_stride % ldw == 0;
  if (backend == 1 && !tc_ok) { set_error("qt_gemm_f16: tcgen05 backend needs sm_100, 16-byte aligned pitches and plane strides that are whole rows"); return QT_EUNSUPPORTED; }
  const bool use_tc = backend == 1 || (backend == 0 && tc_ok && (double)M * (double)N * (double)K >= 2.0e5);
  if (!use_tc) return simt_gemm_f16(a, lda, a_plane_stride, w, ldw, w_plane_stride, fmt, npass, pa, pw, M, N, K, ep, stream);

  const int bn = pick_bn(N);
  const bool pair = use_cta_pair(M, bn);
  const int cg = pair ? 2 : 1;
  TcArgs g{};
  g.a_plane_rows = max_pa > 0 ? a_plane_stride / lda : 0;
  g.w_plane_rows = max_pw > 0 ? w_plane_stride / ldw : 0;
  const uint64_t a_rows = (uint64_t)(max_pa * g.a_plane_rows + M), w_rows = (uint64_t)(max_pw * g.w_plane_rows + N);
  CUtensorMap ma, mw;
  if (int rc = make_map(&ma, a, a_rows, (uint64_t)K * 2, (uint64_t)lda * 2, TC_BM)) return rc;
  if (int rc = make_map(&mw, w, w_rows, (uint64_t)K * 2, (uint64_t)ldw * 2, (uint32_t)(bn / cg))) return rc;
  g.M = M; g.N = N; g.num_kblocks = (int)ceil_div(K * 2, TC_BK_BYTES); g.npass = npass;
  for (int i = 0; i < npass; ++i) { g.pa[i] = pa[i]; g.pw[i] = pw[i]; }
  g.is_int = 0; g.ep = make_epi(ep, M, N);
  // D = f32 (1 << 4), A/B = bf16 (1) or fp16 (0) at bits 7 / 10
  const uint32_t f = fmt == 0 ? 1u : 0u;
  g.idesc = (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((TC_BM * cg) >> 4) << 24);
  if (ep->requant) ep->requant->row_parts = 2 * (int)ceil_div(N, bn);
  return dispatch_tc<1>(ma, mw, g, bn, stream, pair);
}

// ---------------------------------------------------------------------------------------------
// implicit-GEMM convolution on channels-last activations (TMA im2col mode feeds the A operand):
// 8-bit codes on kind::i8, bf16 values on kind::f16
// ---------------------------------------------------------------------------------------------
namespace qt {
template <int KIND, int BKB>
static int dispatch_conv(const CUtensorMap& ma, const CUtensorMap& mw, TcArgs& g, int bn, cudaStream_t stream, bool pair) {
  // stage = 128 x BKB (pixels) + BN x BKB (filters).  CTA pairs (256-pixel tiles, each CTA stages half of the filter rows) for
  // the 64- and 128-byte channel blocks: the TMA engine works per row, and the filter rows are most of them when N is large
  if constexpr (BKB >= 64) {
    if (pair) {
      if (bn == 64) return launch_tc2<64, KIND, 8, 2, BKB, true>(ma, mw, g, stream);
      if (bn == 128) return launch_tc2<128, KIND, (BKB == 128 ? 6 : 8), 2, BKB, true>(ma, mw, g, stream);
      if constexpr (BKB == 128) {
        if (bn == 192) return launch_tc2<192, KIND, 5, 2, BKB, true>(ma, mw, g, stream);     // N = 192 k: no padded filter columns
      }
      if constexpr (BKB == 128) return launch_tc2<256, KIND, 5, 1, BKB, true>(ma, mw, g, stream);
      else return launch_tc2<256, KIND, 8, 2, BKB, true>(ma, mw, g, stream);
    }
  }
  if (bn == 64) return launch_tc<64, KIND, (BKB == 128 ? 6 : 8), BKB, true, 2>(ma, mw, g, stream);
  if (bn == 128) return launch_tc<128, KIND, (BKB == 128 ? 5 : 8), BKB, true, 2>(ma, mw, g, stream);
  if constexpr (BKB == 128) return launch_tc<256, KIND, 3, BKB, true, 2>(ma, mw, g, stream);
  else return launch_tc<256, KIND, (BKB == 64 ? 6 : 8), BKB, true, 2>(ma, mw, g, stream);
}

// elem_bytes 1: int8 / uint8 codes (KIND 0); 2: bf16 (KIND 1).  C, ldw in ELEMENTS.
static int conv_implicit(const void* x_nhwc, int elem_bytes, int a_signed, const QtConvGeom* cg, const void* w, int w_signed,
                         int64_t ldw, int64_t N, const QtEpilogue* ep, cudaStream_t stream, const char* who) {
  QT_REQUIRE(x_nhwc && cg && w, "%s: null argument", who);
  QT_REQUIRE(cg->groups >= 1 && cg->C % cg->groups == 0 && cg->group >= 0 && cg->group < cg->groups, "%s: bad groups", who);
  const int64_t Cg = cg->C / cg->groups, taps = (int64_t)cg->kh * cg->kw, K = taps * Cg;
  const int64_t P = cg->OH * cg->OW, M = cg->B * P;
  if (int rc = check_epi(ep, M, N)) return rc;
  if (M == 0 || N == 0) return QT_OK;
  const int64_t cgb = Cg * elem_bytes;        // bytes of one pixel's channel group
  const int bkb = (cgb % 128 == 0) ? 128 : ((cgb % 64 == 0) ? 64 : ((cgb % 32 == 0) ? 32 : 0));
  const bool ok = tc_available() && bkb != 0 && (cg->C * elem_bytes) % 16 == 0 && al16(x_nhwc) && al16(w) && (ldw * elem_bytes) % 16 == 0 &&
                  ldw >= K && M < (1ll << 31) && cg->dil_w * (cg->kw - 1) < 65536 && cg->dil_h * (cg->kh - 1) < 65536 &&
                  cg->stride_w <= 8 && cg->stride_h <= 8;
  if (!ok) { set_error("%s: shape not supported by the TMA im2col path (needs sm_100, (C/groups) * element size %% 32 == 0, 16-byte aligned operands)", who); return QT_EUNSUPPORTED; }
  EncodeIm2colFn enc = get_encode_im2col_fn();
  if (!enc) { set_error("cuTensorMapEncodeIm2col entry point not available"); return QT_ECUDA; }

  CUtensorMap ma, mw;
  {
    // byte tensor: the channel axis is counted in bytes so that one map type serves both element sizes
    const cuuint64_t cb = (cuuint64_t)cg->C * elem_bytes;
    cuuint64_t dims[4] = {cb, (cuuint64_t)cg->W, (cuuint64_t)cg->H, (cuuint64_t)cg->B};
    cuuint64_t strides[3] = {cb, (cuuint64_t)cg->W * cb, (cuuint64_t)cg->H * cg->W * cb};
    // bounding box of filter-window base positions: [-pad, (size - 1) + pad - (k - 1) * dil]
    int lower[2] = {-cg->pad_w, -cg->pad_h};
    int upper[2] = {cg->pad_w - (cg->kw - 1) * cg->dil_w, cg->pad_h - (cg->kh - 1) * cg->dil_h};
    if (cg->corner_mode) { lower[0] = cg->lower_w; upper[0] = cg->upper_w; }
    cuuint32_t estr[4] = {1, (cuuint32_t)cg->stride_w, (cuuint32_t)cg->stride_h, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x_nhwc), dims, strides, lower, upper,
                     (cuuint32_t)bkb, (cuuint32_t)TC_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bkb),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed with CUresult %d", (int)r); return QT_EUNSUPPORTED; }
  }
  if (g_cta_group < 0) use_cta_pair(0, 0);                     // reads QTB200_CTA_GROUP once
  const bool pair = bkb >= 64 && g_cta_group != 1 && M >= 148 * 128;     // at least one 256-pixel tile per CTA pair
  int bn = pick_bn(N);
  if (pair && bkb == 128 && N % 192 == 0 && N % 256 != 0) bn = 192;
  const int cgn = pair ? 2 : 1;
  if (int rc = make_map(&mw, w, (uint64_t)N, (uint64_t)K * elem_bytes, (uint64_t)ldw * elem_bytes, (uint32_t)(bn / cgn), false, bkb)) return rc;
  TcArgs g{};
  g.M = M; g.N = N; g.npass = 1; g.pa[0] = g.pw[0] = 0; g.a_plane_rows = g.w_plane_rows = 0; g.is_int = elem_bytes == 1;
  g.ep = make_epi(ep, M, N);
  g.cv_cblocks = (int)(cgb / bkb);
  g.num_kblocks = (int)taps * g.cv_cblocks;
  g.cv_OW = (int)cg->OW; g.cv_OHW = (int)P;
  fastdiv_make((uint32_t)g.cv_OHW, g.fd_ohw);
  fastdiv_make((uint32_t)g.cv_OW, g.fd_ow);
  g.cv_sh = cg->stride_h; g.cv_sw = cg->stride_w; g.cv_ph = cg->pad_h; g.cv_pw = cg->corner_mode ? -cg->lower_w : cg->pad_w;
  g.cv_dh = cg->dil_h; g.cv_dw = cg->dil_w; g.cv_kw = cg->kw; g.cv_c0 = (int)(cgb * cg->group);
  if (ep->requant) ep->requant->row_parts = 2 * (int)ceil_div(N, bn);
  if (elem_bytes == 1) {
    g.idesc = (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((w_signed ? 1u : 0u) << 10) | ((uint32_t)(bn >> 3) << 17) |
              ((uint32_t)((TC_BM * cgn) >> 4) << 24);
    if (bkb == 128) return dispatch_conv<0, 128>(ma, mw, g, bn, stream, pair);
    if (bkb == 64) return dispatch_conv<0, 64>(ma, mw, g, bn, stream, pair);
    return dispatch_conv<0, 32>(ma, mw, g, bn, stream, false);
  }
  // D = f32 (1 << 4), A/B = bf16 (1) at bits 7 / 10
  g.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((TC_BM * cgn) >> 4) << 24);
  if (bkb == 128) return dispatch_conv<1, 128>(ma, mw, g, bn, stream, pair);
  if (bkb == 64) return dispatch_conv<1, 64>(ma, mw, g, bn, stream, pair);
  return dispatch_conv<1, 32>(ma, mw, g, bn, stream, false);
}
}  // namespace qt

extern "C" int qt_conv_i8(const void* x_nhwc, int a_signed, const QtConvGeom* cg, const void* w, int w_signed, int64_t ldw,
                          int64_t N, const QtEpilogue* ep, void* stream_) {
  return conv_implicit(x_nhwc, 1, a_signed, cg, w, w_signed, ldw, N, ep, (cudaStream_t)stream_, "qt_conv_i8");
}

extern "C" int qt_conv_bf16(const void* x_nhwc, const QtConvGeom* cg, const void* w, int64_t ldw, int64_t N, const QtEpilogue* ep,
                            void* stream_) {
  return conv_implicit(x_nhwc, 2, 1, cg, w, 1, ldw, N, ep, (cudaStream_t)stream_, "qt_conv_bf16");
}
