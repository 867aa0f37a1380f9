// Channels-last helpers of the conv path (sm_100a): everything here is HBM-bound elementwise / window work.
//
//   qt_image_planes   fp32 NCHW image -> zero-padded channels-last bf16 "plane pixel" tensor (hi / mid / lo parts of every
//                     channel side by side in one 32-byte pixel): the A operand of the first-layer implicit GEMM
//   qt_image_windows  the same image as row-window records (the kw pixels of a filter row side by side): one k-block per
//                     filter ROW instead of per tap
//   qt_pool_codes     max-pool on channels-last 8-bit activation codes (per-channel max or min)
//   qt_pool_quant_f32 max-pool of a channels-last fp32 activation fused with the next activation quantizer
//   qt_head_f32       global average pool + fp32 Linear head, batch-invariant summation order
#include "qt_common.cuh"

namespace qt {

static bool al(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---------------------------------------------------------------------------------------------
// image planes.  One thread per output pixel (b, hp, wp): reads its C channel values (coalesced along w for every c),
// splits each into P bf16 parts, writes 16 slots = 32 bytes (two 16-byte stores; consecutive threads -> consecutive
// pixels).  Border pixels (the conv's zero padding, materialised) and unused slots are zero.
// ---------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256) image_planes_kernel(const float* __restrict__ x, int B, int C, int H, int W, int Hp, int Wp,
                                                           int pad_h, int pad_w, int fh, int fw, __nv_bfloat16* __restrict__ out) {
  // grid: x = (b, hp) padded rows, y = 256-pixel slabs of the row (no per-pixel 64-bit divisions)
  const int wp = (int)(blockIdx.y * blockDim.x + threadIdx.x);
  if (wp < Wp) {
    const int64_t t = blockIdx.x;
    const int hp = (int)(t % Hp);
    const int64_t b = t / Hp;
    const int h = hp - pad_h, w = wp - pad_w;
    __align__(16) __nv_bfloat16 s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = __float2bfloat16_rn(0.f);
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const float* xp = x + ((b * C) * H + h) * (int64_t)W + w;
      for (int c = 0; c < C; ++c) {
        const float v = __ldg(xp + (int64_t)c * H * W);
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        s[c] = hi;
        if (P >= 2) {
          const float r1 = v - __bfloat162float(hi);
          const __nv_bfloat16 mi = __float2bfloat16_rn(r1);
          s[C + c] = mi;
          if (P >= 3) s[2 * C + c] = __float2bfloat16_rn(r1 - __bfloat162float(mi));
        }
      }
    }
    // space-to-depth: an fh x fw block of pixels is one super pixel of fh * fw * 16 slots (fh == 1: plain row-major pixels)
    const int64_t sp = (b * (Hp / fh) + hp / fh) * (Wp / fw) + wp / fw;
    uint4* o = reinterpret_cast<uint4*>(out + (sp * (fh * fw) + (hp % fh) * fw + (wp % fw)) * 16);
    o[0] = reinterpret_cast<const uint4*>(s)[0];
    o[1] = reinterpret_cast<const uint4*>(s)[1];
  }
}

// ---------------------------------------------------------------------------------------------
// image row windows.  Record (b, hp, ow) holds the kw horizontally adjacent pixels output column ow of a conv row reads
// (w = ow * sw - pad_w + kx), each as P * C plane slots: slot kx * P * C + p * C + c.  With the row's plane pixels laid out
// pixel after pixel (P * C slots each, zero padded borders) a record is the CONTIGUOUS slice starting at pixel ow * sw:
// one block per image row stages that array in shared memory (coalesced reads of the C channel rows, every value split
// once), then every thread assembles 16-byte chunks of records from it -- consecutive threads write consecutive 16 bytes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) image_windows_kernel(const float* __restrict__ x, int C, int H, int W, int P, int kw,
                                                            int sw, int pad_h, int pad_w, int Hp, int OW, int slots,
                                                            __nv_bfloat16* __restrict__ out) {
  extern __shared__ __nv_bfloat16 row[];          // [n_pix][P * C]
  const int pc = P * C, used = kw * pc, chunks = slots >> 3;
  const int chunk_shift = 31 - __clz(chunks);
  const int n_pix = (OW - 1) * sw + kw;
  const int64_t bh = blockIdx.x;                  // b * Hp + hp
  const int hp = (int)(bh % Hp);
  const int64_t b = bh / Hp;
  const int h = hp - pad_h;
  uint4* orow = reinterpret_cast<uint4*>(out) + bh * OW * chunks;
  if (h < 0 || h >= H) {                          // a row of the conv's zero padding
    for (int id = threadIdx.x; id < OW * chunks; id += blockDim.x) orow[id] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float* xrow = x + ((b * C) * H + h) * (int64_t)W;
  for (int c = 0; c < C; ++c)
  for (int i = threadIdx.x; i < n_pix; i += blockDim.x) {
    const int w = i - pad_w;
    const float v = (w >= 0 && w < W) ? __ldg(xrow + (int64_t)c * H * W + w) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16* dst = row + i * pc + c;
    dst[0] = hi;
    if (P >= 2) {
      const float r1 = v - __bfloat162float(hi);
      const __nv_bfloat16 mi = __float2bfloat16_rn(r1);
      dst[C] = mi;
      if (P >= 3) dst[2 * C] = __float2bfloat16_rn(r1 - __bfloat162float(mi));
    }
  }
  __syncthreads();
  const unsigned short* r16 = reinterpret_cast<const unsigned short*>(row);
  for (int id = threadIdx.x; id < OW * chunks; id += blockDim.x) {
    const int ow = id >> chunk_shift, q = id - (ow << chunk_shift);      // chunks is 2, 4, 8 or 16
    const int base = ow * sw * pc + q * 8;
    uint32_t w4[4];
    if ((base & 1) == 0 && q * 8 + 8 <= used) {          // aligned, fully used chunk: four 32-bit shared loads
      const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row) + (base >> 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) w4[j] = r32[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s0 = q * 8 + 2 * j;
        const uint32_t e0 = s0 < used ? r16[base + 2 * j] : 0u;
        const uint32_t e1 = s0 + 1 < used ? r16[base + 2 * j + 1] : 0u;
        w4[j] = e0 | (e1 << 16);
      }
    }
    orow[id] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// pooling on codes: one thread per (output pixel, 16-channel vector); window taps outside the image are skipped
// (-inf padding of nn.MaxPool2d).  use_min[c] != 0 turns channel c into a min-pool: pooling AFTER a per-channel affine
// with a negative scale (a folded BatchNorm) equals the affine of the min.
// ---------------------------------------------------------------------------------------------
struct PoolArgs {
  const void* x;
  void* out;
  const uint8_t* use_min;
  int B, H, W, C, OH, OW;
  int kh, kw, sh, sw, ph, pw;
};

template <bool UNSIGNED>
__device__ __forceinline__ uint32_t vmax4(uint32_t a, uint32_t b) { return UNSIGNED ? __vmaxu4(a, b) : __vmaxs4(a, b); }
template <bool UNSIGNED>
__device__ __forceinline__ uint32_t vmin4(uint32_t a, uint32_t b) { return UNSIGNED ? __vminu4(a, b) : __vmins4(a, b); }

template <bool UNSIGNED, bool HAS_MIN>
__global__ void __launch_bounds__(256) pool_codes_kernel(PoolArgs a) {
  // grid: x = (b, oh) output rows, y = 256-thread slabs of the row's (ow, 16-channel vector) items
  const int cv = a.C >> 4;
  const int item = (int)(blockIdx.y * blockDim.x + threadIdx.x);
  if (item < a.OW * cv) {
    const int ow = item / cv, v = item - ow * cv;
    const int64_t bo = blockIdx.x;
    const int oh = (int)(bo % a.OH);
    const int64_t b = bo / a.OH;
    const uint32_t lo_id = UNSIGNED ? 0u : 0x80808080u, hi_id = UNSIGNED ? 0xFFFFFFFFu : 0x7F7F7F7Fu;
    uint4 mx = make_uint4(lo_id, lo_id, lo_id, lo_id), mn = make_uint4(hi_id, hi_id, hi_id, hi_id);
    const int h0 = oh * a.sh - a.ph, w0 = ow * a.sw - a.pw;
    for (int ky = 0; ky < a.kh; ++ky) {
      const int h = h0 + ky;
      if (h < 0 || h >= a.H) continue;
      for (int kx = 0; kx < a.kw; ++kx) {
        const int w = w0 + kx;
        if (w < 0 || w >= a.W) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a.x) +
                                                              (((b * a.H + h) * a.W + w) * (int64_t)a.C)) + v);
        mx.x = vmax4<UNSIGNED>(mx.x, q.x); mx.y = vmax4<UNSIGNED>(mx.y, q.y);
        mx.z = vmax4<UNSIGNED>(mx.z, q.z); mx.w = vmax4<UNSIGNED>(mx.w, q.w);
        if (HAS_MIN) {
          mn.x = vmin4<UNSIGNED>(mn.x, q.x); mn.y = vmin4<UNSIGNED>(mn.y, q.y);
          mn.z = vmin4<UNSIGNED>(mn.z, q.z); mn.w = vmin4<UNSIGNED>(mn.w, q.w);
        }
      }
    }
    if (HAS_MIN) {
      // byte mask from the per-channel flags (0 / 1 bytes -> 0x00 / 0xFF)
      const uint4 f = __ldg(reinterpret_cast<const uint4*>(a.use_min) + v);
      const uint32_t k0 = (f.x & 0x01010101u) * 0xFFu, k1 = (f.y & 0x01010101u) * 0xFFu;
      const uint32_t k2 = (f.z & 0x01010101u) * 0xFFu, k3 = (f.w & 0x01010101u) * 0xFFu;
      mx.x = (mx.x & ~k0) | (mn.x & k0); mx.y = (mx.y & ~k1) | (mn.y & k1);
      mx.z = (mx.z & ~k2) | (mn.z & k2); mx.w = (mx.w & ~k3) | (mn.w & k3);
    }
    reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out) + (((b * a.OH + oh) * a.OW + ow) * (int64_t)a.C))[v] = mx;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 channels-last max-pool fused with the activation quantizer that follows it (stem of a residual net:
// conv -> BN -> clamp [epilogue] -> max-pool -> {fp32 residual stream, 8-bit codes of the first block}).
// One thread per (output pixel, 4-channel vector).
// ---------------------------------------------------------------------------------------------
struct PoolQuantArgs {
  const float* x;
  float* out;          // optional pooled fp32 [B, OH, OW, C]
  void* codes;         // optional [B, OH, OW, C] int8 / uint8
  int32_t* overflow;
  int mode, codes_kind;
  float n;
  int B, H, W, C, OH, OW;
  int kh, kw, sh, sw, ph, pw;
};

__device__ __forceinline__ int pq_code(int mode, float n, float v, float lo, float hi, bool& ovf) {
  if (mode == QT_Q_SIGN) return (v < 0.f) ? -1 : 1;
  if (mode == QT_Q_TERNARY) {
    const float s = (v < 0.f) ? -1.f : 1.f;
    return (((v < 0.f) ? -1 : 1) + ((v - 0.5f * s < 0.f) ? -1 : 1)) >> 1;
  }
  float c = rintf(n * v);
  if (!(c >= lo && c <= hi)) { ovf = true; c = (c != c) ? 0.f : fminf(fmaxf(c, lo), hi); }
  return (int)c;
}

// KH x KW > 0: compile-time window, every tap's load issued before the first compare (the run-time loops serialise one 16-byte
// load per branch: a third of HBM speed on the 3 x 3 / 2 stem pool); KH == 0: run-time window.
// Grid: x = (b, oh) output rows, y = 256-thread slabs of the row's (ow, 4-channel vector) items -- one 32-bit division per
// thread instead of three 64-bit ones (the kernel was issue-bound on its index arithmetic), max.NaN.f32 for torch's
// NaN-propagating max.
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

template <int KH, int KW>
__global__ void __launch_bounds__(256) pool_quant_f32_kernel(PoolQuantArgs a) {
  const int cv = a.C >> 2;
  const int item = (int)(blockIdx.y * blockDim.x + threadIdx.x);
  const float lo = (a.codes_kind == 1) ? -128.f : 0.f, hi = (a.codes_kind == 1) ? 127.f : 255.f;
  const float ninf = -__int_as_float(0x7f800000);
  bool ovf = false;
  if (item < a.OW * cv) {
    const int ow = item / cv, v = item - ow * cv;
    const int64_t bo = blockIdx.x;                 // b * OH + oh
    const int oh = (int)(bo % a.OH);
    const int64_t b = bo / a.OH;
    float4 m = make_float4(ninf, ninf, ninf, ninf);
    const int h0 = oh * a.sh - a.ph, w0 = ow * a.sw - a.pw;
    const float4* xb = reinterpret_cast<const float4*>(a.x + b * a.H * a.W * (int64_t)a.C) + v;
#define QT_POOL_TAKE(q) m.x = max_nan(m.x, q.x); m.y = max_nan(m.y, q.y); m.z = max_nan(m.z, q.z); m.w = max_nan(m.w, q.w);
    if (KH > 0) {
      float4 q[(KH > 0 ? KH * KW : 1)];
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          const int h = h0 + ky, w = w0 + kx;
          const bool in = h >= 0 && h < a.H && w >= 0 && w < a.W;
          q[ky * KW + kx] = in ? __ldg(xb + (h * a.W + w) * cv) : make_float4(ninf, ninf, ninf, ninf);
        }
      }
#pragma unroll
      for (int i = 0; i < KH * KW; ++i) { QT_POOL_TAKE(q[i]) }
    } else {
      for (int ky = 0; ky < a.kh; ++ky) {
        const int h = h0 + ky;
        if (h < 0 || h >= a.H) continue;
        for (int kx = 0; kx < a.kw; ++kx) {
          const int w = w0 + kx;
          if (w < 0 || w >= a.W) continue;
          const float4 q = __ldg(xb + (h * a.W + w) * cv);
          QT_POOL_TAKE(q)
        }
      }
    }
#undef QT_POOL_TAKE
    const int64_t o = (bo * a.OW + ow) * (int64_t)a.C;
    if (a.out) __stcs(reinterpret_cast<float4*>(a.out + o) + v, m);
    if (a.codes) {
      const int k0 = pq_code(a.mode, a.n, m.x, lo, hi, ovf), k1 = pq_code(a.mode, a.n, m.y, lo, hi, ovf);
      const int k2 = pq_code(a.mode, a.n, m.z, lo, hi, ovf), k3 = pq_code(a.mode, a.n, m.w, lo, hi, ovf);
      reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(a.codes) + o)[v] =
          (uint32_t)(k0 & 0xff) | ((uint32_t)(k1 & 0xff) << 8) | ((uint32_t)(k2 & 0xff) << 16) | ((uint32_t)(k3 & 0xff) << 24);
    }
  }
  if (a.overflow && __syncthreads_or(ovf) && threadIdx.x == 0) atomicOr(a.overflow, 1);
}

// ---------------------------------------------------------------------------------------------
// fp32 classifier head: out[b, n] = bias[n] + sum_c mean_hw(x[b, hw, c]) * w[n, c]   (global average pool + nn.Linear of the
// residual nets, models/Resnet/Resnet_bin.py:104-107; hw = 1: a plain fp32 Linear).  One block per sample, every sum in a
// FIXED order that depends on (HW, C) only: the result for a sample does not depend on the batch it arrives in -- a library
// GEMM picks split-K by batch size, which made the logits of a 2048-image batch differ in the last bit from the same images
// run as four shards.  The work is tiny (B x N x C MACs).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_f32_kernel(const float* __restrict__ x, int HW, int C, const float* __restrict__ w,
                                                       int64_t ldw, const float* __restrict__ bias, int N, float* __restrict__ out,
                                                       int64_t ldo) {
  extern __shared__ float mean[];              // [C]
  const int64_t b = blockIdx.x;
  const float* xb = x + b * (int64_t)HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < HW; ++i) s += __ldg(xb + (int64_t)i * C + c);
    mean[c] = HW > 1 ? s / (float)HW : s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < N; n += 8) {
    const float* wn = w + n * ldw;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(mean[c], __ldg(wn + c), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[b * ldo + n] = acc + (bias ? __ldg(bias + n) : 0.f);
  }
}

static unsigned grid_for(int64_t total, int threads) {
  int64_t blocks = ceil_div(total, threads);
  const int64_t cap = 148ll * 64;          // grid-stride beyond ~8 resident waves
  return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace qt

using namespace qt;

extern "C" int qt_image_planes(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int planes, int pad_h, int pad_w,
                               int64_t Hp, int64_t Wp, int fold_h, int fold_w, void* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x && out, "qt_image_planes: null argument");
  QT_REQUIRE(planes >= 1 && planes <= 3 && C >= 1 && planes * C <= 16, "qt_image_planes: planes * C must fit the 16 slots of a pixel");
  QT_REQUIRE(B >= 0 && H > 0 && W > 0 && Hp > 0 && Wp > 0 && pad_h >= 0 && pad_w >= 0, "qt_image_planes: bad shape");
  QT_REQUIRE(al(out, 16), "qt_image_planes: out must be 16-byte aligned");
  QT_REQUIRE(fold_h >= 1 && fold_w >= 1 && Hp % fold_h == 0 && Wp % fold_w == 0, "qt_image_planes: the folds must divide Hp / Wp");
  QT_REQUIRE(B * Hp * Wp < (1ll << 40) && H * W < (1ll << 31), "qt_image_planes: tensor too large");
  if (B == 0) return QT_OK;
  QT_REQUIRE(B * Hp < (1ll << 31) && ceil_div(Wp, 256) <= 65535, "qt_image_planes: image too large");
  const dim3 grid((unsigned)(B * Hp), (unsigned)ceil_div(Wp, 256));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (planes == 1) image_planes_kernel<1><<<grid, 256, 0, stream>>>(x, (int)B, (int)C, (int)H, (int)W, (int)Hp, (int)Wp, pad_h, pad_w, fold_h, fold_w, o);
  else if (planes == 2) image_planes_kernel<2><<<grid, 256, 0, stream>>>(x, (int)B, (int)C, (int)H, (int)W, (int)Hp, (int)Wp, pad_h, pad_w, fold_h, fold_w, o);
  else image_planes_kernel<3><<<grid, 256, 0, stream>>>(x, (int)B, (int)C, (int)H, (int)W, (int)Hp, (int)Wp, pad_h, pad_w, fold_h, fold_w, o);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_image_windows(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int planes, int kw, int stride_w,
                                int pad_h, int pad_w, int64_t Hp, int64_t OW, int slots, void* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x && out, "qt_image_windows: null argument");
  QT_REQUIRE(planes >= 1 && planes <= 3 && C >= 1 && kw >= 1 && stride_w >= 1, "qt_image_windows: bad planes / kw / stride");
  QT_REQUIRE((slots == 16 || slots == 32 || slots == 64 || slots == 128) && (int64_t)kw * planes * C <= slots, "qt_image_windows: kw * planes * C must fit the record's slots (16, 32, 64 or 128)");
  QT_REQUIRE(B >= 0 && H > 0 && W > 0 && Hp > 0 && OW > 0 && pad_h >= 0 && pad_w >= 0, "qt_image_windows: bad shape");
  QT_REQUIRE(al(out, 16), "qt_image_windows: out must be 16-byte aligned");
  QT_REQUIRE(B * Hp * OW * slots < (1ll << 42) && C * H * W < (1ll << 31), "qt_image_windows: tensor too large");
  if (B == 0) return QT_OK;
  const int64_t n_pix = (OW - 1) * stride_w + kw;
  const size_t smem = (size_t)(n_pix * planes * C) * 2;
  QT_REQUIRE(smem <= 48 * 1024 && B * Hp < (1ll << 31), "qt_image_windows: image row too wide for the shared-memory row buffer");
  image_windows_kernel<<<(unsigned)(B * Hp), 128, smem, stream>>>(x, (int)C, (int)H, (int)W, planes, kw, stride_w, pad_h, pad_w, (int)Hp,
                                                                    (int)OW, slots, reinterpret_cast<__nv_bfloat16*>(out));
  QT_LAUNCH_CHECK();
  return QT_OK;
}

static int pool_geom_ok(const QtPoolGeom* g, const char* who) {
  QT_REQUIRE(g, "%s: null geometry", who);
  QT_REQUIRE(g->B >= 0 && g->H > 0 && g->W > 0 && g->C > 0 && g->kh > 0 && g->kw > 0 && g->stride_h > 0 && g->stride_w > 0 &&
             g->pad_h >= 0 && g->pad_w >= 0 && 2 * g->pad_h <= g->kh && 2 * g->pad_w <= g->kw, "%s: bad geometry", who);
  QT_REQUIRE(g->OH == (g->H + 2 * g->pad_h - g->kh) / g->stride_h + 1 && g->OW == (g->W + 2 * g->pad_w - g->kw) / g->stride_w + 1 &&
             g->OH > 0 && g->OW > 0, "%s: OH / OW do not match floor((size + 2 pad - k) / stride) + 1", who);
  QT_REQUIRE(g->B < (1ll << 31) && g->H < (1 << 20) && g->W < (1 << 20) && g->C < (1 << 24), "%s: tensor too large", who);
  return QT_OK;
}

extern "C" int qt_pool_codes(const void* x_nhwc, int is_unsigned, const QtPoolGeom* g, const uint8_t* use_min, void* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x_nhwc && out, "qt_pool_codes: null argument");
  if (int rc = pool_geom_ok(g, "qt_pool_codes")) return rc;
  QT_REQUIRE(g->C % 16 == 0 && al(x_nhwc, 16) && al(out, 16) && (!use_min || al(use_min, 16)),
             "qt_pool_codes: needs C %% 16 == 0 and 16-byte aligned buffers");
  if (g->B == 0) return QT_OK;
  PoolArgs a;
  a.x = x_nhwc; a.out = out; a.use_min = use_min;
  a.B = (int)g->B; a.H = (int)g->H; a.W = (int)g->W; a.C = (int)g->C; a.OH = (int)g->OH; a.OW = (int)g->OW;
  a.kh = g->kh; a.kw = g->kw; a.sh = g->stride_h; a.sw = g->stride_w; a.ph = g->pad_h; a.pw = g->pad_w;
  QT_REQUIRE(ceil_div(g->OW * (g->C / 16), 256) <= 65535, "qt_pool_codes: image too large");
  const dim3 grid((unsigned)(g->B * g->OH), (unsigned)ceil_div(g->OW * (g->C / 16), 256));
  if (is_unsigned) {
    if (use_min) pool_codes_kernel<true, true><<<grid, 256, 0, stream>>>(a);
    else pool_codes_kernel<true, false><<<grid, 256, 0, stream>>>(a);
  } else {
    if (use_min) pool_codes_kernel<false, true><<<grid, 256, 0, stream>>>(a);
    else pool_codes_kernel<false, false><<<grid, 256, 0, stream>>>(a);
  }
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_head_f32(const float* x, int64_t B, int64_t HW, int64_t C, const float* w, int64_t ldw, const float* bias,
                           int64_t N, float* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x && w && out, "qt_head_f32: null argument");
  QT_REQUIRE(B >= 0 && HW >= 1 && C >= 1 && N >= 1 && ldw >= C && ldo >= N, "qt_head_f32: bad shape");
  QT_REQUIRE(C <= 12288 && HW < (1 << 20) && N < (1 << 20) && B < (1ll << 31), "qt_head_f32: C must fit the 48 KB mean buffer");
  if (B == 0) return QT_OK;
  head_f32_kernel<<<(unsigned)B, 256, (size_t)C * 4, stream>>>(x, (int)HW, (int)C, w, ldw, bias, (int)N, out, ldo);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_pool_quant_f32(const float* x_nhwc, const QtPoolGeom* g, float* out, int mode, int bit_width, void* codes,
                                 int codes_kind, int32_t* overflow, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x_nhwc && (out || codes), "qt_pool_quant_f32: null argument");
  if (int rc = pool_geom_ok(g, "qt_pool_quant_f32")) return rc;
  QT_REQUIRE(g->C % 4 == 0 && al(x_nhwc, 16) && (!out || al(out, 16)) && (!codes || al(codes, 4)),
             "qt_pool_quant_f32: needs C %% 4 == 0 and aligned buffers");
  if (codes) {
    QT_REQUIRE(mode == QT_Q_SIGN || mode == QT_Q_TERNARY || mode == QT_Q_DOREFA, "qt_pool_quant_f32: SIGN / TERNARY / DOREFA codes only");
    QT_REQUIRE(codes_kind == 1 || codes_kind == 2, "qt_pool_quant_f32: int8 / uint8 code lanes only");
    QT_REQUIRE(mode != QT_Q_DOREFA || (bit_width >= 2 && bit_width <= 8), "qt_pool_quant_f32: DoReFa bit width must be 2..8");
    QT_REQUIRE(codes_kind == 1 || mode == QT_Q_DOREFA, "qt_pool_quant_f32: sign / ternary codes need the int8 lane");
  }
  if (g->B == 0) return QT_OK;
  PoolQuantArgs a;
  a.x = x_nhwc; a.out = out; a.codes = codes; a.overflow = overflow; a.mode = mode; a.codes_kind = codes_kind;
  a.n = (mode == QT_Q_DOREFA) ? (float)((1 << bit_width) - 1) : 1.f;
  a.B = (int)g->B; a.H = (int)g->H; a.W = (int)g->W; a.C = (int)g->C; a.OH = (int)g->OH; a.OW = (int)g->OW;
  a.kh = g->kh; a.kw = g->kw; a.sh = g->stride_h; a.sw = g->stride_w; a.ph = g->pad_h; a.pw = g->pad_w;
  QT_REQUIRE(g->H * g->W * (g->C / 4) < (1ll << 31) && ceil_div(g->OW * (g->C / 4), 256) <= 65535, "qt_pool_quant_f32: image too large");
  const dim3 grid((unsigned)(g->B * g->OH), (unsigned)ceil_div(g->OW * (g->C / 4), 256));
  if (g->kh == 3 && g->kw == 3) pool_quant_f32_kernel<3, 3><<<grid, 256, 0, stream>>>(a);
  else if (g->kh == 2 && g->kw == 2) pool_quant_f32_kernel<2, 2><<<grid, 256, 0, stream>>>(a);
  else pool_quant_f32_kernel<0, 0><<<grid, 256, 0, stream>>>(a);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

// ---------------------------------------------------------------------------------------------
// Logits gather over NVLink, push form (the only collective of the path is one all-gather of the fp32 logits).
// One small kernel copies this rank's shard into its rows of EVERY peer's gathered buffer with 16-byte stores to peer-mapped
// addresses: the shard is read once, the stores are posted (no round trip), and a few dozen CTAs without shared memory sit
// beside the persistent tcgen05 CTAs of the next step, so the transfer overlaps the contractions instead of trailing them.
// ---------------------------------------------------------------------------------------------
namespace qt {
struct PushArgs {
  const uint4* src;
  uint4* dst[8];
  int ndst;
  int64_t nvec;
};

__global__ void __launch_bounds__(256) peer_push_kernel(PushArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.nvec; i += stride * 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < a.nvec) v[u] = __ldcs(a.src + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i + u * stride < a.nvec) {
        for (int d = 0; d < a.ndst; ++d) a.dst[d][i + u * stride] = v[u];
      }
    }
  }
  __threadfence_system();
}
}  // namespace qt

extern "C" int qt_peer_push(const void* src, void* const* dst, int ndst, int64_t bytes, int ctas, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(src && dst && ndst >= 1 && ndst <= 8, "qt_peer_push: 1..8 destinations");
  QT_REQUIRE(bytes >= 0 && bytes % 16 == 0 && al(src, 16), "qt_peer_push: 16-byte aligned source and size");
  PushArgs a;
  a.src = reinterpret_cast<const uint4*>(src);
  a.ndst = ndst;
  a.nvec = bytes / 16;
  for (int d = 0; d < ndst; ++d) {
    QT_REQUIRE(dst[d] && al(dst[d], 16), "qt_peer_push: destinations must be 16-byte aligned");
    a.dst[d] = reinterpret_cast<uint4*>(dst[d]);
  }
  if (a.nvec == 0) return QT_OK;
  if (ctas <= 0) ctas = 32;
  peer_push_kernel<<<ctas, 256, 0, stream>>>(a);
  QT_LAUNCH_CHECK();
  return QT_OK;
}
