// Activation quantizers, weight packers/expanders and the im2col gather.
// All kernels are HBM-bound streaming passes: 128-bit coalesced loads, one warp per row (or row chunk),
// warp-shuffle reductions, and every derived output (fp32 fake-quant value, low-bit codes, packed bits,
// row sums) produced from a single read of the fp32 source.
#include <math.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "qt_common.cuh"

namespace qt {

// ---------------------------------------------------------------------------------------------
// per-element quantizer semantics (must match oracle/quanttorch_oracle.py bit for bit)
// ---------------------------------------------------------------------------------------------
struct QParams {
  int mode;
  float n;        // DoReFa 2^k - 1
  float inv_n;    // fl(1 / n)
  float lo_e, hi_e;  // Log: exponent clamp
  float step, maxv;  // Lin
  int with_sign;
  // fused BatchNorm(eval) + clamp in front of the quantizer
  const float* pre_scale;
  const float* pre_shift;
  int64_t pre_channels, pre_hw;
  int pre_clamp;
  float pre_lo, pre_hi;
};

__device__ __forceinline__ float pre_apply(const QParams& q, float x, int64_t ch) {
  x = fmaf(x, __ldg(q.pre_scale + ch), __ldg(q.pre_shift + ch));
  if (q.pre_clamp) x = fminf(fmaxf(x, q.pre_lo), q.pre_hi);
  return x;
}
template <bool PRE>
__device__ __forceinline__ float pre_col(const QParams& q, float x, int64_t col) {
  if (!PRE) return x;
  return pre_apply(q, x, (col / q.pre_hw) % q.pre_channels);
}

struct QOut {
  float y;     // fake-quant fp32 value (what the reference op returns)
  float code;  // integer-valued code
};

__device__ __forceinline__ float sign3(float x) { return (float)((x > 0.f) - (x < 0.f)); }  // torch.sign (NaN -> 0)
__device__ __forceinline__ float sign_safe(float x) { return (x < 0.f) ? -1.f : 1.f; }      // safeSign

template <int MODE>
__device__ __forceinline__ QOut quant_elem(const QParams& q, float x, float row_mean) {
  QOut o;
  switch (MODE) {
    case QT_Q_SIGN: {
      o.y = o.code = sign_safe(x);
      break;
    }
    case QT_Q_TERNARY: {
      float s = sign_safe(x);
      float t = x - 0.5f * s;
      o.y = o.code = (s + sign_safe(t)) * 0.5f;
      break;
    }
    case QT_Q_DOREFA: {
      float c = rintf(q.n * x);  // torch.round == round-half-to-even
      o.code = c;
      o.y = q.inv_n * c;
      break;
    }
    case QT_Q_XNOR_ROW: {
      float s = sign3(x);
      o.code = s;
      o.y = s * row_mean;
      break;
    }
    case QT_Q_LOG: {
      float e = fminf(fmaxf(rintf(log2f(fabsf(x))), q.lo_e), q.hi_e);
      float p = exp2f(e);
      o.y = q.with_sign ? sign3(x) * p : p;
      o.code = o.y;                 // a power of two (or 0): exact in a bf16 lane
      break;
    }
    case QT_Q_LIN: {
      if (q.with_sign) {
        o.y = sign3(x) * fminf(fmaxf(rintf(fabsf(x) / q.step) * q.step, 0.f), q.maxv);
      } else {
        o.y = fminf(fmaxf(rintf(x / q.step) * q.step, 0.f), q.maxv);
      }
      o.code = o.y / q.step;        // integer in [-2^bw, 2^bw]: step is a power of two, the division is exact
      break;
    }
    default: {  // QT_Q_SPLIT
      o.y = x;
      o.code = x;
      break;
    }
  }
  return o;
}

struct ActArgs {
  QParams q;
  const float* x;
  int64_t rows, cols, ld_x;
  float* y;
  int64_t ld_y;
  void* codes;
  int codes_kind;
  int64_t ld_codes;
  uint32_t* bits;
  int64_t ld_bits;
  int32_t* row_sum;
  float* row_scale;
  int32_t* overflow;
  int64_t chunk;    // columns per warp task (multiple of 128)
  int nchunks;
};

__device__ __forceinline__ int code_to_lane(float c, int kind, bool& ovf) {
  // kind 1: int8, kind 2: uint8.  NaN / out-of-lane codes saturate and raise the sticky overflow flag.
  float lo = (kind == 1) ? -128.f : 0.f, hi = (kind == 1) ? 127.f : 255.f;
  if (!(c >= lo && c <= hi)) {
    ovf = true;
    c = (c != c) ? 0.f : fminf(fmaxf(c, lo), hi);
  }
  return (int)c;
}

// integer code in [-4, 4] -> e2m1 nibble (sign | magnitude code: 0->0, 1->2, 2->4, 3->5, 4->6); `k` receives the integer
__device__ __forceinline__ uint32_t code_to_f4(float c, int& k, bool& ovf) {
  if (!(c >= -4.f && c <= 4.f)) {
    ovf = true;
    c = (c != c) ? 0.f : fminf(fmaxf(c, -4.f), 4.f);
  }
  k = (int)c;
  const int m = k < 0 ? -k : k;
  return ((0x65420u >> (4 * m)) & 0xFu) | (k < 0 ? 8u : 0u);
}
__device__ __forceinline__ uint32_t f4_nibble(float v) {   // exact small integers only (weights)
  int k; bool o = false;
  return code_to_f4(v, k, o);
}

// Mode-aware variants: the codes of the sign / ternary / XnorNet quantizers are in {-1, 0, +1}, so the integer comes from
// two predicates (no float->int conversion on the quarter-rate pipe, no range check).
template <int MODE>
__device__ __forceinline__ int code_lane_m(float c, int kind, bool& ovf) {
  if (MODE == QT_Q_SIGN || MODE == QT_Q_TERNARY || MODE == QT_Q_XNOR_ROW) {
    const int k = (int)(c > 0.f) - (int)(c < 0.f);
    return (kind == 2 && k < 0) ? (ovf = true, 0) : k;
  }
  return code_to_lane(c, kind, ovf);
}
template <int MODE>
__device__ __forceinline__ uint32_t code_f4_m(float c, int& k, bool& ovf) {
  if (MODE == QT_Q_SIGN || MODE == QT_Q_TERNARY || MODE == QT_Q_XNOR_ROW) {
    const bool pos = c > 0.f, neg = c < 0.f;
    k = (int)pos - (int)neg;
    return (pos ? 0x2u : 0u) | (neg ? 0xAu : 0u);
  }
  return code_to_f4(c, k, ovf);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int MODE, bool VEC, bool PRE>
__global__ void __launch_bounds__(256, (MODE == QT_Q_XNOR_ROW) ? 3 : 4) act_quant_kernel(ActArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= a.rows * a.nchunks) return;
  const int64_t row = task / a.nchunks;
  const int chunk_id = (int)(task - row * a.nchunks);
  const int64_t c0 = (int64_t)chunk_id * a.chunk;
  const int64_t c1 = min(a.cols, c0 + a.chunk);
  const float* xr = a.x + row * a.ld_x;

  float row_mean = 0.f;
  // XnorNet: the codes sign(x) do not depend on the row mean; only the fp32 fake-quant tensor sign(x) * mean does.  Without
  // a `y` output (code-only chains) the row is therefore read ONCE: the sum is accumulated inside the main loop below.
  const bool xnor_one_pass = (MODE == QT_Q_XNOR_ROW) && a.y == nullptr;
  double xsum = 0.0;
  if (MODE == QT_Q_XNOR_ROW && !xnor_one_pass) {  // never chunked: the whole row is reduced by this warp
    double s = 0.0;
    if (VEC) {
      // 8 independent 16-byte loads in flight per lane (the reduction is latency-bound otherwise); fp32 partial sums of 4
      // are folded into the fp64 accumulator
      for (int64_t c0r = 4 * lane; c0r < a.cols; c0r += 128 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t c = c0r + u * 128;
          v[u] = (c < a.cols) ? __ldg(reinterpret_cast<const float4*>(xr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float part = 0.f;      // fp32 over 32 values, folded into the fp64 accumulator once per batch
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t c = c0r + u * 128;
          if (c < a.cols)
            part += (pre_col<PRE>(a.q, v[u].x, c) + pre_col<PRE>(a.q, v[u].y, c + 1)) +
                    (pre_col<PRE>(a.q, v[u].z, c + 2) + pre_col<PRE>(a.q, v[u].w, c + 3));
        }
        s += (double)part;
      }
    } else {
      for (int64_t c = lane; c < a.cols; c += 32) s += (double)pre_col<PRE>(a.q, __ldg(xr + c), c);
    }
    s = warp_sum_d(s);
    row_mean = (float)(s / (double)a.cols);
    if (a.row_scale && lane == 0) a.row_scale[row] = row_mean;
  }

  int isum = 0;
  bool ovf = false;
  float* yr = a.y ? a.y + row * a.ld_y : nullptr;
  int8_t* c8 = (a.codes_kind == 1 || a.codes_kind == 2) ? reinterpret_cast<int8_t*>(a.codes) + row * a.ld_codes : nullptr;
  __nv_bfloat16* cb = (a.codes_kind >= 3 && a.codes_kind <= 6) ? reinterpret_cast<__nv_bfloat16*>(a.codes) + row * a.ld_codes : nullptr;
  uint8_t* c4 = (a.codes_kind == 7) ? reinterpret_cast<uint8_t*>(a.codes) + row * (a.ld_codes >> 1) : nullptr;   // e2m1 pairs
  const bool f16 = a.codes_kind == 5;   // same 2-byte lanes, IEEE half encoding
  __nv_bfloat16* cb_lo = (a.codes_kind == 4 || a.codes_kind == 6) ? cb + a.rows * a.ld_codes : nullptr;
  __nv_bfloat16* cb_lo2 = (a.codes_kind == 6) ? cb_lo + a.rows * a.ld_codes : nullptr;
  uint32_t* br = a.bits ? a.bits + row * a.ld_bits : nullptr;

  if (VEC) {
    constexpr int U = 4;   // 4 independent 16-byte loads in flight per lane (2 KB per warp) before any use
    for (int64_t base0 = c0; base0 < c1; base0 += 128 * U) {
     float4 vv[U];
     float xpart = 0.f;     // XnorNet one-pass: fp32 sum of this batch (U x 4 values per lane), then one fp64 add
#pragma unroll
     for (int u = 0; u < U; ++u) {
       const int64_t cu = base0 + u * 128 + 4 * lane;
       vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
       if (cu < c1) vv[u] = __ldcs(reinterpret_cast<const float4*>(xr + cu));
     }
#pragma unroll
     for (int u = 0; u < U; ++u) {
      const int64_t base = base0 + u * 128;
      if (base >= c1) break;   // warp-uniform
      const int64_t c = base + 4 * lane;
      const bool valid = c < c1;
      const float4 v = vv[u];
      const float p0 = pre_col<PRE>(a.q, v.x, c), p1 = pre_col<PRE>(a.q, v.y, c + 1);
      const float p2 = pre_col<PRE>(a.q, v.z, c + 2), p3 = pre_col<PRE>(a.q, v.w, c + 3);
      if (MODE == QT_Q_XNOR_ROW && xnor_one_pass && valid) xpart += (p0 + p1) + (p2 + p3);   // fp32 over <= 16 values
      QOut o0 = quant_elem<MODE>(a.q, p0, row_mean), o1 = quant_elem<MODE>(a.q, p1, row_mean);
      QOut o2 = quant_elem<MODE>(a.q, p2, row_mean), o3 = quant_elem<MODE>(a.q, p3, row_mean);
      if (valid) {
        if (yr) __stcs(reinterpret_cast<float4*>(yr + c), make_float4(o0.y, o1.y, o2.y, o3.y));   // streamed: never re-read here
        if (c8) {
          int k0 = code_lane_m<MODE>(o0.code, a.codes_kind, ovf), k1 = code_lane_m<MODE>(o1.code, a.codes_kind, ovf);
          int k2 = code_lane_m<MODE>(o2.code, a.codes_kind, ovf), k3 = code_lane_m<MODE>(o3.code, a.codes_kind, ovf);
          isum += k0 + k1 + k2 + k3;
          uint32_t w = (uint32_t)(k0 & 0xff) | ((uint32_t)(k1 & 0xff) << 8) | ((uint32_t)(k2 & 0xff) << 16) |
                       ((uint32_t)(k3 & 0xff) << 24);
          *reinterpret_cast<uint32_t*>(c8 + c) = w;
        } else if (cb) {
          __nv_bfloat16 h0 = __float2bfloat16_rn(o0.code), h1 = __float2bfloat16_rn(o1.code);
          __nv_bfloat16 h2 = __float2bfloat16_rn(o2.code), h3 = __float2bfloat16_rn(o3.code);
          __nv_bfloat162 p0(h0, h1), p1(h2, h3);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&p0);
          u.y = *reinterpret_cast<uint32_t*>(&p1);
          if (f16) {
            __half2 q0 = __floats2half2_rn(o0.code, o1.code), q1 = __floats2half2_rn(o2.code, o3.code);
            u.x = *reinterpret_cast<uint32_t*>(&q0);
            u.y = *reinterpret_cast<uint32_t*>(&q1);
          }
          *reinterpret_cast<uint2*>(cb + c) = u;
          if (cb_lo) {
            __nv_bfloat162 q0(__float2bfloat16_rn(o0.code - __bfloat162float(h0)),
                              __float2bfloat16_rn(o1.code - __bfloat162float(h1)));
            __nv_bfloat162 q1(__float2bfloat16_rn(o2.code - __bfloat162float(h2)),
                              __float2bfloat16_rn(o3.code - __bfloat162float(h3)));
            uint2 l;
            l.x = *reinterpret_cast<uint32_t*>(&q0);
            l.y = *reinterpret_cast<uint32_t*>(&q1);
            *reinterpret_cast<uint2*>(cb_lo + c) = l;
            if (cb_lo2) {
              __nv_bfloat162 r0(__float2bfloat16_rn(o0.code - __bfloat162float(h0) - __bfloat162float(q0.x)),
                                __float2bfloat16_rn(o1.code - __bfloat162float(h1) - __bfloat162float(q0.y)));
              __nv_bfloat162 r1(__float2bfloat16_rn(o2.code - __bfloat162float(h2) - __bfloat162float(q1.x)),
                                __float2bfloat16_rn(o3.code - __bfloat162float(h3) - __bfloat162float(q1.y)));
              uint2 l2;
              l2.x = *reinterpret_cast<uint32_t*>(&r0);
              l2.y = *reinterpret_cast<uint32_t*>(&r1);
              *reinterpret_cast<uint2*>(cb_lo2 + c) = l2;
            }
          }
          if (a.row_sum) isum += (int)o0.code + (int)o1.code + (int)o2.code + (int)o3.code;
        }
      }
      if (c4) {  // 4 nibbles per lane; lane pairs merge into one 32-bit store (8 codes)
        uint32_t h = 0;
        if (valid) {
          int k0, k1, k2, k3;
          h = code_f4_m<MODE>(o0.code, k0, ovf) | (code_f4_m<MODE>(o1.code, k1, ovf) << 4) | (code_f4_m<MODE>(o2.code, k2, ovf) << 8) |
              (code_f4_m<MODE>(o3.code, k3, ovf) << 12);
          isum += k0 + k1 + k2 + k3;
        }
        const uint32_t hi = __shfl_down_sync(0xffffffffu, h, 1);
        if (valid && !(lane & 1)) *reinterpret_cast<uint32_t*>(c4 + (c >> 1)) = h | (hi << 16);
      }
      if (br) {  // 4 bits per lane -> 4 words per 128-column step
        uint32_t nib = 0;
        if (valid) nib = (uint32_t)(o0.code > 0.f) | ((uint32_t)(o1.code > 0.f) << 1) |
                         ((uint32_t)(o2.code > 0.f) << 2) | ((uint32_t)(o3.code > 0.f) << 3);
        uint32_t w = nib << (4 * (lane & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        const int64_t widx = (base >> 5) + (lane >> 3);
        if ((lane & 7) == 0 && widx * 32 < c1) br[widx] = w;
      }
     }
     if (MODE == QT_Q_XNOR_ROW && xnor_one_pass) xsum += (double)xpart;
    }
  } else {
    for (int64_t base = c0; base < c1; base += 32) {
      const int64_t c = base + lane;
      const bool valid = c < c1;
      float v = valid ? pre_col<PRE>(a.q, __ldg(xr + c), c) : 0.f;
      if (MODE == QT_Q_XNOR_ROW && xnor_one_pass && valid) xsum += (double)v;
      QOut o = quant_elem<MODE>(a.q, v, row_mean);
      if (valid) {
        if (yr) yr[c] = o.y;
        if (c8) {
          int k = code_lane_m<MODE>(o.code, a.codes_kind, ovf);
          isum += k;
          c8[c] = (int8_t)k;
        } else if (cb) {
          __nv_bfloat16 h = __float2bfloat16_rn(o.code);
          if (f16) {
            __half hh = __float2half_rn(o.code);
            h = *reinterpret_cast<__nv_bfloat16*>(&hh);
          }
          cb[c] = h;
          if (cb_lo) {
            __nv_bfloat16 m = __float2bfloat16_rn(o.code - __bfloat162float(h));
            cb_lo[c] = m;
            if (cb_lo2) cb_lo2[c] = __float2bfloat16_rn(o.code - __bfloat162float(h) - __bfloat162float(m));
          }
          if (a.row_sum) isum += (int)o.code;
        }
      }
      if (c4) {
        uint32_t nib = 0;
        if (valid) {
          int k;
          nib = code_f4_m<MODE>(o.code, k, ovf);
          isum += k;
        }
        const uint32_t hi = __shfl_down_sync(0xffffffffu, nib, 1);
        if (valid && !(lane & 1)) c4[c >> 1] = (uint8_t)(nib | (hi << 4));
      }
      if (br) {
        uint32_t w = __ballot_sync(0xffffffffu, valid && o.code > 0.f);
        if (lane == 0) br[base >> 5] = w;
      }
    }
  }

  if (MODE == QT_Q_XNOR_ROW && xnor_one_pass) {
    xsum = warp_sum_d(xsum);
    if (a.row_scale && lane == 0) {
      if (a.nchunks == 1) a.row_scale[row] = (float)(xsum / (double)a.cols);
      else a.row_scale[(int64_t)chunk_id * a.rows + row] = (float)xsum;     // partial sums: the consumer adds them in order
    }
  }
  // zero-fill the padding columns / words (last chunk only)
  if (chunk_id == a.nchunks - 1) {
    if (c8) for (int64_t c = a.cols + lane; c < a.ld_codes; c += 32) c8[c] = 0;
    if (cb) for (int64_t c = a.cols + lane; c < a.ld_codes; c += 32) {
      cb[c] = __float2bfloat16_rn(0.f);
      if (cb_lo) cb_lo[c] = __float2bfloat16_rn(0.f);
      if (cb_lo2) cb_lo2[c] = __float2bfloat16_rn(0.f);
    }
    if (c4) for (int64_t b = ((a.cols + 1) >> 1) + lane; b < (a.ld_codes >> 1); b += 32) c4[b] = 0;
    if (br) for (int64_t w = ((a.cols + 31) >> 5) + lane; w < a.ld_bits; w += 32) br[w] = 0u;
  }
  if (a.row_sum) {
    isum = warp_sum_i(isum);
    if (lane == 0) {
      if (a.nchunks == 1) a.row_sum[row] = isum;
      else atomicAdd(a.row_sum + row, isum);
    }
  }
  if (a.overflow) {
    unsigned any = __ballot_sync(0xffffffffu, ovf);
    if (any && lane == 0) atomicOr(a.overflow, 1);
  }
}

// Channels-last code emitter for NCHW activations (conv inputs).  One CTA = one image x 128 channels x 32 pixels:
// coalesced fp32 reads along the pixel axis (one warp = 32 consecutive pixels of one channel), the fp32 fake-quant
// result is written with the same pattern, codes are transposed through a 4 KB shared tile and leave as 128-byte
// channel runs per pixel (16 bytes per thread).
struct NhwcArgs {
  QParams q;
  const float* x;
  float* y;
  int8_t* codes;
  int codes_kind;
  int64_t B, C, HW;
  int32_t* overflow;
};

template <int MODE>
__global__ void __launch_bounds__(256) act_quant_nhwc_kernel(NhwcArgs a) {
  // tile[pixel][channel] as 32-bit words of 4 channel codes; 33-word rows make both the transposing stores
  // (lane = pixel, fixed channel word) and the row reads of the write phase bank-conflict free
  __shared__ uint32_t tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 128, b = blockIdx.z;
  const float* xb = a.x + b * a.C * a.HW;
  float* yb = a.y ? a.y + b * a.C * a.HW : nullptr;
  bool ovf = false;
  const int64_t p = p0 + lane;
  const bool pix_ok = p < a.HW;
  // 16 channels per warp: all 16 loads are issued before any is used (64 B in flight per lane)
  float xv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int64_t c = c0 + warp * 16 + i;
    xv[i] = (pix_ok && c < a.C) ? __ldcs(xb + c * a.HW + p) : 0.f;
  }
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = q4 * 4 + j;
      const int64_t c = c0 + warp * 16 + i;
      int code = 0;
      if (pix_ok && c < a.C) {
        float v = xv[i];
        if (a.q.pre_scale) v = pre_apply(a.q, v, c % a.q.pre_channels);
        QOut o = quant_elem<MODE>(a.q, v, 0.f);
        if (yb) __stcs(yb + c * a.HW + p, o.y);
        code = code_lane_m<MODE>(o.code, a.codes_kind, ovf);
      }
      word |= (uint32_t)(code & 0xff) << (8 * j);
    }
    tile[lane][warp * 4 + q4] = word;
  }
  __syncthreads();
  // write phase: thread -> (pixel = tid / 8, 16 channels = 4 words starting at (tid % 8) * 4): 128-byte runs per pixel
  const int px = threadIdx.x >> 3, wg = (threadIdx.x & 7) * 4;
  const int64_t pp = p0 + px;
  if (pp < a.HW && c0 + wg * 4 < a.C) {
    uint4 v = make_uint4(tile[px][wg], tile[px][wg + 1], tile[px][wg + 2], tile[px][wg + 3]);
    int8_t* dst = a.codes + (b * a.HW + pp) * a.C + c0 + wg * 4;
    if ((a.C & 15) == 0) {
      *reinterpret_cast<uint4*>(dst) = v;
    } else {
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
      for (int j = 0; j < 16 && c0 + wg * 4 + j < a.C; ++j) dst[j] = (int8_t)((w4[j >> 2] >> (8 * (j & 3))) & 0xff);
    }
  }
  if (a.overflow) {
    unsigned any = __ballot_sync(0xffffffffu, ovf);
    if (any && lane == 0) atomicOr(a.overflow, 1);
  }
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---------------------------------------------------------------------------------------------
// weight-side reductions
// ---------------------------------------------------------------------------------------------
// stats[0] = max|tanh w|, stats[1] = sum|w| (then turned into the mean), stats[2] = #zeros, stats[3] = max|w|
__global__ void __launch_bounds__(256) weight_stats_kernel(const float* __restrict__ w, int64_t n, int64_t k, int64_t ld,
                                                           float* stats, double* dsum) {
  double s = 0.0;
  float mx = 0.f;
  int zeros = 0;
  const int64_t total = n * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / k, c = i - r * k;
    float v = __ldg(w + r * ld + c);
    float av = fabsf(v);
    s += (double)av;
    mx = fmaxf(mx, av);
    zeros += (v == 0.f);
  }
  // tanh is monotone, so max|tanh w| = |tanh(max|w|)|; evaluated once, in double, below.
  s = warp_sum_d(s);
  zeros = warp_sum_i(zeros);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(dsum, s);
    atomicAdd(reinterpret_cast<int*>(stats + 8), zeros);
    atomicMax(reinterpret_cast<int*>(stats + 3), __float_as_int(mx));  // non-negative floats order as ints
  }
}

__global__ void weight_stats_finish_kernel(float* stats, const double* dsum, double count) {
  float mx = stats[3];
  stats[0] = (float)tanh((double)mx);
  stats[1] = (float)(*dsum / count);
  stats[2] = (float)(*reinterpret_cast<int*>(stats + 8));
}

// alpha[k] = mean over rows of |w[:, k]|  (xnor_connect.py:111, DIM = 0)
__global__ void __launch_bounds__(256) col_absmean_kernel(const float* __restrict__ w, int64_t n, int64_t k, int64_t ld,
                                                          float* __restrict__ alpha) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  double s = 0.0;
  if (c < k)
    for (int64_t r = ty; r < n; r += 8) s += (double)fabsf(__ldg(w + r * ld + c));
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    alpha[c] = (float)(t / (double)n);
  }
}

struct PackArgs {
  int mode, bit_width, lane_bits;
  const float* w;
  int64_t n, k, ld_w;
  uint8_t* packed;
  int64_t ld_packed;  // bytes
  const float* stats;
  float* wq;
  const float* alpha;
};

// DoReFa weight code, dorefa_connect.py:108-110: c = round(n * (tanh(w) / (2 max|tanh w|) + 0.5))
__device__ __forceinline__ float dorefa_wcode(float w, float two_maxt, float n) {
  float t = (float)tanh((double)w);
  t = t / two_maxt + 0.5f;
  return rintf(n * t);
}

// One warp per row; lanes stride over 32-column groups so that every lane owns whole output words.
__global__ void __launch_bounds__(256) weight_pack_kernel(PackArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.n) return;
  const float* wr = a.w + row * a.ld_w;
  float* qr = a.wq ? a.wq + row * a.ld_w : nullptr;
  uint8_t* pr = a.packed + row * a.ld_packed;
  const int64_t words = a.ld_packed / 4;

  if (a.mode == QT_W_SIGN || (a.mode == QT_W_DOREFA && a.bit_width == 1)) {
    const float e = (a.mode == QT_W_DOREFA) ? a.stats[1] : 1.f;
    uint32_t* out = reinterpret_cast<uint32_t*>(pr);
    for (int64_t base = 0; base < words * 32; base += 32) {
      int64_t c = base + lane;
      bool valid = c < a.k;
      float v = valid ? __ldg(wr + c) : -1.f;
      bool pos = !(v < 0.f);
      if (valid && qr) qr[c] = (pos ? 1.f : -1.f) * e;
      uint32_t wbits = __ballot_sync(0xffffffffu, valid && pos);
      if (lane == 0) out[base >> 5] = wbits;
    }
  } else if (a.mode == QT_W_TERNARY || a.mode == QT_W_XNOR) {
    uint32_t* nz = reinterpret_cast<uint32_t*>(pr);
    uint32_t* sg = reinterpret_cast<uint32_t*>(pr + a.n * a.ld_packed);
    for (int64_t base = 0; base < words * 32; base += 32) {
      int64_t c = base + lane;
      bool valid = c < a.k;
      float v = valid ? __ldg(wr + c) : 0.f;
      float t;
      if (a.mode == QT_W_TERNARY) {
        float s = sign_safe(v);
        t = (s + sign_safe(v - 0.5f * s)) * 0.5f;
        if (valid && qr) qr[c] = t;
      } else {
        t = sign3(v);
        if (valid && qr) qr[c] = t * __ldg(a.alpha + c);
      }
      uint32_t bnz = __ballot_sync(0xffffffffu, valid && t != 0.f);
      uint32_t bsg = __ballot_sync(0xffffffffu, valid && t > 0.f);
      if (lane == 0) {
        nz[base >> 5] = bnz;
        sg[base >> 5] = bsg;
      }
    }
  } else {  // DoReFa k >= 2
    const float n = (float)((1 << a.bit_width) - 1);
    const float two_maxt = 2.f * a.stats[0];
    const float inv_n = 1.0f / n;
    const int per_word = 32 / a.lane_bits;
    uint32_t* out = reinterpret_cast<uint32_t*>(pr);
    for (int64_t wi = lane; wi < words; wi += 32) {
      uint32_t acc = 0;
      for (int j = 0; j < per_word; ++j) {
        int64_t c = wi * per_word + j;
        if (c < a.k) {
          float code = dorefa_wcode(__ldg(wr + c), two_maxt, n);
          if (qr) qr[c] = 2.f * (inv_n * code) - 1.f;
          acc |= ((uint32_t)code & ((1u << a.lane_bits) - 1u)) << (j * a.lane_bits);
        }
      }
      out[wi] = acc;
    }
  }
}

struct ExpandArgs {
  int mode, bit_width, lane_bits;
  const uint8_t* packed;
  int64_t n, k, ld_packed;
  const float* alpha;
  void* out;
  int out_kind;
  int64_t ld_out;
};

// One thread per 16-byte output vector (CPV = 8 columns of a 16-bit operand, 16 of an 8-bit one, 32 e2m1 nibbles), so that
// consecutive threads write consecutive 16-byte vectors (every store instruction of a warp covers 512 contiguous bytes).
// The packed bits / codes of those columns are 1..4 words fetched once (neighbouring threads share them: broadcast loads).
// packed words one output vector needs (fetched up front for several vectors per thread: the kernel is latency-bound)
struct ExpandWords { uint32_t w[4]; };

template <int CPV>
__device__ __forceinline__ ExpandWords expand_fetch(const ExpandArgs& a, int64_t row, int64_t c0, bool one_bit) {
  ExpandWords e;
  e.w[0] = e.w[1] = e.w[2] = e.w[3] = 0u;
  if (c0 >= a.k) return e;
  const uint8_t* pr = a.packed + row * a.ld_packed;
  if (one_bit) {
    e.w[0] = __ldg(reinterpret_cast<const uint32_t*>(pr) + (c0 >> 5));
  } else if (a.mode == QT_W_TERNARY || a.mode == QT_W_XNOR) {
    e.w[0] = __ldg(reinterpret_cast<const uint32_t*>(pr) + (c0 >> 5));
    e.w[1] = __ldg(reinterpret_cast<const uint32_t*>(pr + a.n * a.ld_packed) + (c0 >> 5));
  } else {
    const int lb = a.lane_bits;                      // 2, 4 or 8: CPV columns span at most 4 words (CPV * lb <= 128 bits)
    const int64_t bit0 = c0 * lb;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(pr) + (bit0 >> 5);
    const int nwords = (int)(((bit0 & 31) + (int64_t)CPV * lb + 31) >> 5);
    const int64_t words_left = (a.ld_packed >> 2) - (bit0 >> 5);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nwords && i < words_left) e.w[i] = __ldg(wp + i);
  }
  return e;
}

// bit i of the low byte of b -> bit 4 i (one bit per nibble)
__device__ __forceinline__ uint32_t spread8_nibbles(uint32_t b) {
  uint32_t x = b & 0xFFu;
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}
// bit i of the low nibble of b -> bit 8 i (one bit per byte)
__device__ __forceinline__ uint32_t spread4_bytes(uint32_t b) { return ((b & 0xFu) * 0x00204081u) & 0x01010101u; }

// Sign / ternary / XnorNet weights straight from their bit planes with integer operations (no per-element float work):
//   e2m1 nibble  +1 = 0x2, -1 = 0xA;   int8  +1 = 0x01, -1 = 0xFF;   fp16  +-alpha[k] = half(alpha[k]) ^ (neg << 15).
// Returns false when the (mode, out_kind) pair needs the generic path.
template <int CPV>
__device__ __forceinline__ bool expand_emit_bits(const ExpandArgs& a, int64_t row, int64_t c0, bool one_bit, const ExpandWords& e) {
  const bool planes2 = a.mode == QT_W_TERNARY || a.mode == QT_W_XNOR;
  if (!one_bit && !planes2) return false;
  const int sh = (int)(c0 & 31);
  const int64_t left = a.k - c0;                                   // valid columns from c0 on (may be <= 0 or > CPV)
  const uint32_t vm = left <= 0 ? 0u : (left >= 32 ? 0xFFFFFFFFu : ((1u << (int)left) - 1u));
  // nz: element is non-zero; neg: element is negative (bit planes: 1 <-> +1 in the sign plane)
  const uint32_t nz = (one_bit ? 0xFFFFFFFFu : (e.w[0] >> sh)) & vm;
  const uint32_t neg = ~((one_bit ? e.w[0] : e.w[1]) >> sh) & nz;
  if (CPV == 32 && a.out_kind == 7) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = (spread8_nibbles(nz >> (8 * q)) << 1) | (spread8_nibbles(neg >> (8 * q)) << 3);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out) + row * (a.ld_out >> 1) + (c0 >> 1)) = make_uint4(w[0], w[1], w[2], w[3]);
    return true;
  }
  if (CPV == 16 && a.out_kind == 1) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = spread4_bytes(nz >> (4 * q)) | (spread4_bytes(neg >> (4 * q)) * 0xFEu);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out) + row * a.ld_out + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    return true;
  }
  if (CPV == 8 && (a.out_kind == 5 || a.out_kind == 6)) {
    // fp16: kind 5 = alpha[k] * sign (XnorNet), kind 6 = exact +-1 / 0
    uint32_t h[4];
    if (a.out_kind == 5) {
      float al[8];
      if (left >= 8 && (c0 & 3) == 0) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(a.alpha + c0)), a1 = __ldg(reinterpret_cast<const float4*>(a.alpha + c0 + 4));
        al[0] = a0.x; al[1] = a0.y; al[2] = a0.z; al[3] = a0.w; al[4] = a1.x; al[5] = a1.y; al[6] = a1.z; al[7] = a1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) al[j] = (j < left) ? __ldg(a.alpha + c0 + j) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        __half2 hh = __floats2half2_rn(al[2 * q], al[2 * q + 1]);
        h[q] = *reinterpret_cast<uint32_t*>(&hh);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = 0x3C003C00u;      // half(1.0) twice
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t n2 = (nz >> (2 * q)) & 3u, g2 = (neg >> (2 * q)) & 3u;
      const uint32_t keep = ((n2 & 1u) ? 0x0000FFFFu : 0u) | ((n2 & 2u) ? 0xFFFF0000u : 0u);
      const uint32_t flip = ((g2 & 1u) << 15) | ((g2 & 2u) << 30);
      h[q] = (h[q] ^ flip) & keep;
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + row * a.ld_out + c0) = make_uint4(h[0], h[1], h[2], h[3]);
    return true;
  }
  return false;
}

template <int CPV>
__device__ __forceinline__ void expand_emit(const ExpandArgs& a, int64_t row, int64_t c0, bool one_bit, const ExpandWords& e) {
  if (expand_emit_bits<CPV>(a, row, c0, one_bit, e)) return;
  float v[CPV];
  if (c0 >= a.k) {
#pragma unroll
    for (int j = 0; j < CPV; ++j) v[j] = 0.f;
  } else if (one_bit) {
    const uint32_t w = e.w[0] >> (c0 & 31);
#pragma unroll
    for (int j = 0; j < CPV; ++j) v[j] = (c0 + j < a.k) ? (((w >> j) & 1u) ? 1.f : -1.f) : 0.f;
  } else if (a.mode == QT_W_TERNARY || a.mode == QT_W_XNOR) {
    const uint32_t nz = e.w[0] >> (c0 & 31);
    const uint32_t sg = e.w[1] >> (c0 & 31);
#pragma unroll
    for (int j = 0; j < CPV; ++j) v[j] = (c0 + j < a.k && ((nz >> j) & 1u)) ? (((sg >> j) & 1u) ? 1.f : -1.f) : 0.f;
  } else {
    const int lb = a.lane_bits;
    const int64_t bit0 = c0 * lb;
    const uint32_t mask = (1u << lb) - 1u;
    const float n = (float)((1 << a.bit_width) - 1);
#pragma unroll
    for (int j = 0; j < CPV; ++j) {
      const int bit = (int)(bit0 & 31) + j * lb;
      const int wi = bit >> 5;                 // selects, not an indexed array: the words stay in registers
      const uint32_t word = (wi == 0) ? e.w[0] : ((wi == 1) ? e.w[1] : ((wi == 2) ? e.w[2] : e.w[3]));
      const uint32_t code = (word >> (bit & 31)) & mask;
      float x = (a.out_kind == 2) ? (float)code : 2.f * (float)code - n;
      v[j] = (c0 + j < a.k) ? x : 0.f;
    }
  }
  if (CPV == 32) {          // e2m1 nibbles: 32 columns -> 16 bytes
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t acc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc |= f4_nibble(v[(8 * q + j) % CPV]) << (4 * j);
      w[q] = acc;
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out) + row * (a.ld_out >> 1) + (c0 >> 1)) = make_uint4(w[0], w[1], w[2], w[3]);
  } else if (CPV == 16) {   // int8 / uint8 codes
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      w[q] = ((uint32_t)((int)v[(4 * q) % CPV] & 0xff)) | ((uint32_t)((int)v[(4 * q + 1) % CPV] & 0xff) << 8) |
             ((uint32_t)((int)v[(4 * q + 2) % CPV] & 0xff) << 16) | ((uint32_t)((int)v[(4 * q + 3) % CPV] & 0xff) << 24);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out) + row * a.ld_out + c0) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {                  // 16-bit planes
    if (a.out_kind == 4 || a.out_kind == 5) {     // XnorNet: alpha[k] * sign
      float al[8];
      if (c0 + 8 <= a.k && (c0 & 3) == 0) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(a.alpha + c0)), a1 = __ldg(reinterpret_cast<const float4*>(a.alpha + c0 + 4));
        al[0] = a0.x; al[1] = a0.y; al[2] = a0.z; al[3] = a0.w; al[4] = a1.x; al[5] = a1.y; al[6] = a1.z; al[7] = a1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) al[j] = (c0 + j < a.k) ? __ldg(a.alpha + c0 + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j % CPV] *= al[j];
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x0 = v[(2 * q) % CPV], x1 = v[(2 * q + 1) % CPV];
      if (a.out_kind >= 5) {
        __half2 hh = __floats2half2_rn(x0, x1);
        h[q] = *reinterpret_cast<uint32_t*>(&hh);
        l[q] = 0;
      } else {
        const __nv_bfloat16 b0 = __float2bfloat16_rn(x0), b1 = __float2bfloat16_rn(x1);
        __nv_bfloat162 hh(b0, b1);
        __nv_bfloat162 ll(__float2bfloat16_rn(x0 - __bfloat162float(b0)), __float2bfloat16_rn(x1 - __bfloat162float(b1)));
        h[q] = *reinterpret_cast<uint32_t*>(&hh);
        l[q] = *reinterpret_cast<uint32_t*>(&ll);
      }
    }
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + row * a.ld_out + c0;
    *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (a.out_kind == 4) *reinterpret_cast<uint4*>(o + a.n * a.ld_out) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// One thread per FOUR 16-byte output vectors (CPV = 8 columns of a 16-bit operand, 16 of an 8-bit one, 32 e2m1 nibbles), 256
// vectors apart, so that every store instruction of a warp covers 512 contiguous bytes and the packed words of all four are
// in flight before the first is unpacked (with one vector per thread the kernel ran 7 latency-bound waves).
template <int CPV>
__global__ void __launch_bounds__(256) weight_expand_kernel(ExpandArgs a) {
  constexpr int U = 4;
  // 32-bit index arithmetic (the launcher checks n * vec_per_row < 2^31): a 64-bit division per vector costs more than
  // the unpacking itself
  const uint32_t vec_per_row = (uint32_t)(a.ld_out / CPV);
  const uint32_t total = (uint32_t)a.n * vec_per_row;
  const uint32_t base = blockIdx.x * (256u * U) + threadIdx.x;
  const bool one_bit = a.mode == QT_W_SIGN || (a.mode == QT_W_DOREFA && a.bit_width == 1);
  uint32_t row[U], c0[U];
  ExpandWords e[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t gid = base + u * 256u;
    row[u] = gid / vec_per_row;
    c0[u] = (gid - row[u] * vec_per_row) * CPV;
    if (gid < total) e[u] = expand_fetch<CPV>(a, row[u], c0[u], one_bit);
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (base + u * 256u < total) expand_emit<CPV>(a, row[u], c0[u], one_bit, e[u]);
}

// ---------------------------------------------------------------------------------------------
// im2col
// ---------------------------------------------------------------------------------------------
struct Im2colArgs {
  const uint8_t* x;
  int eb;
  int64_t B, C, H, W, OH, OW;
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int64_t c_begin, cg;   // first channel of the group, channels per group
  uint8_t* out;
  int64_t ld_out, kcols;
  int nhwc;
};

// NHWC fast path: every 16-byte vector of an output row is 16 bytes of one pixel's channel run (or zeros).
__global__ void __launch_bounds__(256) im2col_nhwc_vec_kernel(Im2colArgs a) {
  const int64_t vec_per_row = (a.ld_out * a.eb) / 16;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t M = a.B * a.OH * a.OW;
  if (gid >= M * vec_per_row) return;
  const int64_t m = gid / vec_per_row, v = gid - m * vec_per_row;
  const int64_t col0 = v * (16 / a.eb);
  uint4 val = make_uint4(0, 0, 0, 0);
  if (col0 < a.kcols) {
    const int64_t b = m / (a.OH * a.OW), r = m - b * (a.OH * a.OW);
    const int64_t oh = r / a.OW, ow = r - oh * a.OW;
    const int tap = (int)(col0 / a.cg);
    const int64_t c = col0 - (int64_t)tap * a.cg;
    const int ky = tap / a.kw, kx = tap - ky * a.kw;
    const int64_t ih = oh * a.sh - a.ph + (int64_t)ky * a.dh, iw = ow * a.sw - a.pw + (int64_t)kx * a.dw;
    if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
      val = __ldg(reinterpret_cast<const uint4*>(a.x + (((b * a.H + ih) * a.W + iw) * a.C + a.c_begin + c) * a.eb));
  }
  *reinterpret_cast<uint4*>(a.out + (m * a.ld_out) * a.eb + v * 16) = val;
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) im2col_kernel(Im2colArgs a) {
  const int64_t vec_per_row = a.ld_out / VEC;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t M = a.B * a.OH * a.OW;
  if (gid >= M * vec_per_row) return;
  const int64_t m = gid / vec_per_row, col0 = (gid - m * vec_per_row) * VEC;
  const int64_t b = m / (a.OH * a.OW), r = m - b * (a.OH * a.OW);
  const int64_t oh = r / a.OW, ow = r - oh * a.OW;
  const T* x = reinterpret_cast<const T*>(a.x);
  __align__(16) T v[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    int64_t col = col0 + j;
    T val = T(0);
    if (col < a.kcols) {
      int tap = (int)(col / a.cg);
      int64_t c = col - (int64_t)tap * a.cg;           // column = (kh, kw, c): channel fastest
      int ky = tap / a.kw, kx = tap - ky * a.kw;
      int64_t ih = oh * a.sh - a.ph + (int64_t)ky * a.dh;
      int64_t iw = ow * a.sw - a.pw + (int64_t)kx * a.dw;
      if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
        val = a.nhwc ? x[((b * a.H + ih) * a.W + iw) * a.C + a.c_begin + c]
                     : x[((b * a.C + a.c_begin + c) * a.H + ih) * a.W + iw];
    }
    v[j] = val;
  }
  T* o = reinterpret_cast<T*>(a.out) + m * a.ld_out + col0;
  if (sizeof(T) * VEC == 16) {
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(v);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[j] = v[j];
  }
}

// First-layer gather: fp32 NCHW image -> three bf16 planes (hi/mid/lo) of the im2col matrix in one pass.
// The column -> (input offset, dy, dx) decomposition is tabulated once per CTA in shared memory, so the inner loop is
// table lookups + L1-resident gathers; each thread emits 8 consecutive columns (one 16-byte store per plane).
template <int NPLANES>
__global__ void __launch_bounds__(256) im2col_split3_kernel(Im2colArgs a, int64_t plane_stride) {
  extern __shared__ int tab[];                 // [kcols] offset, [kcols] (dy << 16 | dx)
  int* toff = tab;
  int* tdyx = tab + a.kcols;
  for (int col = threadIdx.x; col < a.kcols; col += blockDim.x) {
    int tap = col / (int)a.cg, c = col - tap * (int)a.cg;
    int ky = tap / a.kw, kx = tap - ky * a.kw;
    int dy = ky * a.dh, dx = kx * a.dw;
    toff[col] = (int)(((a.c_begin + c) * a.H + dy) * a.W + dx);
    tdyx[col] = (dy << 16) | dx;
  }
  __syncthreads();
  const int64_t vec_per_row = a.ld_out / 8;
  const int64_t M = a.B * a.OH * a.OW;
  for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < M * vec_per_row;
       gid += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = gid / vec_per_row;
    const int col0 = (int)(gid - m * vec_per_row) * 8;
    const int64_t b = m / (a.OH * a.OW), r = m - b * (a.OH * a.OW);
    const int oh = (int)(r / a.OW), ow = (int)(r - (int64_t)oh * a.OW);
    const int ih0 = oh * a.sh - a.ph, iw0 = ow * a.sw - a.pw;
    const float* xb = reinterpret_cast<const float*>(a.x) + b * a.C * a.H * a.W + (int64_t)ih0 * a.W + iw0;
    __align__(16) __nv_bfloat16 h[8], mi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + j;
      float v = 0.f;
      if (col < a.kcols) {
        const int dyx = tdyx[col];
        const int ih = ih0 + (dyx >> 16), iw = iw0 + (dyx & 0xffff);
        if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) v = __ldg(xb + toff[col]);
      }
      h[j] = __float2bfloat16_rn(v);
      float r1 = v - __bfloat162float(h[j]);
      mi[j] = __float2bfloat16_rn(r1);
      lo[j] = __float2bfloat16_rn(r1 - __bfloat162float(mi[j]));
    }
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + m * a.ld_out + col0;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(o + plane_stride) = *reinterpret_cast<uint4*>(mi);
    if (NPLANES == 3) *reinterpret_cast<uint4*>(o + 2 * plane_stride) = *reinterpret_cast<uint4*>(lo);
  }
}


// ---------------------------------------------------------------------------------------------
// Backward (STE) helpers
// ---------------------------------------------------------------------------------------------
// fp32 [rows, cols] -> bf16 planes of the TRANSPOSE, out[p][c][r] (plane stride = cols * ld_out), p = 0 hi, 1 lo (, 2 lo2).
// Feeds the tcgen05 bf16 GEMM for the gradient contractions, whose reduction runs over the batch (grad_W = g^T x) or over the
// output features (grad_x = g W_q): the operand that is contracted along its leading dimension has to be K-major.
// 64 x 64 tile through shared memory: coalesced fp32 reads, 128-byte bf16 rows on the way out.
template <int NPLANES>
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_x,
                                                              __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  __shared__ float tile[64][65];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int r = idx >> 6, c = idx & 63;
    float v = 0.f;
    if (r0 + r < rows && c0 + c < cols) v = __ldg(x + (r0 + r) * ld_x + c0 + c);
    tile[r][c] = v;
  }
  __syncthreads();
  const int64_t plane = cols * ld_out;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int c = idx >> 6, r = idx & 63;
    if (c0 + c < cols && r0 + r < ld_out) {          // columns rows..ld_out-1 of the transposed rows are zero-filled
      const float v = tile[r][c];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const int64_t o = (c0 + c) * ld_out + r0 + r;
      out[o] = h;
      if (NPLANES >= 2) {
        const float r1 = v - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        out[plane + o] = m;
        if (NPLANES >= 3) out[2 * plane + o] = __float2bfloat16_rn(r1 - __bfloat162float(m));
      }
    }
  }
}

// out = (|x| <= thresh) ? g : 0   -- the clip-mask straight-through estimator of BinaryConnect / TernaryConnect
// (binary_connect.py:30-38, terner_connect.py:29-34: g.clone(); g[abs(x) > 1.001] = 0), one pass instead of four.
__global__ void __launch_bounds__(256) ste_clip_kernel(const float* __restrict__ g, const float* __restrict__ x, float thresh,
                                                       float* __restrict__ out, int64_t n) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 4 <= n && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g + i4)), xv = __ldcs(reinterpret_cast<const float4*>(x + i4));
    float4 o;
    o.x = (fabsf(xv.x) > thresh) ? 0.f : gv.x; o.y = (fabsf(xv.y) > thresh) ? 0.f : gv.y;
    o.z = (fabsf(xv.z) > thresh) ? 0.f : gv.z; o.w = (fabsf(xv.w) > thresh) ? 0.f : gv.w;
    *reinterpret_cast<float4*>(out + i4) = o;
  } else {
    for (int64_t i = i4; i < n && i < i4 + 4; ++i) out[i] = (fabsf(x[i]) > thresh) ? 0.f : g[i];
  }
}


// LogLin weights: int8 codes in HBM -> bf16 operand.
//   lin: value = code * 2^(fsr - bw)            -> the operand is the integer code itself (exact), the step goes to the epilogue
//   log: value = sign(code) * 2^(emin + |code| - 1), code 0 <-> 0   -> the operand is the power of two (exact in bf16)
__global__ void __launch_bounds__(256) loglin_expand_kernel(const int8_t* __restrict__ codes, int64_t n, int64_t k, int64_t ld_codes,
                                                            int is_log, int emin, __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  const int64_t vec_per_row = ld_out / 8;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * vec_per_row) return;
  const int64_t row = gid / vec_per_row, c0 = (gid - row * vec_per_row) * 8;
  __align__(16) __nv_bfloat16 h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = 0.f;
    if (c0 + j < k) {
      const int c = codes[row * ld_codes + c0 + j];
      if (!is_log) v = (float)c;
      else if (c != 0) v = ldexpf(c < 0 ? -1.f : 1.f, emin + (c < 0 ? -c : c) - 1);
    }
    h[j] = __float2bfloat16_rn(v);
  }
  *reinterpret_cast<uint4*>(out + row * ld_out + c0) = *reinterpret_cast<uint4*>(h);
}

template <bool UNSIGNED>
__global__ void __launch_bounds__(256) rowsum_i8_kernel(const uint8_t* __restrict__ a, int64_t rows, int64_t ld,
                                                        int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(a + row * ld);
  int s = 0;
  for (int64_t i = lane; i < ld / 4; i += 32) {
    uint32_t w = __ldg(p + i);
    if (UNSIGNED) s = __dp4a(w, 0x01010101u, (unsigned)s);
    else s = __dp4a((int)w, 0x01010101, s);
  }
  s = warp_sum_i(s);
  if (lane == 0) out[row] = s;
}

// S[b, h, w] = sum over the group's channels of the 8-bit codes (one thread per pixel, 16-byte loads, dp4a with ones)
template <bool UNSIGNED>
__global__ void __launch_bounds__(256) chan_sum_kernel(const uint8_t* __restrict__ x, int64_t npix, int64_t C, int64_t c0, int64_t cg,
                                                       int32_t* __restrict__ S) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const uint8_t* px = x + p * C + c0;
  int s = 0;
  if ((cg & 15) == 0 && ((C | c0) & 15) == 0) {
    for (int64_t c = 0; c < cg; c += 16) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(px + c));
      if (UNSIGNED) {
        s = __dp4a(v.x, 0x01010101u, (unsigned)s); s = __dp4a(v.y, 0x01010101u, (unsigned)s);
        s = __dp4a(v.z, 0x01010101u, (unsigned)s); s = __dp4a(v.w, 0x01010101u, (unsigned)s);
      } else {
        s = __dp4a((int)v.x, 0x01010101, s); s = __dp4a((int)v.y, 0x01010101, s);
        s = __dp4a((int)v.z, 0x01010101, s); s = __dp4a((int)v.w, 0x01010101, s);
      }
    }
  } else {
    for (int64_t c = 0; c < cg; ++c) s += UNSIGNED ? (int)px[c] : (int)(int8_t)px[c];
  }
  S[p] = s;
}

// row_sum[m] = sum over in-image filter taps of S
__global__ void __launch_bounds__(256) box_sum_kernel(const int32_t* __restrict__ S, Im2colArgs a, int32_t* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t M = a.B * a.OH * a.OW;
  if (m >= M) return;
  const int64_t b = m / (a.OH * a.OW), r = m - b * (a.OH * a.OW);
  const int64_t oh = r / a.OW, ow = r - oh * a.OW;
  int s = 0;
  for (int ky = 0; ky < a.kh; ++ky) {
    const int64_t ih = oh * a.sh - a.ph + (int64_t)ky * a.dh;
    if (ih < 0 || ih >= a.H) continue;
    for (int kx = 0; kx < a.kw; ++kx) {
      const int64_t iw = ow * a.sw - a.pw + (int64_t)kx * a.dw;
      if (iw >= 0 && iw < a.W) s += __ldg(S + (b * a.H + ih) * a.W + iw);
    }
  }
  out[m] = s;
}

}  // namespace qt

using namespace qt;

extern "C" int qt_patch_rowsum(const void* x_nhwc, int is_unsigned, const QtConvGeom* g, int32_t* chan_sum, int32_t* row_sum,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x_nhwc && g && chan_sum && row_sum, "qt_patch_rowsum: null argument");
  QT_REQUIRE(g->groups >= 1 && g->C % g->groups == 0 && g->group >= 0 && g->group < g->groups, "qt_patch_rowsum: bad groups");
  const int64_t npix = g->B * g->H * g->W, cg = g->C / g->groups;
  if (npix == 0) return QT_OK;
  if (is_unsigned) chan_sum_kernel<true><<<(unsigned)ceil_div(npix, 256), 256, 0, stream>>>((const uint8_t*)x_nhwc, npix, g->C, cg * g->group, cg, chan_sum);
  else chan_sum_kernel<false><<<(unsigned)ceil_div(npix, 256), 256, 0, stream>>>((const uint8_t*)x_nhwc, npix, g->C, cg * g->group, cg, chan_sum);
  QT_LAUNCH_CHECK();
  Im2colArgs a{};
  a.B = g->B; a.C = g->C; a.H = g->H; a.W = g->W; a.OH = g->OH; a.OW = g->OW;
  a.kh = g->kh; a.kw = g->kw; a.sh = g->stride_h; a.sw = g->stride_w; a.ph = g->pad_h; a.pw = g->pad_w;
  a.dh = g->dil_h; a.dw = g->dil_w;
  const int64_t M = g->B * g->OH * g->OW;
  if (M == 0) return QT_OK;
  box_sum_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, stream>>>(chan_sum, a, row_sum);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_quant_xnor_parts(int64_t cols, int has_y, int capacity) {
  if (has_y || cols < 2048) return 1;
  const int parts = (int)((cols + 1023) / 1024);
  return capacity >= parts ? parts : 1;
}

namespace qt {
// ---------------------------------------------------------------------------------------------
// Lean code-only quantizers (inference chains: no fp32 output, no bit plane, no pre-transform, cols % 1024 == 0).
// Same per-element semantics as quant_elem; the generic kernel above spends ~100 instructions per 16-byte load on output
// dispatch and 64-bit indexing and becomes issue-bound at ~4 TB/s, this one stays at the HBM rate (scratch/quant_probe.cu).
// One warp per 1024-column chunk of a row; 4 x 16-byte loads in flight per lane, two batches.
// ---------------------------------------------------------------------------------------------
struct LeanArgs {
  const float* x;
  void* codes;
  int32_t* row_sum;      // DoReFa: atomically accumulated (pre-zeroed by the launcher) when nchunks > 1
  float* row_scale;      // XnorNet: [nchunks, rows] partial sums, or [rows] means when nchunks == 1
  int32_t* overflow;
  uint32_t rows, nchunks;
  int64_t ld_x, ld_codes;   // elements
  float n;               // DoReFa levels
  float inv_cols;
  int max_ctas;          // 0: one CTA per 8 tasks; > 0: grid bound (persistent form)
  int32_t* ready;        // optional progress counters: ready[row / ready_rows] += 1 per finished task
  uint32_t ready_rows;
};

template <int MODE, int CK>
__global__ void __launch_bounds__(256) act_quant_lean_kernel(LeanArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t total = a.rows * a.nchunks;
  bool ovf = false;
  // one task per warp when the grid covers them all; a bounded grid (QtActQuant.max_ctas) walks the tasks with a stride
  for (uint32_t task = blockIdx.x * 8u + (threadIdx.x >> 5); task < total; task += gridDim.x * 8u) {
  const uint32_t row = task / a.nchunks, ch = task - row * a.nchunks;
  const float* xr = a.x + (int64_t)row * a.ld_x + ch * 1024u + 4u * lane;
  float psum = 0.f;
  int isum = 0;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(xr + it * 512 + u * 128));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      const uint32_t col = ch * 1024u + (uint32_t)(it * 512 + u * 128) + 4u * lane;     // first of this lane's 4 columns
      if (CK == 5 || CK == 3) {           // 16-bit float codes of {-1, 0, +1}
        float c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (MODE == QT_Q_XNOR_ROW) c[j] = (float)((xs[j] > 0.f) - (xs[j] < 0.f));
          else if (MODE == QT_Q_SIGN) c[j] = (xs[j] < 0.f) ? -1.f : 1.f;
          else { const float sg = (xs[j] < 0.f) ? -1.f : 1.f; c[j] = (sg + ((xs[j] - 0.5f * sg < 0.f) ? -1.f : 1.f)) * 0.5f; }
        }
        if (MODE == QT_Q_XNOR_ROW) psum += (xs[0] + xs[1]) + (xs[2] + xs[3]);
        uint2 o;
        if (CK == 5) {
          __half2 h0 = __floats2half2_rn(c[0], c[1]), h1 = __floats2half2_rn(c[2], c[3]);
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        } else {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(c[0], c[1]), h1 = __floats2bfloat162_rn(c[2], c[3]);
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        }
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.codes) + (int64_t)row * a.ld_codes + col) = o;
      } else {
        int k[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (MODE == QT_Q_SIGN) k[j] = (xs[j] < 0.f) ? -1 : 1;
          else if (MODE == QT_Q_TERNARY) { const float sg = (xs[j] < 0.f) ? -1.f : 1.f; k[j] = ((xs[j] < 0.f) ? -1 : 1) + ((xs[j] - 0.5f * sg < 0.f) ? -1 : 1); k[j] >>= 1; }
          else {                          // DoReFa: rint(n x), saturating to the lane with a sticky flag (NaN -> 0 + flag)
            float c = rintf(a.n * xs[j]);
            const float lo = (CK == 7) ? -4.f : ((CK == 1) ? -128.f : 0.f), hi = (CK == 7) ? 4.f : ((CK == 1) ? 127.f : 255.f);
            if (!(c >= lo && c <= hi)) { ovf = true; c = (c != c) ? 0.f : fminf(fmaxf(c, lo), hi); }
            k[j] = (int)c;
          }
          isum += k[j];
        }
        if (CK == 7) {                    // e2m1 nibbles; lane pairs merge into one 32-bit store
          uint32_t h = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = k[j] < 0 ? -k[j] : k[j];
            const uint32_t nib = (MODE == QT_Q_DOREFA) ? (((0x65420u >> (4 * m)) & 0xFu) | (k[j] < 0 ? 8u : 0u))
                                                       : ((k[j] > 0 ? 0x2u : 0u) | (k[j] < 0 ? 0xAu : 0u));
            h |= nib << (4 * j);
          }
          const uint32_t hi2 = __shfl_down_sync(0xffffffffu, h, 1);
          if (!(lane & 1u)) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(a.codes) + (((int64_t)row * a.ld_codes + col) >> 1)) = h | (hi2 << 16);
        } else {
          const uint32_t w = (uint32_t)(k[0] & 0xff) | ((uint32_t)(k[1] & 0xff) << 8) | ((uint32_t)(k[2] & 0xff) << 16) | ((uint32_t)(k[3] & 0xff) << 24);
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(a.codes) + (int64_t)row * a.ld_codes + col) = w;
        }
      }
    }
  }
  if (MODE == QT_Q_XNOR_ROW && a.row_scale) {
    double s = warp_sum_d((double)psum);
    if (lane == 0) {
      if (a.nchunks == 1) a.row_scale[row] = (float)(s / 1024.0);
      else a.row_scale[(int64_t)ch * a.rows + row] = (float)s;
    }
  }
  if (a.row_sum) {
    isum = warp_sum_i(isum);
    if (lane == 0) {
      if (a.nchunks == 1) a.row_sum[row] = isum;
      else atomicAdd(a.row_sum + row, isum);
    }
  }
  if (a.ready) {           // publish the chunk: every lane's stores (and the row vectors above) before the counter
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicAdd(a.ready + row / a.ready_rows, 1);
    }
  }
  }   // task loop
  if (MODE == QT_Q_DOREFA && a.overflow) {
    const unsigned any = __ballot_sync(0xffffffffu, ovf);
    if (any && lane == 0) atomicOr(a.overflow, 1);
  }
}

template <int MODE, int CK>
static int launch_lean(const LeanArgs& a, cudaStream_t stream) {
  const uint32_t tasks = a.rows * a.nchunks;
  uint32_t grid = (tasks + 7u) / 8u;
  if (a.max_ctas > 0 && grid > (uint32_t)a.max_ctas) grid = (uint32_t)a.max_ctas;
  act_quant_lean_kernel<MODE, CK><<<grid, 256, 0, stream>>>(a);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

// Returns 1 when the call was served by a lean kernel, 0 when the generic kernel must run, < 0 on error.
static int try_lean_quant(const QtActQuant* p, cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("QTB200_LEAN_QUANT"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
  if (!enabled) return 0;
  if (p->y || p->bits || p->pre_scale || p->nhwc_c || !p->codes) return 0;
  if (p->cols % 1024 != 0 || p->ld_x % 4 != 0 || !aligned(p->x, 16) || p->ld_codes != p->cols) return 0;
  if (p->rows * (p->cols / 1024) >= (1ll << 31) || p->cols >= (1ll << 31)) return 0;
  const int ck = p->codes_kind;
  if (!(ck == 1 || ck == 2 || ck == 3 || ck == 5 || ck == 7)) return 0;
  if ((ck == 3 || ck == 5) && (p->ld_codes % 4 != 0 || !aligned(p->codes, 8))) return 0;
  if ((ck == 1 || ck == 2) && (p->ld_codes % 4 != 0 || !aligned(p->codes, 4))) return 0;
  if (ck == 7 && (p->ld_codes % 8 != 0 || !aligned(p->codes, 4))) return 0;
  LeanArgs a;
  a.x = p->x; a.codes = p->codes; a.row_sum = p->row_sum; a.row_scale = p->row_scale; a.overflow = p->overflow;
  a.rows = (uint32_t)p->rows; a.nchunks = (uint32_t)(p->cols / 1024); a.ld_x = p->ld_x; a.ld_codes = p->ld_codes;
  a.n = 1.f; a.inv_cols = 1.f / (float)p->cols; a.max_ctas = p->max_ctas;
  a.ready = p->ready; a.ready_rows = p->ready ? (uint32_t)p->ready_rows : 1u;
  if (p->mode == QT_Q_XNOR_ROW) {
    if (!(ck == 3 || ck == 5)) return 0;
    // the partial-sum contract of qt_quant_xnor_parts: chunks of 1024 columns only when the caller provided room for them
    const int parts = qt_quant_xnor_parts(p->cols, 0, p->row_parts);
    if (p->row_scale && (uint32_t)parts != a.nchunks) return 0;
    return (ck == 5 ? launch_lean<QT_Q_XNOR_ROW, 5>(a, stream) : launch_lean<QT_Q_XNOR_ROW, 3>(a, stream)) == QT_OK ? 1 : QT_ECUDA;
  }
  if (a.nchunks > 1 && p->row_sum) {
    if (cudaMemsetAsync(p->row_sum, 0, sizeof(int32_t) * p->rows, stream) != cudaSuccess) return 0;
  }
  int rc = 0;
  if (p->mode == QT_Q_SIGN) {
    if (ck == 7) rc = launch_lean<QT_Q_SIGN, 7>(a, stream);
    else if (ck == 1) rc = launch_lean<QT_Q_SIGN, 1>(a, stream);
    else return 0;
  } else if (p->mode == QT_Q_TERNARY) {
    if (ck == 7) rc = launch_lean<QT_Q_TERNARY, 7>(a, stream);
    else if (ck == 1) rc = launch_lean<QT_Q_TERNARY, 1>(a, stream);
    else return 0;
  } else if (p->mode == QT_Q_DOREFA) {
    if (p->bit_width < 2 || p->bit_width > 8) return 0;
    a.n = (float)((1 << p->bit_width) - 1);
    if (ck == 7 && p->bit_width == 2) rc = launch_lean<QT_Q_DOREFA, 7>(a, stream);
    else if (ck == 1) rc = launch_lean<QT_Q_DOREFA, 1>(a, stream);
    else if (ck == 2) rc = launch_lean<QT_Q_DOREFA, 2>(a, stream);
    else return 0;
  } else {
    return 0;
  }
  return rc == QT_OK ? 1 : rc;
}
}  // namespace qt

extern "C" int qt_quant_act(const QtActQuant* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(p && p->x, "qt_quant_act: null argument");
  QT_REQUIRE(p->mode >= QT_Q_SIGN && p->mode <= QT_Q_SPLIT, "qt_quant_act: bad mode %d", p->mode);
  QT_REQUIRE(p->rows >= 0 && p->cols >= 0 && p->ld_x >= p->cols, "qt_quant_act: bad shape");
  if (p->rows == 0 || p->cols == 0) return QT_OK;
  ActArgs a;
  a.q.mode = p->mode;
  a.q.n = a.q.inv_n = 1.f;
  a.q.lo_e = a.q.hi_e = a.q.step = a.q.maxv = 0.f;
  a.q.with_sign = p->with_sign;
  a.q.pre_scale = p->pre_scale; a.q.pre_shift = p->pre_shift;
  a.q.pre_channels = p->pre_channels > 0 ? p->pre_channels : 1; a.q.pre_hw = p->pre_hw > 0 ? p->pre_hw : 1;
  a.q.pre_clamp = p->pre_clamp; a.q.pre_lo = p->pre_lo; a.q.pre_hi = p->pre_hi;
  if (p->pre_scale) QT_REQUIRE(p->pre_shift && p->pre_channels > 0, "qt_quant_act: pre-transform needs pre_shift and pre_channels");
  if (p->mode == QT_Q_DOREFA) {
    QT_REQUIRE(p->bit_width >= 2 && p->bit_width <= 16, "qt_quant_act: DoReFa bit width %d not in 2..16", p->bit_width);
    a.q.n = (float)((1 << p->bit_width) - 1);
    a.q.inv_n = 1.0f / a.q.n;
    if (p->codes_kind == 1 || p->codes_kind == 2)
      QT_REQUIRE(p->bit_width <= 8, "qt_quant_act: 8-bit code lanes need bit width <= 8");
  }
  if (p->mode == QT_Q_LOG) {
    a.q.lo_e = (float)(p->fsr - (1 << p->bit_width));
    a.q.hi_e = (float)p->fsr;
  }
  if (p->mode == QT_Q_LIN) {
    a.q.step = ldexpf(1.f, p->fsr - p->bit_width);
    a.q.maxv = ldexpf(1.f, p->fsr);
  }
  if (p->mode == QT_Q_SPLIT) QT_REQUIRE(p->codes_kind == 4 || p->codes_kind == 6, "qt_quant_act: QT_Q_SPLIT needs codes_kind 4 or 6");
  if (p->mode == QT_Q_LOG)
    QT_REQUIRE((p->codes_kind == 0 || p->codes_kind == 3) && !p->bits && !p->row_sum,
               "qt_quant_act: the Log quantizer emits fp32 and / or bf16 values (codes_kind 3)");
  if (p->mode == QT_Q_LIN) {
    QT_REQUIRE((p->codes_kind >= 0 && p->codes_kind <= 3) && !p->bits, "qt_quant_act: the Lin quantizer emits int8 / uint8 / bf16 codes");
    if (p->codes_kind == 1) QT_REQUIRE(p->bit_width <= 6, "qt_quant_act: int8 Lin codes need bit_width <= 6");
    if (p->codes_kind == 2) QT_REQUIRE(p->bit_width <= 7 && !p->with_sign, "qt_quant_act: uint8 Lin codes need with_sign = 0, bit_width <= 7");
  }
  QT_REQUIRE(p->codes_kind >= 0 && p->codes_kind <= 7, "qt_quant_act: bad codes_kind");
  if (p->codes_kind == 7) {
    QT_REQUIRE(p->mode == QT_Q_SIGN || p->mode == QT_Q_TERNARY || (p->mode == QT_Q_DOREFA && p->bit_width == 2),
               "qt_quant_act: fp4 codes hold SIGN / TERNARY / DoReFa-2 codes only");
    QT_REQUIRE(p->ld_codes % 32 == 0 && p->nhwc_c == 0, "qt_quant_act: fp4 code rows must be multiples of 32 elements (row-major only)");
  }
  if (p->codes_kind) QT_REQUIRE(p->codes && p->ld_codes >= p->cols, "qt_quant_act: bad codes buffer");
  if (p->bits) QT_REQUIRE(p->mode == QT_Q_SIGN && p->ld_bits * 32 >= p->cols, "qt_quant_act: bits need QT_Q_SIGN and ld_bits*32 >= cols");
  if (p->y) QT_REQUIRE(p->ld_y >= p->cols, "qt_quant_act: ld_y < cols");
  a.x = p->x; a.rows = p->rows; a.cols = p->cols; a.ld_x = p->ld_x;
  a.y = p->y; a.ld_y = p->ld_y; a.codes = p->codes; a.codes_kind = p->codes_kind; a.ld_codes = p->ld_codes;
  a.bits = p->bits; a.ld_bits = p->ld_bits; a.row_sum = p->row_sum; a.row_scale = p->row_scale; a.overflow = p->overflow;

  if (p->ready) QT_REQUIRE(p->ready_rows > 0, "qt_quant_act: ready needs ready_rows > 0");
  if (int lean = try_lean_quant(p, stream)) return lean > 0 ? QT_OK : lean;
  if (p->ready) { set_error("qt_quant_act: progress counters (ready) are served by the lean code-only kernels only (cols %% 1024 == 0, no y / bits)"); return QT_EUNSUPPORTED; }

  if (p->nhwc_c > 0) {
    QT_REQUIRE(p->codes && (p->codes_kind == 1 || p->codes_kind == 2), "qt_quant_act: NHWC output needs int8/uint8 codes");
    QT_REQUIRE(p->mode == QT_Q_SIGN || p->mode == QT_Q_TERNARY || p->mode == QT_Q_DOREFA || p->mode == QT_Q_LIN,
               "qt_quant_act: NHWC output: SIGN / TERNARY / DOREFA / LIN only");
    QT_REQUIRE(p->cols % p->nhwc_c == 0 && p->ld_x == p->cols && (!p->y || p->ld_y == p->cols) && !p->bits && !p->row_sum,
               "qt_quant_act: NHWC output needs dense NCHW input and no bits / row sums");
    NhwcArgs n;
    n.q = a.q; n.x = p->x; n.y = p->y; n.codes = (int8_t*)p->codes; n.codes_kind = p->codes_kind;
    n.B = p->rows; n.C = p->nhwc_c; n.HW = p->cols / p->nhwc_c; n.overflow = p->overflow;
    QT_REQUIRE(n.B <= 65535 && ceil_div(n.C, 128) <= 65535, "qt_quant_act: NHWC grid too large");
    dim3 grid((unsigned)ceil_div(n.HW, 32), (unsigned)ceil_div(n.C, 128), (unsigned)n.B);
    if (p->mode == QT_Q_SIGN) act_quant_nhwc_kernel<QT_Q_SIGN><<<grid, 256, 0, stream>>>(n);
    else if (p->mode == QT_Q_TERNARY) act_quant_nhwc_kernel<QT_Q_TERNARY><<<grid, 256, 0, stream>>>(n);
    else if (p->mode == QT_Q_LIN) act_quant_nhwc_kernel<QT_Q_LIN><<<grid, 256, 0, stream>>>(n);
    else act_quant_nhwc_kernel<QT_Q_DOREFA><<<grid, 256, 0, stream>>>(n);
    QT_LAUNCH_CHECK();
    return QT_OK;
  }
  bool vec = (p->cols % 4 == 0) && (p->ld_x % 4 == 0) && aligned(p->x, 16);
  if (p->y) vec = vec && (p->ld_y % 4 == 0) && aligned(p->y, 16);
  if (p->codes_kind == 1 || p->codes_kind == 2) vec = vec && (p->ld_codes % 4 == 0) && aligned(p->codes, 4);
  if (p->codes_kind >= 3 && p->codes_kind <= 6) vec = vec && (p->ld_codes % 4 == 0) && aligned(p->codes, 8) && ((p->rows * p->ld_codes) % 4 == 0);
  if (p->codes_kind == 7) vec = vec && aligned(p->codes, 4);

  // split long rows into chunks when there are too few rows to fill the machine
  // warp tasks of <= 2048 columns: many small tasks keep the last wave short (8192 x 4096 -> 16384 tasks, ~3.5 waves of
  // 4 CTAs/SM instead of 1.7); the XNOR row mean needs the whole row in one warp
  a.chunk = p->cols; a.nchunks = 1;
  if (p->mode != QT_Q_XNOR_ROW && p->cols >= 2048) {
    static int qchunk = 0;
    if (qchunk == 0) {
      const char* e = getenv("QTB200_QCHUNK");
      qchunk = e ? atoi(e) : 2048;
      if (qchunk < 128 || qchunk % 128) qchunk = 2048;
    }
    a.chunk = qchunk;
    a.nchunks = (int)ceil_div(p->cols, a.chunk);
  }
  if (p->mode == QT_Q_XNOR_ROW) {
    // one-pass form (no fp32 output): chunks of 1024 columns, each writing a partial row SUM (qt_quant_xnor_parts)
    a.nchunks = qt_quant_xnor_parts(p->cols, p->y != nullptr, p->row_parts);
    if (a.nchunks > 1) a.chunk = 1024;
  }
  if (a.nchunks > 1 && p->row_sum) QT_CUDA_OK(cudaMemsetAsync(p->row_sum, 0, sizeof(int32_t) * p->rows, stream));
  const int warps_per_block = 8;
  int64_t tasks = p->rows * a.nchunks;
  dim3 grid((unsigned)ceil_div(tasks, warps_per_block)), block(32 * warps_per_block);
#define QT_ACT_LAUNCH(MODE_)                                                              \
  case MODE_:                                                                            \
    if (p->pre_scale) {                                                                  \
      if (vec) act_quant_kernel<MODE_, true, true><<<grid, block, 0, stream>>>(a);       \
      else act_quant_kernel<MODE_, false, true><<<grid, block, 0, stream>>>(a);          \
    } else {                                                                             \
      if (vec) act_quant_kernel<MODE_, true, false><<<grid, block, 0, stream>>>(a);      \
      else act_quant_kernel<MODE_, false, false><<<grid, block, 0, stream>>>(a);         \
    }                                                                                    \
    break;
  switch (p->mode) {
    QT_ACT_LAUNCH(QT_Q_SIGN)
    QT_ACT_LAUNCH(QT_Q_TERNARY)
    QT_ACT_LAUNCH(QT_Q_DOREFA)
    QT_ACT_LAUNCH(QT_Q_XNOR_ROW)
    QT_ACT_LAUNCH(QT_Q_LOG)
    QT_ACT_LAUNCH(QT_Q_LIN)
    QT_ACT_LAUNCH(QT_Q_SPLIT)
  }
#undef QT_ACT_LAUNCH
  QT_LAUNCH_CHECK();
  return QT_OK;
}

static int lane_bits_for(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }

extern "C" int qt_pack_weight(const QtWeightPack* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(p && p->w && p->packed && p->stats, "qt_pack_weight: null argument");
  QT_REQUIRE(p->mode >= QT_W_SIGN && p->mode <= QT_W_XNOR, "qt_pack_weight: bad mode %d", p->mode);
  QT_REQUIRE(p->n > 0 && p->k > 0 && p->ld_w >= p->k, "qt_pack_weight: bad shape");
  QT_REQUIRE(p->ld_packed % 4 == 0, "qt_pack_weight: ld_packed must be a multiple of 4 bytes");
  PackArgs a;
  a.mode = p->mode; a.bit_width = p->bit_width; a.lane_bits = 1;
  if (p->mode == QT_W_DOREFA) {
    QT_REQUIRE(p->bit_width >= 1 && p->bit_width <= 8, "qt_pack_weight: DoReFa bit width %d not in 1..8", p->bit_width);
    a.lane_bits = lane_bits_for(p->bit_width);
  }
  QT_REQUIRE(p->ld_packed * 8 >= p->k * a.lane_bits, "qt_pack_weight: ld_packed too small");
  if (p->mode == QT_W_XNOR) QT_REQUIRE(p->alpha, "qt_pack_weight: QT_W_XNOR needs alpha");
  a.w = p->w; a.n = p->n; a.k = p->k; a.ld_w = p->ld_w;
  a.packed = (uint8_t*)p->packed; a.ld_packed = p->ld_packed; a.stats = p->stats; a.wq = p->wq; a.alpha = p->alpha;

  // reductions: stats[0..3] + int zero counter at stats[8] + double accumulator at stats[12..13]
  QT_CUDA_OK(cudaMemsetAsync(p->stats, 0, 16 * sizeof(float), stream));
  double* dsum = reinterpret_cast<double*>(p->stats + 12);
  QT_REQUIRE(aligned(dsum, 8), "qt_pack_weight: stats must be 8-byte aligned");
  int blocks = (int)std::min<int64_t>(ceil_div(p->n * p->k, 256 * 8), 148 * 8);
  weight_stats_kernel<<<blocks, 256, 0, stream>>>(p->w, p->n, p->k, p->ld_w, p->stats, dsum);
  QT_LAUNCH_CHECK();
  weight_stats_finish_kernel<<<1, 1, 0, stream>>>(p->stats, dsum, (double)p->n * (double)p->k);
  QT_LAUNCH_CHECK();
  if (p->mode == QT_W_XNOR && !p->alpha_is_input) {
    col_absmean_kernel<<<(unsigned)ceil_div(p->k, 32), 256, 0, stream>>>(p->w, p->n, p->k, p->ld_w, p->alpha);
    QT_LAUNCH_CHECK();
  }
  weight_pack_kernel<<<(unsigned)ceil_div(p->n, 8), 256, 0, stream>>>(a);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_col_absmean(const float* w, int64_t n, int64_t k, int64_t ld_w, float* alpha, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(w && alpha && n > 0 && k > 0 && ld_w >= k, "qt_col_absmean: bad argument");
  col_absmean_kernel<<<(unsigned)ceil_div(k, 32), 256, 0, stream>>>(w, n, k, ld_w, alpha);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_expand_weight(const QtWeightExpand* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(p && p->packed && p->out, "qt_expand_weight: null argument");
  QT_REQUIRE(p->mode >= QT_W_SIGN && p->mode <= QT_W_XNOR, "qt_expand_weight: bad mode");
  QT_REQUIRE(p->out_kind >= 1 && p->out_kind <= 7, "qt_expand_weight: bad out_kind");
  if (p->out_kind == 7) {
    QT_REQUIRE(p->mode == QT_W_SIGN || p->mode == QT_W_TERNARY || (p->mode == QT_W_DOREFA && p->bit_width <= 2),
               "qt_expand_weight: fp4 operands hold sign / ternary / DoReFa k<=2 weights only");
    QT_REQUIRE(p->ld_out % 32 == 0, "qt_expand_weight: fp4 rows must be multiples of 32 elements");
  }
  QT_REQUIRE(p->ld_out % 16 == 0 && p->ld_out >= p->k, "qt_expand_weight: ld_out must be a multiple of 16 and >= k");
  QT_REQUIRE(aligned(p->out, 16), "qt_expand_weight: out must be 16-byte aligned");
  if (p->out_kind == 4 || p->out_kind == 5) QT_REQUIRE(p->alpha && p->mode == QT_W_XNOR, "qt_expand_weight: kinds 4/5 need XNOR alpha");
  if (p->out_kind == 2) QT_REQUIRE(p->mode == QT_W_DOREFA && p->bit_width >= 2, "qt_expand_weight: uint8 codes are DoReFa only");
  if (p->out_kind == 1 && p->mode == QT_W_DOREFA) QT_REQUIRE(p->bit_width <= 7, "qt_expand_weight: centred int8 codes need k <= 7");
  ExpandArgs a;
  a.mode = p->mode; a.bit_width = p->bit_width;
  a.lane_bits = (p->mode == QT_W_DOREFA) ? lane_bits_for(p->bit_width) : 1;
  a.packed = (const uint8_t*)p->packed; a.n = p->n; a.k = p->k; a.ld_packed = p->ld_packed; a.alpha = p->alpha;
  a.out = p->out; a.out_kind = p->out_kind; a.ld_out = p->ld_out;
  const int cpv = (p->out_kind == 7) ? 32 : ((p->out_kind == 1 || p->out_kind == 2) ? 16 : 8);
  const int64_t threads = p->n * (p->ld_out / cpv);
  QT_REQUIRE(threads < (1ll << 31), "qt_expand_weight: weight matrix too large (n * ld_out / %d >= 2^31)", cpv);
  const unsigned blocks = (unsigned)ceil_div(threads, 256 * 4);     // four vectors per thread
  if (cpv == 32) weight_expand_kernel<32><<<blocks, 256, 0, stream>>>(a);
  else if (cpv == 16) weight_expand_kernel<16><<<blocks, 256, 0, stream>>>(a);
  else weight_expand_kernel<8><<<blocks, 256, 0, stream>>>(a);
  QT_LAUNCH_CHECK();
  return QT_OK;
}


extern "C" int qt_transpose_split(const float* x, int64_t rows, int64_t cols, int64_t ld_x, void* out, int64_t ld_out, int planes,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(x && out, "qt_transpose_split: null argument");
  QT_REQUIRE(rows >= 0 && cols >= 0 && ld_x >= cols && ld_out >= rows, "qt_transpose_split: bad shape");
  QT_REQUIRE(planes >= 1 && planes <= 3, "qt_transpose_split: planes must be 1, 2 or 3");
  if (rows == 0 || cols == 0) return QT_OK;
  dim3 grid((unsigned)ceil_div(cols, 64), (unsigned)ceil_div(ld_out, 64));
  QT_REQUIRE(grid.y <= 65535, "qt_transpose_split: too many rows");
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (planes == 1) transpose_split_kernel<1><<<grid, 256, 0, stream>>>(x, rows, cols, ld_x, o, ld_out);
  else if (planes == 2) transpose_split_kernel<2><<<grid, 256, 0, stream>>>(x, rows, cols, ld_x, o, ld_out);
  else transpose_split_kernel<3><<<grid, 256, 0, stream>>>(x, rows, cols, ld_x, o, ld_out);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_ste_clip(const float* g, const float* x, float thresh, float* out, int64_t n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(g && x && out && n >= 0, "qt_ste_clip: bad argument");
  if (n == 0) return QT_OK;
  ste_clip_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, stream>>>(g, x, thresh, out, n);
  QT_LAUNCH_CHECK();
  return QT_OK;
}


extern "C" int qt_expand_loglin(const void* codes, int64_t n, int64_t k, int64_t ld_codes, int is_log, int emin, void* out,
                                int64_t ld_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(codes && out, "qt_expand_loglin: null argument");
  QT_REQUIRE(n >= 0 && k >= 0 && ld_codes >= k && ld_out >= k && ld_out % 8 == 0 && aligned(out, 16), "qt_expand_loglin: bad shape");
  if (n == 0 || k == 0) return QT_OK;
  const int64_t threads = n * (ld_out / 8);
  loglin_expand_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, stream>>>((const int8_t*)codes, n, k, ld_codes, is_log, emin,
                                                                              (__nv_bfloat16*)out, ld_out);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_rowsum_codes(const void* codes, int is_unsigned, int64_t rows, int64_t ld, int32_t* row_sum, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(codes && row_sum, "qt_rowsum_codes: null argument");
  QT_REQUIRE(rows >= 0 && ld > 0 && ld % 4 == 0 && aligned(codes, 4), "qt_rowsum_codes: rows of 8-bit codes must be 4-byte aligned multiples of 4");
  if (rows == 0) return QT_OK;
  const uint8_t* a = reinterpret_cast<const uint8_t*>(codes);
  if (is_unsigned) rowsum_i8_kernel<true><<<(unsigned)ceil_div(rows, 8), 256, 0, stream>>>(a, rows, ld, row_sum);
  else rowsum_i8_kernel<false><<<(unsigned)ceil_div(rows, 8), 256, 0, stream>>>(a, rows, ld, row_sum);
  QT_LAUNCH_CHECK();
  return QT_OK;
}

extern "C" int qt_im2col(const QtIm2col* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  QT_REQUIRE(p && p->x && p->out, "qt_im2col: null argument");
  QT_REQUIRE(p->elem_bytes == 1 || p->elem_bytes == 2 || p->elem_bytes == 4, "qt_im2col: elem_bytes must be 1, 2 or 4");
  QT_REQUIRE(p->groups >= 1 && p->C % p->groups == 0 && p->group >= 0 && p->group < p->groups, "qt_im2col: bad groups");
  Im2colArgs a;
  a.x = (const uint8_t*)p->x; a.eb = p->elem_bytes;
  a.B = p->B; a.C = p->C; a.H = p->H; a.W = p->W; a.OH = p->OH; a.OW = p->OW;
  a.kh = p->kh; a.kw = p->kw; a.sh = p->stride_h; a.sw = p->stride_w; a.ph = p->pad_h; a.pw = p->pad_w;
  a.dh = p->dil_h; a.dw = p->dil_w;
  a.cg = p->C / p->groups; a.c_begin = a.cg * p->group;
  a.out = (uint8_t*)p->out; a.ld_out = p->ld_out; a.kcols = a.cg * p->kh * p->kw;
  a.nhwc = p->nhwc;
  QT_REQUIRE(p->ld_out >= a.kcols, "qt_im2col: ld_out < C/groups*kh*kw");
  const int64_t M = p->B * p->OH * p->OW;
  if (M == 0) return QT_OK;
  if (p->split3) {
    QT_REQUIRE(p->split3 == 1 || p->split3 == 2 || p->split3 == 3, "qt_im2col: split3 must be 0, 1 (= 3 planes), 2 or 3");
    QT_REQUIRE(p->elem_bytes == 4 && !p->nhwc && !p->row_sum, "qt_im2col: split3 needs fp32 NCHW input and no row sums");
    QT_REQUIRE(p->ld_out % 8 == 0 && aligned(p->out, 16) && (M * p->ld_out) % 8 == 0, "qt_im2col: split3 needs ld_out % 8 == 0");
    QT_REQUIRE(a.kcols <= 4096 && p->C * p->H * p->W < (1ll << 31), "qt_im2col: split3 supports up to 4096 gathered columns");
    int64_t vecs = M * (p->ld_out / 8);
    unsigned nb = (unsigned)std::min<int64_t>(ceil_div(vecs, 256), 148 * 16);
    if (p->split3 == 2) im2col_split3_kernel<2><<<nb, 256, 2 * a.kcols * sizeof(int), stream>>>(a, M * p->ld_out);
    else im2col_split3_kernel<3><<<nb, 256, 2 * a.kcols * sizeof(int), stream>>>(a, M * p->ld_out);
    QT_LAUNCH_CHECK();
    return QT_OK;
  }
  const int vec = 16 / p->elem_bytes;
  QT_REQUIRE(p->ld_out % vec == 0 && aligned(p->out, 16), "qt_im2col: ld_out*elem_bytes must be a multiple of 16, out 16B aligned");
  int64_t threads = M * (p->ld_out / vec);
  unsigned blocks = (unsigned)ceil_div(threads, 256);
  const bool fast = p->nhwc && ((a.cg * p->elem_bytes) % 16 == 0) && ((p->C * p->elem_bytes) % 16 == 0) && aligned(p->x, 16);
  if (fast) im2col_nhwc_vec_kernel<<<blocks, 256, 0, stream>>>(a);
  else if (p->elem_bytes == 1) im2col_kernel<uint8_t, 16><<<blocks, 256, 0, stream>>>(a);
  else if (p->elem_bytes == 2) im2col_kernel<uint16_t, 8><<<blocks, 256, 0, stream>>>(a);
  else im2col_kernel<uint32_t, 4><<<blocks, 256, 0, stream>>>(a);
  QT_LAUNCH_CHECK();
  if (p->row_sum) {
    QT_REQUIRE(p->elem_bytes == 1, "qt_im2col: row_sum needs 1-byte codes");
    if (p->is_unsigned) rowsum_i8_kernel<true><<<(unsigned)ceil_div(M, 8), 256, 0, stream>>>(a.out, M, p->ld_out, p->row_sum);
    else rowsum_i8_kernel<false><<<(unsigned)ceil_div(M, 8), 256, 0, stream>>>(a.out, M, p->ld_out, p->row_sum);
    QT_LAUNCH_CHECK();
  }
  return QT_OK;
}
