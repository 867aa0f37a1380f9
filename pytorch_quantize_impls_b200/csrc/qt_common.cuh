// Shared helpers for the qtb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/qtb200.h"

namespace qt {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define QT_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      ::qt::set_error(__VA_ARGS__);                            \
      return QT_EINVAL;                                        \
    }                                                          \
  } while (0)

#define QT_CUDA_OK(expr)                                                          \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::qt::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));           \
      return QT_ECUDA;                                                            \
    }                                                                             \
  } while (0)

#define QT_LAUNCH_CHECK()                                                         \
  do {                                                                            \
    cudaError_t e__ = cudaPeekAtLastError();                                      \
    if (e__ != cudaSuccess) {                                                     \
      ::qt::set_error("kernel launch failed: %s", cudaGetErrorString(e__));       \
      (void)cudaGetLastError();                                                   \
      return QT_ECUDA;                                                            \
    }                                                                             \
    ::qt::count_launch();                                                         \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Device-side copy of the epilogue parameters (POD, passed by value to kernels).
struct Epi {
  const float* bias;
  const float* row_scale;
  const float* col_scale;
  const int32_t* row_sum;
  float scale;
  int32_t acc_mul, rs_mul;
  float* out;
  int64_t ldo;
  int out_mode;
  int64_t nchw_inner;
  int32_t* acc_out;
  int64_t M, N;
  // fused re-quantisation of the output (QtRequant) and partial-sum row operands (tcgen05 kernels only)
  int rq_mode;            // -1: none, else QT_Q_SIGN / QT_Q_TERNARY / QT_Q_DOREFA / QT_Q_XNOR_ROW
  int rq_codes_kind;
  void* rq_codes;
  int64_t rq_ld;
  int64_t rq_cover;       // columns [N, rq_cover) of a row receive zero codes (<= rq_ld)
  int rq_clamp;
  float rq_lo, rq_hi, rq_n;
  float* rq_row_part;
  int32_t* rq_row_sum_part;
  int32_t* rq_overflow;
  int row_scale_parts, row_sum_parts;
  float row_scale_mul;
  int out_clamp;
  float out_lo, out_hi;
  const float* residual;   // optional fp32 [M, ld_res] added to y before the clamp (row-major outputs only)
  int64_t ld_res;
  const int32_t* a_ready;  // progress counters of a concurrently running producer of the A rows (tcgen05 GEMM routes)
  int a_ready_rows, a_ready_target;
};

static inline Epi make_epi(const QtEpilogue* e, int64_t M, int64_t N) {
  Epi d;
  d.bias = e->bias; d.row_scale = e->row_scale; d.col_scale = e->col_scale; d.row_sum = e->row_sum;
  d.scale = e->scale; d.acc_mul = e->acc_mul; d.rs_mul = e->rs_mul;
  d.out = e->out; d.ldo = e->ldo; d.out_mode = e->out_mode; d.nchw_inner = e->nchw_inner;
  d.acc_out = e->acc_out; d.M = M; d.N = N;
  d.rq_mode = -1; d.rq_codes_kind = 0; d.rq_codes = nullptr; d.rq_ld = 0; d.rq_cover = 0; d.rq_clamp = 0; d.rq_lo = d.rq_hi = 0.f; d.rq_n = 1.f;
  d.rq_row_part = nullptr; d.rq_row_sum_part = nullptr; d.rq_overflow = nullptr;
  if (const QtRequant* r = e->requant) {
    d.rq_mode = r->mode; d.rq_codes_kind = r->codes_kind; d.rq_codes = r->codes; d.rq_ld = r->ld_codes; d.rq_cover = r->cover > 0 ? r->cover : r->ld_codes;
    d.rq_clamp = r->clamp; d.rq_lo = r->lo; d.rq_hi = r->hi;
    d.rq_n = (r->mode == QT_Q_DOREFA) ? (float)((1u << r->bit_width) - 1u) : 1.f;
    d.rq_row_part = r->row_part; d.rq_row_sum_part = r->row_sum_part; d.rq_overflow = r->overflow;
  }
  d.row_scale_parts = e->row_scale_parts; d.row_sum_parts = e->row_sum_parts; d.row_scale_mul = e->row_scale_mul;
  d.out_clamp = e->out_clamp; d.out_lo = e->out_lo; d.out_hi = e->out_hi;
  d.residual = e->residual; d.ld_res = e->ld_res;
  d.a_ready = e->a_ready; d.a_ready_rows = e->a_ready_rows; d.a_ready_target = e->a_ready_target;
  return d;
}

int check_epi(const QtEpilogue* e, int64_t M, int64_t N);
// true when the epilogue asks for something only the tcgen05 kernels implement (requant, partial-sum row operands, residual)
static inline bool epi_needs_tc(const Epi& e) {
  return e.rq_mode >= 0 || e.row_scale_parts > 0 || e.row_sum_parts > 0 || e.residual != nullptr || e.a_ready != nullptr;
}

// y = float(acc_mul*acc + rs_mul*row_sum[m]) * scale * row_scale[m] * col_scale[n] + bias[n]
// The integer part is exact; float(t) * 1.0f + bias is a single rounding, which is what makes the
// BinaryNet / Terner outputs bit-identical to the reference's fp32 addmm on +-1 operands.
__device__ __forceinline__ float epi_int(const Epi& e, int64_t m, int64_t n, int32_t acc) {
  int32_t t = e.acc_mul * acc;
  if (e.row_sum) t += e.rs_mul * __ldg(e.row_sum + m);
  float y = (float)t * e.scale;
  if (e.row_scale) y *= __ldg(e.row_scale + m);
  if (e.col_scale) y *= __ldg(e.col_scale + n);
  if (e.bias) y += __ldg(e.bias + n);
  if (e.out_clamp) y = fminf(fmaxf(y, e.out_lo), e.out_hi);
  return y;
}
__device__ __forceinline__ float epi_f32(const Epi& e, int64_t m, int64_t n, float acc) {
  float y = acc * e.scale;
  if (e.row_scale) y *= __ldg(e.row_scale + m);
  if (e.col_scale) y *= __ldg(e.col_scale + n);
  if (e.bias) y += __ldg(e.bias + n);
  if (e.out_clamp) y = fminf(fmaxf(y, e.out_lo), e.out_hi);
  return y;
}
__device__ __forceinline__ int64_t epi_addr(const Epi& e, int64_t m, int64_t n) {
  if (e.out_mode == 0) return m * e.ldo + n;
  int64_t img = m / e.nchw_inner, r = m - img * e.nchw_inner;
  return (img * e.ldo + n) * e.nchw_inner + r;
}

}  // namespace qt
