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
};

static inline Epi make_epi(const QtEpilogue* e, int64_t M, int64_t N) {
  Epi d;
  d.bias = e->bias; d.row_scale = e->row_scale; d.col_scale = e->col_scale; d.row_sum = e->row_sum;
  d.scale = e->scale; d.acc_mul = e->acc_mul; d.rs_mul = e->rs_mul;
  d.out = e->out; d.ldo = e->ldo; d.out_mode = e->out_mode; d.nchw_inner = e->nchw_inner;
  d.acc_out = e->acc_out; d.M = M; d.N = N;
  return d;
}

int check_epi(const QtEpilogue* e, int64_t M, int64_t N);

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
  return y;
}
__device__ __forceinline__ float epi_f32(const Epi& e, int64_t m, int64_t n, float acc) {
  float y = acc * e.scale;
  if (e.row_scale) y *= __ldg(e.row_scale + m);
  if (e.col_scale) y *= __ldg(e.col_scale + n);
  if (e.bias) y += __ldg(e.bias + n);
  return y;
}
__device__ __forceinline__ int64_t epi_addr(const Epi& e, int64_t m, int64_t n) {
  if (e.out_mode == 0) return m * e.ldo + n;
  int64_t img = m / e.nchw_inner, r = m - img * e.nchw_inner;
  return (img * e.ldo + n) * e.nchw_inner + r;
}

}  // namespace qt
