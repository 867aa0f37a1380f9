// Library-level entry points: version, error string, device capabilities, launch counter.
#include <stdarg.h>
#include <string.h>
#include "qt_common.cuh"

namespace qt {
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

int check_epi(const QtEpilogue* e, int64_t M, int64_t N) {
  QT_REQUIRE(e != nullptr, "epilogue is NULL");
  QT_REQUIRE(e->out != nullptr || e->acc_out != nullptr || e->requant != nullptr, "epilogue: neither out, acc_out nor requant given");
  QT_REQUIRE(e->row_scale_parts >= 0 && e->row_sum_parts >= 0, "epilogue: negative partial-sum count");
  QT_REQUIRE(e->row_scale_parts == 0 || e->row_scale != nullptr, "epilogue: row_scale_parts without row_scale");
  QT_REQUIRE(e->row_sum_parts == 0 || e->row_sum != nullptr, "epilogue: row_sum_parts without row_sum");
  if (const QtRequant* r = e->requant) {
    QT_REQUIRE(r->mode == QT_Q_SIGN || r->mode == QT_Q_TERNARY || r->mode == QT_Q_DOREFA || r->mode == QT_Q_XNOR_ROW,
               "requant: mode must be QT_Q_SIGN, QT_Q_TERNARY, QT_Q_DOREFA or QT_Q_XNOR_ROW");
    QT_REQUIRE(r->mode != QT_Q_DOREFA || (r->bit_width >= 2 && r->bit_width <= 8), "requant: DoReFa bit width must be 2..8");
    QT_REQUIRE(r->codes != nullptr, "requant: codes is NULL");
    QT_REQUIRE(r->codes_kind == 1 || r->codes_kind == 2 || r->codes_kind == 3 || r->codes_kind == 5 || r->codes_kind == 7,
               "requant: codes_kind must be 1 (int8), 2 (uint8), 3 (bf16), 5 (fp16) or 7 (fp4)");
    QT_REQUIRE(r->codes_kind != 7 || r->mode == QT_Q_SIGN || r->mode == QT_Q_TERNARY || (r->mode == QT_Q_DOREFA && r->bit_width == 2) ||
               r->mode == QT_Q_XNOR_ROW, "requant: fp4 codes hold integers in [-4, 4] only");
    QT_REQUIRE(r->mode != QT_Q_DOREFA || r->codes_kind == 1 || r->codes_kind == 2 || r->codes_kind == 7,
               "requant: DoReFa codes go to int8, uint8 or fp4 lanes");
    QT_REQUIRE(r->mode == QT_Q_DOREFA || r->codes_kind != 2, "requant: uint8 lanes hold DoReFa codes only");
    QT_REQUIRE((r->mode == QT_Q_XNOR_ROW) == (r->codes_kind == 3 || r->codes_kind == 5),
               "requant: XnorNet codes go to bf16 / fp16 lanes, integer codes to int8 / uint8 / fp4 lanes");
    QT_REQUIRE(r->ld_codes % 32 == 0 && r->ld_codes >= (N + 31) / 32 * 32, "requant: ld_codes must be a multiple of 32 and >= N rounded up to 32");
    QT_REQUIRE((reinterpret_cast<uintptr_t>(r->codes) & 15) == 0, "requant: codes must be 16-byte aligned");
    QT_REQUIRE(r->mode != QT_Q_XNOR_ROW || r->row_part != nullptr, "requant: QT_Q_XNOR_ROW needs row_part");
    QT_REQUIRE(r->cover == 0 || (r->cover % 32 == 0 && r->cover <= r->ld_codes && r->cover >= (N + 31) / 32 * 32),
               "requant: cover must be a multiple of 32 in [N rounded up to 32, ld_codes]");
  }
  QT_REQUIRE(e->out_mode == 0 || e->out_mode == 1, "epilogue: out_mode must be 0 or 1");
  if (e->residual) QT_REQUIRE(e->out_mode == 0 && e->ld_res >= N, "epilogue: residual needs out_mode 0 and ld_res >= N");
  if (e->a_ready) QT_REQUIRE(e->a_ready_rows >= 128 && e->a_ready_rows % 128 == 0 && e->a_ready_target > 0, "epilogue: a_ready needs a_ready_rows % 128 == 0 and a positive target");
  if (e->out && e->out_mode == 0) QT_REQUIRE(e->ldo >= N, "epilogue: ldo (%lld) < N (%lld)", (long long)e->ldo, (long long)N);
  if (e->out && e->out_mode == 1)
    QT_REQUIRE(e->ldo >= N && e->nchw_inner > 0 && M % e->nchw_inner == 0, "epilogue: NCHW inner (%lld) must divide M (%lld)",
               (long long)e->nchw_inner, (long long)M);
  return QT_OK;
}
}  // namespace qt

extern "C" {

int qt_version(void) { return QT_VERSION; }

int qt_sizeof(const char* name) {
  if (!name) return -1;
#define QT_SZ(T) if (strcmp(name, #T) == 0) return (int)sizeof(T);
  QT_SZ(QtActQuant) QT_SZ(QtWeightPack) QT_SZ(QtWeightExpand) QT_SZ(QtIm2col) QT_SZ(QtRequant) QT_SZ(QtEpilogue) QT_SZ(QtConvGeom) QT_SZ(QtPoolGeom)
#undef QT_SZ
  return -1;
}

int qt_requant_max_parts(int64_t N) { return (int)(2 * ((N + 127) / 128) + 2); }

const char* qt_last_error(void) { return qt::g_err; }

int64_t qt_launch_count(int reset) {
  int64_t v = qt::g_launches;
  if (reset) qt::g_launches = 0;
  return v;
}

int qt_device_caps(int device, int* sm_major, int* sm_minor, int* num_sms, int* has_tcgen05) {
  cudaDeviceProp p;
  QT_CUDA_OK(cudaGetDeviceProperties(&p, device));
  if (sm_major) *sm_major = p.major;
  if (sm_minor) *sm_minor = p.minor;
  if (num_sms) *num_sms = p.multiProcessorCount;
  if (has_tcgen05) *has_tcgen05 = (p.major == 10 && p.minor == 0) ? 1 : 0;
  return QT_OK;
}

}  // extern "C"
