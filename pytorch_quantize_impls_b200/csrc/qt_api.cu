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
  QT_REQUIRE(e->out != nullptr || e->acc_out != nullptr, "epilogue: neither out nor acc_out given");
  QT_REQUIRE(e->out_mode == 0 || e->out_mode == 1, "epilogue: out_mode must be 0 or 1");
  if (e->out && e->out_mode == 0) QT_REQUIRE(e->ldo >= N, "epilogue: ldo (%lld) < N (%lld)", (long long)e->ldo, (long long)N);
  if (e->out && e->out_mode == 1)
    QT_REQUIRE(e->ldo >= N && e->nchw_inner > 0 && M % e->nchw_inner == 0, "epilogue: NCHW inner (%lld) must divide M (%lld)",
               (long long)e->nchw_inner, (long long)M);
  return QT_OK;
}
}  // namespace qt

extern "C" {

int qt_version(void) { return QT_VERSION; }

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
