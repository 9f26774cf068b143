// Error reporting + device queries for libqst.
#include "qst_common.cuh"
#include <string.h>

namespace qst {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace qst

extern "C" int qst_version(void) { return QST_VERSION; }
extern "C" const char* qst_last_error(void) { return qst::g_err; }

extern "C" int qst_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  QST_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) { QST_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
  if (cc_major) { QST_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
  if (cc_minor) { QST_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
  return QST_OK;
}
