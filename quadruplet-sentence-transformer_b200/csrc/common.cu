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

// ---- launch counter --------------------------------------------------------------------------------
#include <atomic>
namespace qst {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace qst
extern "C" long long qst_launch_count(void) { return qst::g_launches.load(std::memory_order_relaxed); }

// ---- peer-visible device buffers (CUDA IPC) for the cross-rank threshold hints ------------------
extern "C" int qst_peer_buffer_create(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  QST_CHECK_ARG(dev_ptr && handle64 && bytes > 0, "peer_buffer_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == QST_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  QST_CUDA(cudaMalloc(&p, bytes));
  QST_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    qst::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return QST_ERR_CUDA;
  }
  memcpy(handle64, &h, sizeof(h));
  *dev_ptr = p;
  return QST_OK;
}

extern "C" int qst_peer_buffer_open(const unsigned char* handle64, void** dev_ptr) {
  QST_CHECK_ARG(dev_ptr && handle64, "peer_buffer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    qst::set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    return QST_ERR_CUDA;
  }
  *dev_ptr = p;
  return QST_OK;
}

extern "C" int qst_peer_buffer_clear(void* dev_ptr, size_t offset, size_t bytes, qst_stream_t stream) {
  QST_CHECK_ARG(dev_ptr != nullptr, "peer_buffer_clear: null pointer");
  QST_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned char*>(dev_ptr) + offset, 0, bytes, reinterpret_cast<cudaStream_t>(stream)));
  return QST_OK;
}

// Device-to-device copy into (or out of) a peer-mapped buffer, executed by the COPY ENGINES: no SM is
// involved, so it runs underneath a persistent kernel that occupies every SM (K2).
extern "C" int qst_peer_copy(void* dst, const void* src, size_t bytes, qst_stream_t stream) {
  QST_CHECK_ARG(dst && src, "peer_copy: null pointer");
  QST_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return QST_OK;
}

extern "C" int qst_peer_buffer_close(void* peer_ptr) {
  if (peer_ptr) QST_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return QST_OK;
}

extern "C" int qst_peer_buffer_destroy(void* dev_ptr) {
  if (dev_ptr) QST_CUDA(cudaFree(dev_ptr));
  return QST_OK;
}
