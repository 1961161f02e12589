// Runtime plumbing of the C-ABI: errors, device memory, copies.  No kernels here.
#include "vpb_common.cuh"
#include <stdarg.h>
#include <atomic>

namespace vpb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
int check_cuda(cudaError_t e, const char *what, const char *file, int line) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return (int)e;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace vpb

using namespace vpb;

extern "C" {

int vpb_version(void) { return VPB_VERSION; }
const char *vpb_last_error(void) { return g_err; }
int64_t vpb_launch_count(void) { return g_launches.load(); }

int vpb_device_count(int *count) { VPB_CUDA(cudaGetDeviceCount(count)); return 0; }
int vpb_set_device(int device) { VPB_CUDA(cudaSetDevice(device)); return 0; }
int vpb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_bytes) {
  cudaDeviceProp prop;
  VPB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_bytes) *total_bytes = prop.totalGlobalMem;
  return 0;
}
int vpb_malloc(void **dptr, size_t bytes) { VPB_CUDA(cudaMalloc(dptr, bytes ? bytes : 1)); return 0; }
int vpb_free(void *dptr) { VPB_CUDA(cudaFree(dptr)); return 0; }
int vpb_malloc_host(void **hptr, size_t bytes) { VPB_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1)); return 0; }
int vpb_free_host(void *hptr) { VPB_CUDA(cudaFreeHost(hptr)); return 0; }
int vpb_memset(void *dptr, int value, size_t bytes, void *stream) {
  VPB_CUDA(cudaMemsetAsync(dptr, value, bytes, as_stream(stream))); return 0;
}
int vpb_memcpy_h2d(void *dptr, const void *hptr, size_t bytes, void *stream) {
  VPB_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, as_stream(stream))); return 0;
}
int vpb_memcpy_d2h(void *hptr, const void *dptr, size_t bytes, void *stream) {
  VPB_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, as_stream(stream))); return 0;
}
int vpb_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  VPB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream))); return 0;
}
int vpb_stream_sync(void *stream) { VPB_CUDA(cudaStreamSynchronize(as_stream(stream))); return 0; }
int vpb_device_sync(void) { VPB_CUDA(cudaDeviceSynchronize()); return 0; }

}  // extern "C"
