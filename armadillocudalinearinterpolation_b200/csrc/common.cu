// common.cu — status/error plumbing and device helpers of the C-ABI (include/b200_common.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
unsigned long long count_launch() { return ++g_launches; }

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d in %s", (int)e,
           cudaGetErrorString(e), base ? base + 1 : file, line, what);
  cudaGetLastError();  // clear the sticky-less error state
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? B200_ERR_NO_DEVICE
                                                                      : B200_ERR_CUDA;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(B200_ERR_NO_DEVICE,
                "no CUDA device visible (%s): this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  int dev = 0, major = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return fail(B200_ERR_NO_DEVICE,
                "device %d has compute capability %d.x; kernels are built for sm_100a only", dev,
                major);
  return B200_OK;
}

}  // namespace b200

extern "C" {

const char* b200_last_error(void) { return b200::g_err; }

unsigned long long b200_launch_count(void) { return b200::g_launches.load(); }

int b200_version(void) { return (0 << 16) | (1 << 8) | 0; }

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b200_set_device(int device) {
  B200_CUDA(cudaSetDevice(device));
  return B200_OK;
}

int b200_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return b200::fail(B200_ERR_INVALID_ARG, "b200_host_alloc: ptr is NULL");
  B200_TRY(b200::require_device());
  B200_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return B200_OK;
}

int b200_host_free(void* ptr) {
  if (!ptr) return B200_OK;
  B200_CUDA(cudaFreeHost(ptr));
  return B200_OK;
}

int b200_synchronize(void) {
  B200_CUDA(cudaDeviceSynchronize());
  return B200_OK;
}

}  // extern "C"
