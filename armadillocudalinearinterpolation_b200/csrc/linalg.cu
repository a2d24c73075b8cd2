// linalg.cu — dense eigenvalues of the finite-difference Jacobian (include/b200_edm.h:
// b200_eig_gen_f64).  Replaces the host LAPACK call behind arma::eig_gen in
// Stability::ComputeEigenvalues (/root/reference/Stability.cpp:72, :40): at the 1000 x 1000 size
// of BASELINE config 5 the eigen-solve is the Amdahl term of the stability analysis once the
// Jacobian is sharded over the GPUs, so it runs on the device that already holds the Jacobian,
// through cuSOLVER's 64-bit GEEV (a library call: nothing here is a hot kernel of this repo).
// cuSOLVER is loaded on first use (dlopen), so the C-ABI library itself carries no link-time
// dependency on it.
#include <dlfcn.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "b200_edm.h"

namespace {
using namespace b200;

struct Solver {
  void* so = nullptr;
  decltype(&cusolverDnCreate) create = nullptr;
  decltype(&cusolverDnDestroy) destroy = nullptr;
  decltype(&cusolverDnSetStream) set_stream = nullptr;
  decltype(&cusolverDnCreateParams) create_params = nullptr;
  decltype(&cusolverDnDestroyParams) destroy_params = nullptr;
  decltype(&cusolverDnXgeev_bufferSize) geev_size = nullptr;
  decltype(&cusolverDnXgeev) geev = nullptr;
  bool tried = false, ok = false;
};
Solver g_solver;
std::mutex g_mu;

bool load_solver() {
  std::lock_guard<std::mutex> lk(g_mu);
  Solver& s = g_solver;
  if (s.tried) return s.ok;
  s.tried = true;
  for (const char* name : {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11"}) {
    s.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (s.so) break;
  }
  if (!s.so) return false;
#define B200_SYM(field, sym) s.field = (decltype(s.field))dlsym(s.so, #sym); if (!s.field) return false
  B200_SYM(create, cusolverDnCreate);
  B200_SYM(destroy, cusolverDnDestroy);
  B200_SYM(set_stream, cusolverDnSetStream);
  B200_SYM(create_params, cusolverDnCreateParams);
  B200_SYM(destroy_params, cusolverDnDestroyParams);
  B200_SYM(geev_size, cusolverDnXgeev_bufferSize);
  B200_SYM(geev, cusolverDnXgeev);
#undef B200_SYM
  return s.ok = true;
}

}  // namespace

// cuSOLVER handle, parameters and buffers are kept per device between calls (creating a handle costs more than
// a 1000 x 1000 GEEV); guarded by g_call.
struct DevState {
  int device = -1;
  cusolverDnHandle_t hd = nullptr;
  cusolverDnParams_t pr = nullptr;
  void *dA = nullptr, *dW = nullptr, *dWork = nullptr, *hWork = nullptr;
  int* dInfo = nullptr;
  size_t capA = 0, capWork = 0, capHost = 0;
};
std::vector<DevState> g_states;
std::mutex g_call;

extern "C" int b200_eig_gen_f64(size_t n, const double* a_colmajor, double* w_re, double* w_im) {
  if (!a_colmajor || !w_re || !w_im) return fail(B200_ERR_INVALID_ARG, "eig_gen: NULL argument");
  if (n == 0) return B200_OK;
  B200_TRY(require_device());
  if (!load_solver()) return fail(B200_ERR_UNSUPPORTED, "eig_gen: cuSOLVER (libcusolver.so.11 with cusolverDnXgeev) could not be loaded: %s", dlerror());
  Solver& s = g_solver;
  std::lock_guard<std::mutex> lk(g_call);
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  DevState* st = nullptr;
  for (DevState& d : g_states) if (d.device == dev) st = &d;
  if (!st) {
    g_states.emplace_back();
    st = &g_states.back();
    st->device = dev;
    if (s.create(&st->hd) != CUSOLVER_STATUS_SUCCESS || s.create_params(&st->pr) != CUSOLVER_STATUS_SUCCESS) {
      g_states.pop_back();
      return fail(B200_ERR_CUDA, "eig_gen: cusolverDnCreate failed");
    }
    if (cudaMalloc(&st->dInfo, sizeof(int)) != cudaSuccess) {   // never keep a half-initialised state for the next call
      cudaGetLastError();
      s.destroy_params(st->pr); s.destroy(st->hd);
      g_states.pop_back();
      return fail(B200_ERR_CUDA, "eig_gen: out of device memory");
    }
  }
  if (n * n > st->capA) {
    cudaFree(st->dA); cudaFree(st->dW); st->dA = st->dW = nullptr; st->capA = 0;
    B200_CUDA(cudaMalloc(&st->dA, n * n * sizeof(double)));
    B200_CUDA(cudaMalloc(&st->dW, n * 2 * sizeof(double)));
    st->capA = n * n;
  }
  B200_CUDA(cudaMemcpy(st->dA, a_colmajor, n * n * sizeof(double), cudaMemcpyHostToDevice));
  size_t wdev = 0, whost = 0;
  cusolverStatus_t rc = s.geev_size(st->hd, st->pr, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, (int64_t)n, CUDA_R_64F,
                                    st->dA, (int64_t)n, CUDA_C_64F, st->dW, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1,
                                    CUDA_R_64F, &wdev, &whost);
  if (rc != CUSOLVER_STATUS_SUCCESS) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnXgeev_bufferSize status %d", (int)rc);
  if (wdev > st->capWork) {
    cudaFree(st->dWork); st->dWork = nullptr; st->capWork = 0;
    B200_CUDA(cudaMalloc(&st->dWork, wdev));
    st->capWork = wdev;
  }
  if (whost > st->capHost) {
    free(st->hWork);
    st->hWork = malloc(whost);
    st->capHost = st->hWork ? whost : 0;
    if (!st->hWork) return fail(B200_ERR_INVALID_ARG, "eig_gen: out of host memory");
  }
  rc = s.geev(st->hd, st->pr, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, (int64_t)n, CUDA_R_64F, st->dA, (int64_t)n,
              CUDA_C_64F, st->dW, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, st->dWork, wdev, st->hWork, whost,
              st->dInfo);
  if (rc != CUSOLVER_STATUS_SUCCESS) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnXgeev status %d", (int)rc);
  B200_CUDA(cudaDeviceSynchronize());
  int info = 0;
  B200_CUDA(cudaMemcpy(&info, st->dInfo, sizeof(int), cudaMemcpyDeviceToHost));
  if (info != 0) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnXgeev info %d (QR iteration failed to converge)", info);
  std::vector<double> w(2 * n);
  B200_CUDA(cudaMemcpy(w.data(), st->dW, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) { w_re[i] = w[2 * i]; w_im[i] = w[2 * i + 1]; }
  return B200_OK;
}
