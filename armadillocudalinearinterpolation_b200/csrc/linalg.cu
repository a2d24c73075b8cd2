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

struct Free { void* p = nullptr; bool host = false; ~Free() { if (p) { if (host) free(p); else cudaFree(p); } } };
}  // namespace

extern "C" int b200_eig_gen_f64(size_t n, const double* a_colmajor, double* w_re, double* w_im) {
  if (!a_colmajor || !w_re || !w_im) return fail(B200_ERR_INVALID_ARG, "eig_gen: NULL argument");
  if (n == 0) return B200_OK;
  B200_TRY(require_device());
  if (!load_solver()) return fail(B200_ERR_UNSUPPORTED, "eig_gen: cuSOLVER (libcusolver.so.11 with cusolverDnXgeev) could not be loaded: %s", dlerror());
  Solver& s = g_solver;
  cusolverDnHandle_t hd = nullptr;
  cusolverDnParams_t pr = nullptr;
  if (s.create(&hd) != CUSOLVER_STATUS_SUCCESS) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnCreate failed");
  struct Guard { Solver& s; cusolverDnHandle_t h; cusolverDnParams_t* p; ~Guard() { if (*p) s.destroy_params(*p); s.destroy(h); } } guard{s, hd, &pr};
  if (s.create_params(&pr) != CUSOLVER_STATUS_SUCCESS) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnCreateParams failed");
  Free dA, dW, dWork, dInfo, hWork;
  hWork.host = true;
  B200_CUDA(cudaMalloc(&dA.p, n * n * sizeof(double)));
  B200_CUDA(cudaMalloc(&dW.p, n * 2 * sizeof(double)));
  B200_CUDA(cudaMalloc(&dInfo.p, sizeof(int)));
  B200_CUDA(cudaMemcpy(dA.p, a_colmajor, n * n * sizeof(double), cudaMemcpyHostToDevice));
  size_t wdev = 0, whost = 0;
  cusolverStatus_t st = s.geev_size(hd, pr, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, (int64_t)n, CUDA_R_64F, dA.p,
                                    (int64_t)n, CUDA_C_64F, dW.p, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F,
                                    &wdev, &whost);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnXgeev_bufferSize status %d", (int)st);
  B200_CUDA(cudaMalloc(&dWork.p, wdev ? wdev : 1));
  hWork.p = malloc(whost ? whost : 1);
  if (!hWork.p) return fail(B200_ERR_INVALID_ARG, "eig_gen: out of host memory");
  st = s.geev(hd, pr, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, (int64_t)n, CUDA_R_64F, dA.p, (int64_t)n, CUDA_C_64F,
              dW.p, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dWork.p, wdev, hWork.p, whost, (int*)dInfo.p);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnXgeev status %d", (int)st);
  B200_CUDA(cudaDeviceSynchronize());
  int info = 0;
  B200_CUDA(cudaMemcpy(&info, dInfo.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (info != 0) return fail(B200_ERR_CUDA, "eig_gen: cusolverDnXgeev info %d (QR iteration failed to converge)", info);
  std::vector<double> w(2 * n);
  B200_CUDA(cudaMemcpy(w.data(), dW.p, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) { w_re[i] = w[2 * i]; w_im[i] = w[2 * i + 1]; }
  return B200_OK;
}
