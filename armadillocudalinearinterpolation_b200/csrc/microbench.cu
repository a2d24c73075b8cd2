// microbench.cu — two machine ceilings the roofline discussion needs and MEASURED_PEAKS.json
// does not carry ("the builder must measure"):
//   * random-gather ceiling: independent 32-byte record gathers from a table far larger than L2.
//     On B200 every L2 miss fills a whole 128-byte line, so this — not the streaming HBM rate —
//     bounds interp2 at scattered queries (profiles/interp2_scattered_r1.md);
//   * FP64 FMA issue ceiling: what the event-driven map's arithmetic is measured against.
#include "common.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
random_gather_kernel(const double* __restrict__ table, size_t nrec, double* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double* p = table + 4 * (mix64(i) % nrec);
    double a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
    out[i] = (a + b) + (c + d);
  }
}

__global__ void __launch_bounds__(256)
dfma_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_bench_random_gather(size_t table_bytes, size_t n_gathers, double* ms_out, double* gathers_per_s) {
  if (!ms_out || table_bytes < 4096 || n_gathers == 0) return fail(B200_ERR_INVALID_ARG, "bench_random_gather: bad argument");
  B200_TRY(require_device());
  const size_t nrec = table_bytes / 32;
  double *tab = nullptr, *out = nullptr;
  B200_CUDA(cudaMalloc(&tab, nrec * 32));
  cudaError_t e = cudaMalloc(&out, n_gathers * sizeof(double));
  if (e != cudaSuccess) { cudaFree(tab); return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
  cudaMemset(tab, 0, nrec * 32);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    random_gather_kernel<<<148 * 16, 256>>>(tab, nrec, out, n_gathers);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(tab); cudaFree(out);
  if (e != cudaSuccess) return cuda_fail(e, "random_gather_kernel", __FILE__, __LINE__);
  *ms_out = best;
  if (gathers_per_s) *gathers_per_s = (double)n_gathers / (best * 1e-3);
  return B200_OK;
}

int b200_bench_fp64_fma(double* tflops_out) {
  if (!tflops_out) return fail(B200_ERR_INVALID_ARG, "bench_fp64_fma: NULL");
  B200_TRY(require_device());
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  B200_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (e != cudaSuccess) return cuda_fail(e, "dfma_kernel", __FILE__, __LINE__);
  const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return B200_OK;
}

}  // extern "C"
