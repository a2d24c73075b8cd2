// tools/gather_granularity_test.cu — how fast does a B200 deliver uniformly random table units, as a function of
// the table size, the unit size (32 / 64 / 128 bytes) and the L2 fill granularity hint?  1e8 units per launch,
// one 8-byte store per unit, nothing else.  Decides whether any table layout for the scattered interp2 kernel can
// beat "one 128-byte line per query from the 228 MiB tile table".
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ggt tools/gather_granularity_test.cu && /tmp/ggt
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
template <int HINT> __device__ __forceinline__ double ld32(const double* p) {
  double a, b, c, d;
  if (HINT == 64) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
  else if (HINT == 128) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
  return (a + b) + (c + d);
}
// UNIT bytes per gather (32, 64 or 128), fetched as UNIT/32 vector loads of the same aligned unit
template <int UNIT, int HINT>
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ table, size_t nunits, double* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double* p = table + (UNIT / 8) * (mix64(i) % nunits);
    double s = 0;
#pragma unroll
    for (int k = 0; k < UNIT / 32; ++k) s += ld32<HINT>(p + 4 * k);
    out[i] = s;
  }
}
template <int UNIT, int HINT>
float run(const double* tab, size_t bytes, double* out, size_t n) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    gather_kernel<UNIT, HINT><<<148 * 16, 256>>>(tab, bytes / UNIT, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}
int main() {
  const size_t n = 100000000;
  const size_t sizes_mib[] = {64, 128, 171, 228, 288, 341, 512};
  double *tab, *out;
  cudaMalloc(&tab, (size_t)512 << 20); cudaMemset(tab, 0, (size_t)512 << 20);
  cudaMalloc(&out, n * 8);
  printf("%8s %10s %10s %10s %10s %10s %10s %10s   (ms per 1e8 units)\n", "MiB", "32B", "32B/64", "64B", "64B/64", "128B", "128B/64", "128B/128");
  for (size_t mib : sizes_mib) {
    const size_t b = mib << 20;
    printf("%8zu %10.3f %10.3f %10.3f %10.3f %10.3f %10.3f %10.3f\n", mib, run<32, 0>(tab, b, out, n), run<32, 64>(tab, b, out, n),
           run<64, 0>(tab, b, out, n), run<64, 64>(tab, b, out, n), run<128, 0>(tab, b, out, n), run<128, 64>(tab, b, out, n),
           run<128, 128>(tab, b, out, n));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
