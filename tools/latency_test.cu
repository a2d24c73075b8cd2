// dependent-chain latencies on one warp: DFMA, DADD, DMUL, FFMA, MUFU.RCP64H, LDS, double exp()
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double tab[64];
  tab[threadIdx.x & 63] = seed + threadIdx.x;
  __syncthreads();
  double a = seed, b = 1.0000001, c = 1e-9; float fa = (float)seed;
  long long t0, t1;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64(); cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = a + c;
  t1 = clock64(); cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = a * b;
  t1 = clock64(); cyc[2] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) fa = fmaf(fa, 1.0000001f, 1e-9f);
  t1 = clock64(); cyc[3] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y + 1.5; }
  t1 = clock64(); cyc[4] = t1 - t0;   // rcp + dadd
  t0 = clock64();
  int idx = threadIdx.x & 63;
#pragma unroll
  for (int i = 0; i < 64; ++i) { double y = tab[idx]; idx = ((int)y) & 63; }
  t1 = clock64(); cyc[5] = t1 - t0;   // lds + cvt
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) a = exp(-a * 1e-3);
  t1 = clock64(); cyc[6] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) a = 1.0 / (a + 1.0);
  t1 = clock64(); cyc[7] = t1 - t0;
  out[threadIdx.x] = a + fa + idx;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 8 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(o, c, 1.25);
  long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles/op): DFMA %.1f DADD %.1f DMUL %.1f FFMA %.1f | rcp64h+dadd %.1f | lds+cvt %.1f | exp() %.1f | 1/(a+1) %.1f\n",
         h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 64.0, h[5] / 64.0, h[6] / 32.0, h[7] / 32.0);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
