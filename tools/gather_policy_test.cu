// micro-benchmark: 1e8 random 32-byte gathers from a table of S MiB while 2.4 GB of query/output streams pass
// through L2 (the access pattern of interp2 at scattered queries), as a function of the table size and of the
// fraction of table lines that are given L2::evict_last (the rest evict_first): how much of the 126 MB L2 can
// be made to hold table lines, i.e. how many DRAM row activations per gather can be avoided.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long x){ x += 0x9E3779B97F4A7C15ull; x=(x^(x>>30))*0xBF58476D1CE4E5B9ull; x=(x^(x>>27))*0x94D049BB133111EBull; return x^(x>>31);}
__global__ void __launch_bounds__(512) gather(const double* __restrict__ tab, size_t nrec, const double* __restrict__ xs, const double* __restrict__ ys,
                       double* __restrict__ out, size_t nq, float frac, int mode){
  unsigned long long pol;
  if (mode == 0) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(frac));
  else asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, %1;" : "=l"(pol) : "f"(frac));
  size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  for (; i<nq; i+=stride){
    double x, y;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(x) : "l"(xs + i));
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(y) : "l"(ys + i));
    size_t r = mix(i + (size_t)(x * 0.0)) % nrec; const double* p = tab + 4*r; double a,b,c,d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p),"l"(pol));
    double v = a+b+c+d+y;
    asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(out + i), "d"(v) : "memory");
  }
}
int main(int argc,char**argv){
  cudaSetDevice(0);
  {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("L2 %d MB, persisting max %d MB, access policy max window %d MB\n", pr.l2CacheSize >> 20, pr.persistingL2CacheMaxSize >> 20, pr.accessPolicyMaxWindowSize >> 20);
    if (argc > 1) {
      size_t want = (size_t)atoi(argv[1]) << 20;
      cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
      printf("persisting L2 carve-out: asked %zu MB -> %s, now %zu MB\n", want >> 20, cudaGetErrorString(e), got >> 20);
    }
  }
  size_t nq=100000000; double *tab,*out,*xs,*ys;
  size_t maxb=(size_t)512<<20; cudaMalloc(&tab,maxb); cudaMalloc(&out,nq*8); cudaMalloc(&xs,nq*8); cudaMalloc(&ys,nq*8);
  cudaMemset(tab,0,maxb); cudaMemset(xs,0,nq*8); cudaMemset(ys,0,nq*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sizes[] = {128, 228, 512};
  float fr[] = {1.0f, 0.8f, 0.6f, 0.45f, 0.3f, 0.2f, 0.1f, 0.0f};
  for (int mode = 0; mode < 1; ++mode)
  for (int s : sizes) {
    size_t nrec = ((size_t)s << 20) / 32;
    printf("mode %d table %3d MiB:", mode, s);
    for (float f : fr) {
      float best = 1e9f;
      for(int rep=0;rep<3;rep++){
        cudaEventRecord(e0);
        gather<<<148*4,512>>>(tab,nrec,xs,ys,out,nq,f,mode);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1);
        if (rep && ms < best) best = ms;
      }
      printf("  f=%.2f %.3f", f, best);
    }
    printf("  ms\n");
  }
  printf("%s\n",cudaGetErrorString(cudaGetLastError()));
}
