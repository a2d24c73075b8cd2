// micro-benchmark: random 32-byte record gathers from a 512 MiB table; DRAM bytes per gather
// as a function of cudaLimitMaxL2FetchGranularity and of the load flavour.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long x){ x += 0x9E3779B97F4A7C15ull; x=(x^(x>>30))*0xBF58476D1CE4E5B9ull; x=(x^(x>>27))*0x94D049BB133111EBull; return x^(x>>31);}  
template<int MODE> __global__ void gather(const double* __restrict__ tab, size_t nrec, double* __restrict__ out, size_t nq){
  size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  for (; i<nq; i+=stride){
    size_t r = mix(i) % nrec; const double* p = tab + 4*r; double a,b,c,d;
    if (MODE==0) asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p));
    else if (MODE==1) { asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];":"=d"(a),"=d"(b):"l"(p)); asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];":"=d"(c),"=d"(d):"l"(p+2)); }
    else if (MODE==2) { asm volatile("ld.global.nc.f64 %0, [%1];":"=d"(a):"l"(p)); b=c=d=0; }
    else if (MODE==3) { asm volatile("ld.global.cv.f64 %0, [%1];":"=d"(a):"l"(p)); b=c=d=0; }
    else if (MODE==4) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p));
    else if (MODE==5) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p));
    else if (MODE==6) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p));
    else if (MODE==7) asm volatile("ld.global.L1::no_allocate.L2::evict_first.L2::64B.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p));
    else { asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];":"=d"(a):"l"(p)); b=c=d=0; }
    out[i]=a+b+c+d;
  }
}
int main(int argc,char**argv){
  int gran = argc>1?atoi(argv[1]):0; int mode = argc>2?atoi(argv[2]):0;
  cudaSetDevice(0);
  if (gran) { cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,(size_t)gran); printf("set %d -> %s\n",gran,cudaGetErrorString(e)); }
  size_t g=0; cudaDeviceGetLimit(&g,cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n",g);
  size_t nrec=(size_t)16<<20; size_t nq=100000000; double*tab,*out; cudaMalloc(&tab,nrec*32); cudaMalloc(&out,nq*8); cudaMemset(tab,0,nrec*32);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int rep=0;rep<3;rep++){
    cudaEventRecord(e0);
    switch(mode){case 0: gather<0><<<148*16,256>>>(tab,nrec,out,nq); break; case 1: gather<1><<<148*16,256>>>(tab,nrec,out,nq); break; case 2: gather<2><<<148*16,256>>>(tab,nrec,out,nq); break; case 3: gather<3><<<148*16,256>>>(tab,nrec,out,nq); break;
      case 4: gather<4><<<148*16,256>>>(tab,nrec,out,nq); break; case 5: gather<5><<<148*16,256>>>(tab,nrec,out,nq); break; case 6: gather<6><<<148*16,256>>>(tab,nrec,out,nq); break; case 7: gather<7><<<148*16,256>>>(tab,nrec,out,nq); break; default: gather<8><<<148*16,256>>>(tab,nrec,out,nq);}
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1);
    printf("mode %d gran %d: %.3f ms  %.2f Ggather/s\n",mode,gran,ms,nq/ms/1e6);
  }
  printf("%s\n",cudaGetErrorString(cudaGetLastError()));
}
