// micro-benchmark (round 2): does the B200's split L2 reward die affinity for random gathers?
//  1. sm -> die map from L2-hit latencies (near ~234 / far ~262 cycles), 2. home die of every 2 KB chunk of a
//  table, 3. 1e8 random 32-byte gathers (+ 24 B/query of streams, the interp2 access pattern) where each SM gathers
//  mode 0: anywhere; 1: only from "its" half of the table (virtual halves); 2: only from chunks homed on its own
//  die; 3: only from chunks homed on the other die.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned long long mix(unsigned long long x){ x += 0x9E3779B97F4A7C15ull; x=(x^(x>>30))*0xBF58476D1CE4E5B9ull; x=(x^(x>>27))*0x94D049BB133111EBull; return x^(x>>31);}
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ int lat_of(const unsigned long long* p) {
  // chain of dependent L2 loads on one address (table holds zeros): the clock is read after the 9th load is
  // issued, i.e. after 8 complete round trips; min over repeats, first repeat (DRAM miss) dropped
  int best = 1 << 30; unsigned long long off = 0;
  for (int r = 0; r < 4; ++r) {
    long long t0 = clock64();
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      unsigned long long v;
      asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p + off) : "memory");
      off = v;
    }
    long long t1 = clock64();
    int d = (int)((t1 - t0) / 8); if (r > 0 && d < best) best = d;
  }
  return best + (int)off;
}
__global__ void probe_sm(const unsigned long long* tab, int K, size_t stride_words, int* lat /*[nsm][K]*/, int* claimed) {
  if (threadIdx.x) return;
  unsigned s = smid();
  if (atomicCAS(&claimed[s], 0, 1)) return;
  for (int k = 0; k < K; ++k) lat[s * K + k] = lat_of(tab + (size_t)k * stride_words);
}
__global__ void probe_chunks(const unsigned long long* tab, size_t nchunks, int* lat, unsigned char* by_sm) {
  if (threadIdx.x) return;
  unsigned s = smid();
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) { lat[c] = lat_of(tab + c * 256); by_sm[c] = (unsigned char)s; }
}
__global__ void __launch_bounds__(512) gather(const double* __restrict__ tab, size_t nrec, const double* __restrict__ xs, const double* __restrict__ ys,
    double* __restrict__ out, size_t nq, int mode, const unsigned char* __restrict__ die_of_sm, const unsigned* __restrict__ list0,
    const unsigned* __restrict__ list1, unsigned n0, unsigned n1) {
  int d = die_of_sm[smid()];
  if (mode == 3) d ^= 1;
  const unsigned* lst = d ? list1 : list0; unsigned nl = d ? n1 : n0;
  size_t half = nrec / 2;
  size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  for (; i<nq; i+=stride){
    double x, y;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(x) : "l"(xs + i));
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(y) : "l"(ys + i));
    unsigned long long h = mix(i + (size_t)(x * 0.0));
    size_t r;
    if (mode == 0) r = h % nrec;
    else if (mode == 1) r = (h % half) + (d ? half : 0);
    else { unsigned c = lst[(h >> 8) % nl]; r = (size_t)c * 64 + (h & 63); }   // 64 records of 32 B per 2 KB chunk
    const double* p = tab + 4*r; double a,b,c2,dd;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c2),"=d"(dd):"l"(p));
    double v = a+b+c2+dd+y;
    asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(out + i), "d"(v) : "memory");
  }
}
int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int nsm = pr.multiProcessorCount; printf("SMs %d, L2 %d MB\n", nsm, pr.l2CacheSize >> 20);
  size_t maxb = (size_t)256 << 20; unsigned long long* tab; CK(cudaMalloc(&tab, maxb)); CK(cudaMemset(tab, 0, maxb));
  // ---- 1. sm -> die
  const int K = 48; int *d_lat, *d_claim; CK(cudaMalloc(&d_lat, 256 * K * 4)); CK(cudaMalloc(&d_claim, 256 * 4));
  CK(cudaMemset(d_lat, 0, 256 * K * 4)); CK(cudaMemset(d_claim, 0, 256 * 4));
  size_t stride_words = (6144 + 128) / 8;
  probe_sm<<<nsm * 16, 32, 40 * 1024>>>(tab, K, stride_words, d_lat, d_claim);
  CK(cudaDeviceSynchronize());
  std::vector<int> lat(256 * K), claim(256); CK(cudaMemcpy(lat.data(), d_lat, 256 * K * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(claim.data(), d_claim, 256 * 4, cudaMemcpyDeviceToHost));
  std::vector<int> sms; for (int s = 0; s < 256; ++s) if (claim[s]) sms.push_back(s);
  printf("SMs probed: %zu\n", sms.size());
  // per address: threshold = midpoint of min and max; vote relative to the first probed SM
  std::vector<int> vote(256, 0); int used = 0;
  for (int k = 0; k < K; ++k) {
    int mn = 1 << 30, mx = 0; for (int s : sms) { mn = std::min(mn, lat[s * K + k]); mx = std::max(mx, lat[s * K + k]); }
    if (mx - mn < 12) continue;
    int thr = (mn + mx) / 2; int ref = lat[sms[0] * K + k] > thr; ++used;
    for (int s : sms) vote[s] += ((lat[s * K + k] > thr) != ref) ? 1 : -1;
    if (k < 4) { printf("addr %d: min %d max %d | ", k, mn, mx); for (int s : sms) if (s < 24) printf("%d ", lat[s * K + k]); printf("\n"); }
  }
  std::vector<unsigned char> die(256, 0); int n1 = 0; int weak = 0;
  for (int s : sms) { die[s] = vote[s] > 0; n1 += die[s]; if (abs(vote[s]) < used / 2) ++weak; }
  printf("addresses used %d/%d; die split %d / %d; SMs with weak votes %d\n", used, K, (int)sms.size() - n1, n1, weak);
  printf("die map: "); for (int s : sms) printf("%d", die[s]); printf("\n");
  unsigned char* d_die; CK(cudaMalloc(&d_die, 256)); CK(cudaMemcpy(d_die, die.data(), 256, cudaMemcpyHostToDevice));
  // ---- 2. chunk homes
  size_t nchunks = maxb / 2048; int* d_clat; unsigned char* d_by; CK(cudaMalloc(&d_clat, nchunks * 4)); CK(cudaMalloc(&d_by, nchunks));
  probe_chunks<<<nsm * 4, 32>>>(tab, nchunks, d_clat, d_by); CK(cudaDeviceSynchronize());
  std::vector<int> clat(nchunks); std::vector<unsigned char> by(nchunks);
  CK(cudaMemcpy(clat.data(), d_clat, nchunks * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(by.data(), d_by, nchunks, cudaMemcpyDeviceToHost));
  std::vector<int> sorted(clat); std::sort(sorted.begin(), sorted.end());
  int lo = sorted[nchunks / 20], hi = sorted[nchunks - nchunks / 20], thr = (lo + hi) / 2;
  printf("chunk latency p5 %d p50 %d p95 %d -> threshold %d\n", lo, sorted[nchunks / 2], hi, thr);
  std::vector<unsigned char> home(nchunks); size_t h1 = 0; int runs = 0;
  for (size_t c = 0; c < nchunks; ++c) { home[c] = (clat[c] > thr) ? (die[by[c]] ^ 1) : die[by[c]]; h1 += home[c]; if (c && home[c] != home[c - 1]) ++runs; }
  printf("chunks homed on die1: %.3f; run changes %d of %zu (2 KB grain => ~0.5)\n", (double)h1 / nchunks, runs, nchunks);
  // finer grain check: homes of 256-byte pieces inside the first chunks
  // ---- 3. gathers
  size_t nq = 100000000; double *out, *xs, *ys; CK(cudaMalloc(&out, nq * 8)); CK(cudaMalloc(&xs, nq * 8)); CK(cudaMalloc(&ys, nq * 8));
  CK(cudaMemset(xs, 0, nq * 8)); CK(cudaMemset(ys, 0, nq * 8));
  unsigned *d_l0, *d_l1; CK(cudaMalloc(&d_l0, nchunks * 4)); CK(cudaMalloc(&d_l1, nchunks * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sizes[] = {64, 96, 128, 160, 228};
  for (int s : sizes) {
    size_t bytes = (size_t)s << 20, nrec = bytes / 32, nch = bytes / 2048;
    std::vector<unsigned> l0, l1; for (size_t c = 0; c < nch; ++c) (home[c] ? l1 : l0).push_back((unsigned)c);
    CK(cudaMemcpy(d_l0, l0.data(), l0.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_l1, l1.data(), l1.size() * 4, cudaMemcpyHostToDevice));
    printf("table %3d MiB:", s);
    for (int mode = 0; mode < 4; ++mode) {
      float best = 1e9f;
      if (mode >= 2 && (l0.empty() || l1.empty())) continue;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        gather<<<nsm * 4, 512>>>((const double*)tab, nrec, xs, ys, out, nq, mode, d_die, d_l0, d_l1, (unsigned)l0.size(), (unsigned)l1.size());
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      printf("  mode%d %.3f ms", mode, best);
    }
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
