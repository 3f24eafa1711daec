// Microbenchmark: L1 -> register-file delivery cost of 256-bit and 128-bit loads as a
// function of how many distinct 32-byte rows the 32 lanes of a warp ask for.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int MODE>
__global__ void k(const double *tab, double *out, int iters, int distinct, long long *cyc) {
  const int lane = threadIdx.x & 31;
  // rows chosen inside a 16 KB region so that everything is an L1 hit
  unsigned row = (lane % distinct) * 7 + (threadIdx.x >> 5);
  double s = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    double a[4], b[4], c[4], d[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double *p = tab + ((row + u * 13 + i) & 511) * 4;
      if (MODE == 0) ld256(p, a[u], b[u], c[u], d[u]);
      else if (MODE == 1) { double2 x = __ldg((const double2 *)p), y = __ldg((const double2 *)p + 1); a[u] = x.x; b[u] = x.y; c[u] = y.x; d[u] = y.y; }
      else { a[u] = __ldg(p); b[u] = c[u] = d[u] = 1.0; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) s += a[u] + b[u] + c[u] + d[u];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *tab, *out; long long *cyc;
  cudaMalloc(&tab, 512 * 32); cudaMemset(tab, 0, 512 * 32);
  cudaMalloc(&out, 148 * 4 * 512 * 8); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  for (int mode = 0; mode < 3; mode++)
    for (int distinct : {1, 2, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) k<0><<<148 * 2, 512>>>(tab, out, iters, distinct, cyc);
        if (mode == 1) k<1><<<148 * 2, 512>>>(tab, out, iters, distinct, cyc);
        if (mode == 2) k<2><<<148 * 2, 512>>>(tab, out, iters, distinct, cyc);
        cudaDeviceSynchronize();
      }
      // 2 CTAs x 16 warps per SM, each warp issues iters*4 row loads
      double per = (double)*cyc / (iters * 4.0 * 32.0);
      printf("mode %d (0=256b,1=2x128b,2=64b) distinct rows/warp %2d : %.2f SM-cycles per warp-row-load\n", mode, distinct, per);
    }
  return 0;
}
