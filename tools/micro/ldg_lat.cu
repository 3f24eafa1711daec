// Dependent-load latency (L1 hit) of 64-, 128- and 256-bit global loads, one warp per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const double *tab, long long *cyc, double *sink, int iters) {
  unsigned idx = threadIdx.x & 7;
  double acc = 0;
  // warm L1
  for (int i = 0; i < 512; i++) acc += tab[i * 4];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    const double *p = tab + (size_t)(idx & 511) * 4;
    double a, b, c, d;
    if (MODE == 0) asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
    else if (MODE == 1) { asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p)); c = d = 0; }
    else { asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(a) : "l"(p)); b = c = d = 0; }
    acc += b + c + d;
    idx = (unsigned)__double2int_rn(a) + idx + 1;   // next address depends on the loaded value (table holds zeros)
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = t1 - t0; *sink = acc + idx; }
}
int main() {
  double *tab, *sink; long long *cyc;
  cudaMalloc(&tab, 512 * 32); cudaMemset(tab, 0, 512 * 32);
  cudaMalloc(&sink, 8); cudaMallocManaged(&cyc, 8);
  const int iters = 20000;
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      if (mode == 0) k<0><<<1, 32>>>(tab, cyc, sink, iters);
      if (mode == 1) k<1><<<1, 32>>>(tab, cyc, sink, iters);
      if (mode == 2) k<2><<<1, 32>>>(tab, cyc, sink, iters);
      cudaDeviceSynchronize();
    }
    printf("mode %d (0=256b,1=128b,2=64b): %.1f cycles per dependent load (includes ~12 cycles of address arithmetic)\n", mode, (double)*cyc / iters);
  }
  return 0;
}
