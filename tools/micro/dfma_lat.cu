// FP64 FMA pipe: cycles per warp-DFMA on one SM as a function of resident warps and
// independent chains per warp (how much ILP x TLP the pipe needs to reach 64 FMA/clk/SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH>
__global__ void k(double *out, long long *cyc, int iters) {
  double a[NCH];
#pragma unroll
  for (int c = 0; c < NCH; c++) a[c] = threadIdx.x * 1e-9 + c;
  const double m = 0.999999, b = 1e-7;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) a[c] = fma(a[c], m, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < NCH; c++) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 2048 * 8); cudaMallocManaged(&cyc, 8);
  const int iters = 20000;
  printf("warps/SM chains  cycles per warp-DFMA (SM level)   DFMA lanes/clk/SM\n");
  for (int warps : {4, 8, 16, 32, 64})
    for (int nch : {1, 2, 4, 8}) {
      for (int rep = 0; rep < 2; rep++) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32, blocks = 148 * (warps * 32 / threads);
        if (nch == 1) k<1><<<blocks, threads>>>(out, cyc, iters);
        if (nch == 2) k<2><<<blocks, threads>>>(out, cyc, iters);
        if (nch == 4) k<4><<<blocks, threads>>>(out, cyc, iters);
        if (nch == 8) k<8><<<blocks, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      double per = (double)*cyc / ((double)iters * nch * warps);
      printf("%5d %6d   %8.3f   %8.1f\n", warps, nch, per, 32.0 / per);
    }
  return 0;
}
