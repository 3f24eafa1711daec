// Microbenchmark (round 2): what does a 32-byte-per-lane row load cost when only some lanes of the warp take part?
//   mode 0  ld.global.nc.v4.f64, lanes 0 .. a-1 active (contiguous), all L1 hits
//   mode 1  same, every (32/a)-th lane active (spread)
//   mode 2  two ld.shared.v2.f64 per row, all 32 lanes, `a` distinct rows per warp (broadcast among the lanes of a row)
//   mode 3  ld.global.nc.v2.f64 x 2 (two 128-bit loads), lanes 0 .. a-1 active
// Answers whether lane-private row reuse (reload only on a (class, knot) change) can beat the 8 cycles per
// warp-wide 256-bit load that r01_ldg_bw_microbench.txt measured.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ld128x2(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2+16];" : "=d"(c), "=d"(d) : "l"(p));
}
template <int MODE>
__global__ void k(const double *tab, double *out, int iters, int a, long long *cyc) {
  __shared__ double sh[512 * 4];
  for (int i = threadIdx.x; i < 512 * 4; i += blockDim.x) sh[i] = tab[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  bool act = true;
  if (MODE == 0 || MODE == 3) act = lane < a;
  if (MODE == 1) act = (lane % (32 / a)) == 0;
  unsigned row = (MODE == 2 ? (lane % a) * 7 : lane * 7) + (threadIdx.x >> 5);
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    double x[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned r = (row + u * 13 + i) & 511;
      if (MODE == 2) {
        const double2 p = *reinterpret_cast<const double2 *>(sh + r * 4), q = *reinterpret_cast<const double2 *>(sh + r * 4 + 2);
        x[u][0] = p.x; x[u][1] = p.y; x[u][2] = q.x; x[u][3] = q.y;
      } else if (act) {
        if (MODE == 3) ld128x2(tab + r * 4, x[u][0], x[u][1], x[u][2], x[u][3]);
        else ld256(tab + r * 4, x[u][0], x[u][1], x[u][2], x[u][3]);
      } else { x[u][0] = x[u][1] = x[u][2] = x[u][3] = 0.0; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) { s0 += x[u][0]; s1 += x[u][1]; s2 += x[u][2]; s3 += x[u][3]; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *tab, *out; long long *cyc;
  cudaMalloc(&tab, 512 * 32); cudaMemset(tab, 0, 512 * 32);
  cudaMalloc(&out, 148 * 4 * 512 * 8); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  for (int mode = 0; mode < 4; mode++)
    for (int a : {1, 2, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) k<0><<<148 * 2, 512>>>(tab, out, iters, a, cyc);
        if (mode == 1) k<1><<<148 * 2, 512>>>(tab, out, iters, a, cyc);
        if (mode == 2) k<2><<<148 * 2, 512>>>(tab, out, iters, a, cyc);
        if (mode == 3) k<3><<<148 * 2, 512>>>(tab, out, iters, a, cyc);
        cudaDeviceSynchronize();
      }
      // 2 CTAs x 16 warps per SM, each warp issues iters*4 row loads
      const double per = (double)*cyc / (iters * 4.0 * 32.0);
      const char *nm[] = {"LDG.256, lanes 0..a-1 active", "LDG.256, a lanes spread over the warp", "2 x LDS.128, a distinct rows (broadcast)",
                          "2 x LDG.128, lanes 0..a-1 active"};
      printf("mode %d (%s) a = %2d : %.2f SM-cycles per warp-wide row load\n", mode, nm[mode], a, per);
    }
  return 0;
}
