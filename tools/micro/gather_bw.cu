// Microbenchmark: throughput of gathering class-random 32-byte rows (one per lane, 32 distinct rows per warp,
// 1 MB table resident in L2, the near-field pattern of scan_cells_kernel) with
//   mode 0: ld.global.nc.v4.f64 into registers (what the kernel does), DEPTH rows in flight per lane
//   mode 1: cp.async (LDGSTS) 2 x 16 B per row into shared memory, DEPTH rows per lane per stage, two stages,
//           then 2 x LDS.128
// 2 CTAs x 8 warps per SM as in the kernel.  Prints SM-cycles per warp-wide row gather.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
constexpr int kRows = 32768;     // 1 MB
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(256, 2) k(const double *tab, double *out, int iters, long long *cyc) {
  extern __shared__ double2 dyn[];
  double2 (*h0)[DEPTH][256] = reinterpret_cast<double2 (*)[DEPTH][256]>(dyn);
  double2 (*h1)[DEPTH][256] = reinterpret_cast<double2 (*)[DEPTH][256]>(dyn + 2 * DEPTH * 256);
  unsigned rng = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  auto next_row = [&]() { rng = rng * 1664525u + 1013904223u; return (rng >> 8) % kRows; };
  double s = 0;
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int i = 0; i < iters; i++) {
      double a[DEPTH], b[DEPTH], c[DEPTH], d[DEPTH];
#pragma unroll
      for (int u = 0; u < DEPTH; u++) {
        const double *p = tab + (size_t)next_row() * 4;
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[u]), "=d"(b[u]), "=d"(c[u]), "=d"(d[u]) : "l"(p));
      }
#pragma unroll
      for (int u = 0; u < DEPTH; u++) s += fma(a[u], b[u], fma(c[u], d[u], 1.0));
    }
  } else {
    auto issue = [&](int st) {
#pragma unroll
      for (int u = 0; u < DEPTH; u++) {
        const double *p = tab + (size_t)next_row() * 4;
        __pipeline_memcpy_async(&h0[st][u][threadIdx.x], p, 16);
        __pipeline_memcpy_async(&h1[st][u][threadIdx.x], p + 2, 16);
      }
      __pipeline_commit();
    };
    issue(0);
    for (int i = 0; i < iters; i++) {
      const int st = i & 1;
      if (i + 1 < iters) issue(st ^ 1);
      __pipeline_wait_prior(i + 1 < iters ? 1 : 0);
#pragma unroll
      for (int u = 0; u < DEPTH; u++) {
        const double2 x = h0[st][u][threadIdx.x], y = h1[st][u][threadIdx.x];
        s += fma(x.x, x.y, fma(y.x, y.y, 1.0));
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int DEPTH>
void run(const double *tab, double *out, long long *cyc) {
  const int iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    const int smem = MODE ? 2 * 2 * DEPTH * 256 * 16 : 0;
    cudaFuncSetAttribute(k<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE, DEPTH><<<148 * 2, 256, smem>>>(tab, out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  // per SM: 16 warps x iters x DEPTH warp-wide row gathers in ms
  const double warp_rows = 16.0 * iters * DEPTH;
  printf("mode %d (0 = LDG.256 to registers, 1 = cp.async to shared + LDS) depth %d: %.2f SM-cycles per warp-row (clock64 %.2f), %.1f G rows/s per GPU\n",
         MODE, DEPTH, ms * 1e-3 * 1.965e9 / warp_rows, (double)*cyc / (8.0 * iters * DEPTH) / 2.0, 148 * warp_rows * 32 / (ms * 1e-3) / 1e9);
}
int main() {
  double *tab, *out; long long *cyc;
  cudaMalloc(&tab, (size_t)kRows * 32); cudaMemset(tab, 0, (size_t)kRows * 32);
  cudaMalloc(&out, 148 * 2 * 256 * 8); cudaMallocManaged(&cyc, 8);
  run<0, 2>(tab, out, cyc); run<0, 4>(tab, out, cyc); run<0, 8>(tab, out, cyc);
  run<1, 2>(tab, out, cyc); run<1, 4>(tab, out, cyc); run<1, 6>(tab, out, cyc);
  return 0;
}
