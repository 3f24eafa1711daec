// fscl_tables.cu -- sweep-model tables on the device.
//
// Reference map (slowkoni/fscl):
//   lchoose / log_fact            sm-spline.c:18-46      (log-factorial table built on the host with libm)
//   ascprob_subsample             asc-bias.c:12-25
//   ascbias_adjust_background     asc-bias.c:27-52
//   ascbias_adjust_expect         asc-bias.c:97-109
//   compute_sweep_model_fsp       sm-spline.c:316-484    (pjh :337-351, pbk :354-403, knots :423-462)
//   estimate_spline               sm-spline.c:120-220
//   solve_linear_system           sm-spline.c:63-118
//
// Every sum runs in the reference's order and every expression keeps the reference's
// shape (gcc contracts a - b*c and a + b*c to one FMA, as nvcc does), so the spline
// solve is bit-exact for equal inputs; what differs from the host build is the last
// ulp of device exp()/log() in the knot values, which moves spline values by < 1e-12
// (tests/test_gpu_tables.py) -- far inside the 1e-9 CLR tolerance.
#include "../../include/fscl_b200.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace {

__device__ __forceinline__ double d_logfact(const double *lf, int n) {
  if (n < 0) return -DBL_MAX;
  if (n == 0 || n == 1) return 0.;
  return lf[n];
}
__device__ __forceinline__ double d_lchoose(const double *lf, int n, int k) {
  if (n == 0 && k == 0) return 0.;
  if (k > n || n == 0) return -DBL_MAX;
  return d_logfact(lf, n) - d_logfact(lf, k) - d_logfact(lf, n - k);
}
// asc-bias.c:12-25
__device__ double d_ascprob(const double *lf, int k, int d, int min_obs, int n) {
  double no_asc = 0.;
  for (int i = 0; i < min_obs; i++)
    no_asc += exp(d_lchoose(lf, k, d - i) + d_lchoose(lf, n - k, i)) + exp(d_lchoose(lf, n - k, d - i) + d_lchoose(lf, k, i));
  no_asc /= exp(d_lchoose(lf, n, d));
  return 1.0 - no_asc;
}

// asc-bias.c:27-52 (sequential sums: one thread)
__global__ void adjust_background_kernel(const double *lf, const double *bsf, int n, int asc_depth, int min_obs,
                                         double *asc, double *adj) {
  if (threadIdx.x || blockIdx.x) return;
  asc[0] = asc[n] = 0.;
  double asc_sum = 0., adj_sum = 0.;
  for (int i = 1; i < n; i++) { asc[i] = d_ascprob(lf, i, asc_depth, min_obs, n); asc_sum += asc[i]; }
  for (int i = 1; i < n; i++) asc[i] /= asc_sum;
  adj[0] = adj[n] = 0.;
  for (int i = 1; i < n; i++) { adj[i] = bsf[i] / asc[i]; adj_sum += adj[i]; }
  for (int i = 1; i < n; i++) adj[i] /= adj_sum;
}

// sm-spline.c:337-351; only h >= j is ever read (:396-401)
__global__ void pjh_kernel(const double *lf, const double *fsp, int n, double *pjh) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (h > n || h < j) return;
  double s = 0.;
  for (int i = j; i <= n; i++)
    s += fsp[i] * exp(d_lchoose(lf, i, j) + d_lchoose(lf, n - i, h - j) - d_lchoose(lf, n, h));
  pjh[(size_t)j * (n + 1) + h] = s;
}

// sm-spline.c:354-403
__global__ void pbk_kernel(const double *pjh, int n, double *pbk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (k >= n) return;
  const size_t W = n + 1;
  double s = 0.;
  const int q = b - (n - k) + 1;
  if (q > 0) s += pjh[q * W + k + 1] * (q / (double)(k + 1));
  if (b < k + 1) s += pjh[b * W + k + 1] * ((k + 1 - b) / (double)(k + 1));
  pbk[b * W + k] = s;
}

// sm-spline.c:423-462: one block per knot
__global__ void knots_kernel(const double *lf, const double *fsp, const double *pbk, int n, int spline_pts,
                             int asc_depth, int asc_min_freq, int ascbias_background_only, int include_invariant,
                             double *y /* [n+1][P] */, double *fy /* [n/2+1][P] */) {
  extern __shared__ double sh[];
  double *pk = sh, *p = sh + (n + 1), *ap = sh + 2 * (n + 1);
  const int i = blockIdx.x, P = spline_pts + 1;
  const size_t W = n + 1;
  const double log_ad_step = (4.0 - (-20.0)) / (spline_pts + 1.);
  const double log_ad = -20.0 + i * log_ad_step;
  const double ad = exp(log_ad);
  for (int k = threadIdx.x; k <= n; k += blockDim.x)      // p_kescape sm-spline.c:236-240
    pk[k] = (k == 0) ? exp(-n * ad) : exp(d_lchoose(lf, n, k) + k * log(1.0 - exp(-ad)) - (n - k) * ad);
  __syncthreads();
  for (int f = threadIdx.x; f <= n; f += blockDim.x) {
    double v = pk[n] * fsp[f];
    for (int k = 0; k < n; k++) v += pk[k] * pbk[f * W + k];
    p[f] = v;
    if (asc_depth > 0 && ascbias_background_only == 0) ap[f] = d_ascprob(lf, f, asc_depth, asc_min_freq, n);
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                  // the reference's running sums, in order
    double p_sum = 0.;
    for (int f = 0; f <= n; f++) p_sum += p[f];
    if (!include_invariant) { p_sum -= p[0] + p[n]; p[0] = p[n] = 0.; }
    for (int f = 0; f <= n; f++) p[f] /= p_sum;
    if (asc_depth > 0 && ascbias_background_only == 0) {   // asc-bias.c:97-109
      double asc_sum = 0;
      for (int f = 0; f <= n; f++) asc_sum += p[f] * ap[f];
      for (int f = 0; f <= n; f++) p[f] = p[f] * ap[f] / asc_sum;
    }
  }
  __syncthreads();
  const double lmin = log(DBL_MIN);
  for (int f = threadIdx.x; f <= n; f += blockDim.x) y[(size_t)f * P + i] = (p[f] == 0.) ? lmin : log(p[f]);
  for (int f = threadIdx.x; f <= n - f; f += blockDim.x) {
    double v;
    if (f < n - f) v = (p[f] + p[n - f] == 0.) ? lmin : log(p[f] + p[n - f]);
    else v = (p[f] == 0.) ? lmin : log(p[f]);
    fy[(size_t)f * P + i] = v;
  }
}

// estimate_spline + solve_linear_system, one thread per spline.  The 4n x 4n system is kept as a
// window of 16 columns around the diagonal (all the reference ever touches: columns i..i+7 of rows
// i..i+7), element (r, c) at m[r*16 + (c - r + 8)].
constexpr int kBW = 16;
__device__ __forceinline__ double &M(double *m, int r, int c) { return m[(size_t)r * kBW + (c - r + 8)]; }

__global__ void spline_kernel(const double *yall, int n_splines, int n /* spline_pts */, double *scratch,
                              double *coef_out /* [n_splines][n][4] */, int *flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_splines) return;
  const int N = 4 * n, P = n + 1;
  const double *y = yall + (size_t)s * P;
  double *m = scratch + (size_t)s * ((size_t)(N + 8) * kBW + 2 * (size_t)(N + 8));
  double *v = m + (size_t)(N + 8) * kBW, *b = v + (N + 8);
  for (size_t q = 0; q < (size_t)(N + 8) * kBW + 2 * (size_t)(N + 8); q++) m[q] = 0.;
  const double step = (4.0 - (-20.0)) / (n + 1.);
  auto X = [&](int k) { return -20.0 + k * step; };
  M(m, 0, 0) = 6 * X(0);
  M(m, 0, 1) = 2;
  int i, j, k;
  for (i = 1, j = 0, k = 0; k < n - 1; i += 4, j += 4, k++) {
    const double xk = X(k), x1 = X(k + 1);
    M(m, i, j) = xk * xk * xk; M(m, i, j + 1) = xk * xk; M(m, i, j + 2) = xk; M(m, i, j + 3) = 1.;
    v[i] = y[k];
    M(m, i + 1, j) = x1 * x1 * x1; M(m, i + 1, j + 1) = x1 * x1; M(m, i + 1, j + 2) = x1; M(m, i + 1, j + 3) = 1;
    v[i + 1] = y[k + 1];
    M(m, i + 2, j) = 3 * x1 * x1; M(m, i + 2, j + 1) = 2 * x1; M(m, i + 2, j + 2) = 1; M(m, i + 2, j + 3) = 0;
    M(m, i + 2, j + 4) = -3 * x1 * x1; M(m, i + 2, j + 5) = -2 * x1; M(m, i + 2, j + 6) = -1; M(m, i + 2, j + 7) = 0;
    v[i + 2] = 0;
    M(m, i + 3, j) = 6 * x1; M(m, i + 3, j + 1) = 2; M(m, i + 3, j + 2) = 0; M(m, i + 3, j + 3) = 0;
    M(m, i + 3, j + 4) = -6 * x1; M(m, i + 3, j + 5) = -2; M(m, i + 3, j + 6) = 0; M(m, i + 3, j + 7) = 0;
    v[i + 3] = 0;
  }
  {
    const double xk = X(k), xn = X(n);
    M(m, i, j) = xk * xk * xk; M(m, i, j + 1) = xk * xk; M(m, i, j + 2) = xk; M(m, i, j + 3) = 1;
    v[i] = y[k];
    M(m, i + 1, j) = xn * xn * xn; M(m, i + 1, j + 1) = xn * xn; M(m, i + 1, j + 2) = xn; M(m, i + 1, j + 3) = 1;
    v[i + 1] = y[n];
    M(m, i + 2, j) = 6 * xn; M(m, i + 2, j + 1) = 2; M(m, i + 2, j + 2) = 0; M(m, i + 2, j + 3) = 0;
    v[i + 2] = 0;
  }
  int flag = 0;
  for (i = 0; i < N; i++) {
    if (fabs(M(m, i, i)) < 1e-20) {          // sm-spline.c:67-88 inside the band (rows beyond hold zeros)
      int mx = i;
      for (j = i + 1; j < N && j < i + 8; j++)
        if (fabs(M(m, j, i)) > 0 && (mx == i || fabs(fabs(M(m, j, i)) - 1) < fabs(fabs(M(m, mx, i)) - 1))) mx = j;
      if (mx == i) { flag |= 1; break; }
      for (k = i; k < i + 8 && k < N; k++) if (k - mx < 8 && k - mx >= -8) M(m, i, k) += M(m, mx, k);
      v[i] += v[mx];
      flag |= 2;
    }
    double f = M(m, i, i);
    for (k = i; k < i + 8 && k < N; k++) M(m, i, k) /= f;
    v[i] /= f;
    for (j = i + 1; j < i + 8 && j < N; j++) {
      f = M(m, j, i);
      for (k = i; k < i + 8 && k < N; k++) M(m, j, k) = M(m, j, k) - M(m, i, k) * f;
      v[j] = v[j] - v[i] * f;
    }
  }
  for (i = N - 1; i >= 0; i--) {
    if (fabs(M(m, i, i)) < 1e-10) { b[i] = 0; flag |= 4; continue; }
    b[i] = v[i];
    for (k = i + 1; k < i + 8 && k < N; k++) b[i] -= M(m, i, k) * b[k];
  }
  double *out = coef_out + (size_t)s * N;
  for (i = 0; i < N; i++) out[i] = b[i];
  if (flag) atomicOr(flags, flag);
}

}  // namespace

extern "C" int fscl_set_error(int code, const char *msg);     // fscl_host.cu
static int tfail(int code, const std::string &m) { return fscl_set_error(code, m.c_str()); }

#define TCK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return tfail(FSCL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// compute_sweep_model_tables (sm-spline.c:486-523) + the class-indexed null model
// (scan-chromosome.c:23-37) for all depths; outputs use the class layout of fscl_problem_t.
extern "C" int fscl_build_tables(int device, int n_depths, const int32_t *depths, const double *const *fsp,
                                 int spline_pts, int asc_depth, int asc_min_freq, int ascbias_background_only,
                                 int include_invariant, double *coef_out, double *null_out, int32_t *flags_out) {
  if (n_depths <= 0 || !depths || !fsp || spline_pts < 2 || !coef_out || !null_out)
    return tfail(FSCL_ERR_ARG, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return tfail(FSCL_ERR_NODEVICE, "no CUDA device: fscl_b200 has no CPU path"); }
  TCK(cudaSetDevice(device));
  int max_n = 0;
  for (int d = 0; d < n_depths; d++) max_n = std::max(max_n, depths[d]);
  const int lf_n = std::max(max_n, asc_depth) + 2;
  std::vector<double> lf(lf_n + 1, 0.0);                 // log_fact's running sum, sm-spline.c:30-36
  for (int i = 1; i <= lf_n; i++) lf[i] = lf[i - 1] + std::log((double)i);
  const int K = spline_pts, P = K + 1, N = 4 * K;
  const size_t W = (size_t)max_n + 1;
  const int max_splines = (max_n + 1) + (max_n / 2 + 1);
  const size_t per_spline = (size_t)(N + 8) * kBW + 2 * (size_t)(N + 8);
  double *d_lf = nullptr, *d_fsp = nullptr, *d_bg = nullptr, *d_asc = nullptr, *d_pjh = nullptr, *d_pbk = nullptr,
         *d_y = nullptr, *d_scratch = nullptr, *d_coef = nullptr;
  int *d_flags = nullptr;
  TCK(cudaMalloc(&d_lf, sizeof(double) * lf.size()));
  TCK(cudaMalloc(&d_fsp, sizeof(double) * W));
  TCK(cudaMalloc(&d_bg, sizeof(double) * W));
  TCK(cudaMalloc(&d_asc, sizeof(double) * W));
  TCK(cudaMalloc(&d_pjh, sizeof(double) * W * W));
  TCK(cudaMalloc(&d_pbk, sizeof(double) * W * W));
  TCK(cudaMalloc(&d_y, sizeof(double) * (size_t)max_splines * P));
  TCK(cudaMalloc(&d_scratch, sizeof(double) * (size_t)max_splines * per_spline));
  TCK(cudaMalloc(&d_coef, sizeof(double) * (size_t)max_splines * N));
  TCK(cudaMalloc(&d_flags, sizeof(int)));
  TCK(cudaMemset(d_flags, 0, sizeof(int)));
  TCK(cudaMemcpy(d_lf, lf.data(), sizeof(double) * lf.size(), cudaMemcpyHostToDevice));
  size_t base = 0;
  for (int d = 0; d < n_depths; d++) {
    const int n = depths[d];
    const int ns = (n + 1) + (n / 2 + 1);
    TCK(cudaMemcpy(d_fsp, fsp[d], sizeof(double) * (n + 1), cudaMemcpyHostToDevice));
    const double *bg = d_fsp;
    if (asc_depth > 0) {                                  // sm-spline.c:500-505
      adjust_background_kernel<<<1, 1>>>(d_lf, d_fsp, n, asc_depth, asc_min_freq, d_asc, d_bg);
      bg = d_bg;
    }
    TCK(cudaMemset(d_pjh, 0, sizeof(double) * W * W));
    pjh_kernel<<<dim3((n + 1 + 127) / 128, n + 1), 128>>>(d_lf, bg, n, d_pjh);
    pbk_kernel<<<dim3((n + 127) / 128, n + 1), 128>>>(d_pjh, n, d_pbk);
    knots_kernel<<<P, 128, sizeof(double) * 3 * (n + 1)>>>(d_lf, bg, d_pbk, n, K, asc_depth, asc_min_freq,
                                                            ascbias_background_only, include_invariant, d_y,
                                                            d_y + (size_t)(n + 1) * P);
    spline_kernel<<<(ns + 31) / 32, 32>>>(d_y, ns, K, d_scratch, d_coef, d_flags);
    TCK(cudaGetLastError());
    TCK(cudaMemcpy(coef_out + base * N, d_coef, sizeof(double) * (size_t)ns * N, cudaMemcpyDeviceToHost));
    // null model per class with the host's libm, exactly as compute_snp_null_model
    const double *f = fsp[d];
    for (int q = 0; q <= n; q++) null_out[base + q] = std::log(f[q]);
    for (int q = 0; q <= n / 2; q++)
      null_out[base + n + 1 + q] = (q != n - q) ? std::log(f[q] + f[n - q]) : std::log(f[q]);
    base += ns;
  }
  int flags = 0;
  TCK(cudaMemcpy(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost));
  if (flags_out) *flags_out = flags;
  cudaFree(d_lf); cudaFree(d_fsp); cudaFree(d_bg); cudaFree(d_asc); cudaFree(d_pjh); cudaFree(d_pbk);
  cudaFree(d_y); cudaFree(d_scratch); cudaFree(d_coef); cudaFree(d_flags);
  if (flags & 1) return tfail(FSCL_ERR_ARG, "ill conditioned spline system (sm-spline.c:78)");
  return FSCL_OK;
}
