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

// Everything below runs for ALL sample depths in one launch per stage (blockIdx.z / .y = depth); the
// per-depth inputs and outputs sit in packed arrays addressed through a descriptor table.
struct DepthDesc {
  int n;                 // sample depth
  int spline_base;       // first class (= spline) of this depth; y / coef offsets follow from it
  long long off_vec;     // offset of this depth's (n + 1)-vectors: fsp, bg, asc
  long long off_sq;      // offset of its (n + 1) x (n + 1) matrices: pjh, pbk
};

// asc-bias.c:27-52: the ascertainment probabilities in parallel, the two normalisations by one thread in the
// reference's summation order
__global__ void __launch_bounds__(128) adjust_background_kernel(const double *lf, const DepthDesc *dd,
                                                                const double *fsp_all, int asc_depth, int min_obs,
                                                                double *asc_all, double *bg_all) {
  const DepthDesc D = dd[blockIdx.x];
  const int n = D.n;
  const double *bsf = fsp_all + D.off_vec;
  double *asc = asc_all + D.off_vec, *adj = bg_all + D.off_vec;
  for (int i = 1 + threadIdx.x; i < n; i += blockDim.x) asc[i] = d_ascprob(lf, i, asc_depth, min_obs, n);
  __syncthreads();
  if (threadIdx.x == 0) {
    asc[0] = asc[n] = 0.;
    double asc_sum = 0., adj_sum = 0.;
    for (int i = 1; i < n; i++) asc_sum += asc[i];
    for (int i = 1; i < n; i++) asc[i] /= asc_sum;
    adj[0] = adj[n] = 0.;
    for (int i = 1; i < n; i++) { adj[i] = bsf[i] / asc[i]; adj_sum += adj[i]; }
    for (int i = 1; i < n; i++) adj[i] /= adj_sum;
  }
}

// sm-spline.c:337-351; only h >= j is ever read (:396-401)
__global__ void pjh_kernel(const double *lf, const DepthDesc *dd, const double *bg_all, double *pjh_all) {
  const DepthDesc D = dd[blockIdx.z];
  const int n = D.n;
  const int h = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (j > n || h > n || h < j) return;
  const double *fsp = bg_all + D.off_vec;
  double s = 0.;
  for (int i = j; i <= n; i++)
    s += fsp[i] * exp(d_lchoose(lf, i, j) + d_lchoose(lf, n - i, h - j) - d_lchoose(lf, n, h));
  pjh_all[D.off_sq + (size_t)j * (n + 1) + h] = s;
}

// sm-spline.c:354-403
__global__ void pbk_kernel(const DepthDesc *dd, const double *pjh_all, double *pbk_all) {
  const DepthDesc D = dd[blockIdx.z];
  const int n = D.n;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (k >= n || b > n) return;
  const size_t W = n + 1;
  const double *pjh = pjh_all + D.off_sq;
  double s = 0.;
  const int q = b - (n - k) + 1;
  if (q > 0) s += pjh[q * W + k + 1] * (q / (double)(k + 1));
  if (b < k + 1) s += pjh[b * W + k + 1] * ((k + 1 - b) / (double)(k + 1));
  pbk_all[D.off_sq + b * W + k] = s;
}

// sm-spline.c:423-462: one block per (knot, depth)
__global__ void knots_kernel(const double *lf, const DepthDesc *dd, const double *bg_all, const double *pbk_all,
                             int spline_pts, int asc_depth, int asc_min_freq, int ascbias_background_only,
                             int include_invariant, double *y_all /* [spline][P] */) {
  extern __shared__ double sh[];
  const DepthDesc D = dd[blockIdx.y];
  const int n = D.n;
  double *pk = sh, *p = sh + (n + 1), *ap = sh + 2 * (n + 1);
  const int i = blockIdx.x, P = spline_pts + 1;
  const size_t W = n + 1;
  const double *fsp = bg_all + D.off_vec, *pbk = pbk_all + D.off_sq;
  double *y = y_all + (size_t)D.spline_base * P, *fy = y + (size_t)(n + 1) * P;
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

// estimate_spline + solve_linear_system (sm-spline.c:63-220), ONE WARP PER SPLINE.  The 4n x 4n system is
// kept as a window of 16 columns around the diagonal (all the reference ever touches: columns i..i+7 of rows
// i..i+7), element (r, c) at m[r*16 + (c - r + 8)].  One elimination step of the reference -- normalise the
// pivot row, then subtract it from the 7 rows below over 8 columns -- is 8 + 56 + 7 independent
// element updates: the lanes of the warp take them in parallel, every element with the reference's own
// expression (a / f, a - b * f: one division or one fused multiply-subtract, as gcc emits them), so the
// result equals the sequential elimination bit for bit for equal inputs.  Only the steps are sequential.
constexpr int kBW = 16;
__device__ __forceinline__ double &M(double *m, int r, int c) { return m[(size_t)r * kBW + (c - r + 8)]; }

__global__ void __launch_bounds__(128) spline_kernel(const double *yall, int n_splines, int n /* spline_pts */,
                                                     double *scratch, double *coef_out /* [n_splines][n][4] */,
                                                     int *flags) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= n_splines) return;
  const unsigned full = 0xFFFFFFFFu;
  const int N = 4 * n, P = n + 1;
  const double *y = yall + (size_t)s * P;
  const size_t per = (size_t)(N + 8) * kBW + 2 * (size_t)(N + 8);
  double *m = scratch + (size_t)s * per;
  double *v = m + (size_t)(N + 8) * kBW, *b = v + (N + 8);
  for (size_t q = lane; q < per; q += 32) m[q] = 0.;
  __syncwarp();
  const double step = (4.0 - (-20.0)) / (n + 1.);
  auto X = [&](int k) { return -20.0 + k * step; };
  if (lane == 0) { M(m, 0, 0) = 6 * X(0); M(m, 0, 1) = 2; }
  for (int k = lane; k < n - 1; k += 32) {                 // sm-spline.c:137-180: four rows per interior knot
    const int i = 1 + 4 * k, j = 4 * k;
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
  if (lane == 0) {                                          // the last interval, sm-spline.c:182-197
    const int k = n - 1, i = 1 + 4 * k, j = 4 * k;
    const double xk = X(k), xn = X(n);
    M(m, i, j) = xk * xk * xk; M(m, i, j + 1) = xk * xk; M(m, i, j + 2) = xk; M(m, i, j + 3) = 1;
    v[i] = y[k];
    M(m, i + 1, j) = xn * xn * xn; M(m, i + 1, j + 1) = xn * xn; M(m, i + 1, j + 2) = xn; M(m, i + 1, j + 3) = 1;
    v[i + 1] = y[n];
    M(m, i + 2, j) = 6 * xn; M(m, i + 2, j + 1) = 2; M(m, i + 2, j + 2) = 0; M(m, i + 2, j + 3) = 0;
    v[i + 2] = 0;
  }
  __syncwarp();
  int flag = 0;
  bool dead = false;
  for (int i = 0; i < N && !dead; i++) {
    double piv = M(m, i, i);                               // every lane reads the pivot
    if (fabs(piv) < 1e-20) {                               // sm-spline.c:67-88 inside the band (rows beyond hold zeros): rare, serial
      if (lane == 0) {
        int mx = i;
        for (int j = i + 1; j < N && j < i + 8; j++)
          if (fabs(M(m, j, i)) > 0 && (mx == i || fabs(fabs(M(m, j, i)) - 1) < fabs(fabs(M(m, mx, i)) - 1))) mx = j;
        if (mx == i) flag |= 1;
        else {
          for (int k = i; k < i + 8 && k < N; k++) if (k - mx < 8 && k - mx >= -8) M(m, i, k) += M(m, mx, k);
          v[i] += v[mx];
          flag |= 2;
        }
      }
      __syncwarp();
      flag = __shfl_sync(full, flag, 0);
      if (flag & 1) { dead = true; break; }
      piv = M(m, i, i);
    }
    // lanes 0..7: column i + lane of the pivot row; lanes 0..6 also hold the multiplier of row i + 1 + lane
    const int c8 = lane & 7;
    double prow = 0., fcol = 0.;
    if (lane < 8 && i + lane < N) prow = M(m, i, i + lane) / piv;
    if (lane < 7 && i + 1 + lane < N) fcol = M(m, i + 1 + lane, i);
    if (lane < 8 && i + lane < N) M(m, i, i + lane) = prow;
    double vi = 0.;
    if (lane == 0) { vi = v[i] / piv; v[i] = vi; }
    vi = __shfl_sync(full, vi, 0);
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      const int pr = lane + 32 * pass, r = pr >> 3;        // pair (row offset r + 1, column offset c8), 56 pairs
      const double f = __shfl_sync(full, fcol, r < 7 ? r : 0);
      const double pc = __shfl_sync(full, prow, c8);
      if (r < 7 && i + 1 + r < N && i + c8 < N) {
        double &e = M(m, i + 1 + r, i + c8);
        e = e - pc * f;
      }
    }
    if (lane < 7 && i + 1 + lane < N) v[i + 1 + lane] = v[i + 1 + lane] - vi * fcol;
    __syncwarp();
  }
  if (lane == 0 && !dead) {                                // back substitution: a dependent chain, sm-spline.c:105-116
    for (int i = N - 1; i >= 0; i--) {
      if (fabs(M(m, i, i)) < 1e-10) { b[i] = 0; flag |= 4; continue; }
      b[i] = v[i];
      for (int k = i + 1; k < i + 8 && k < N; k++) b[i] -= M(m, i, k) * b[k];
    }
  }
  __syncwarp();
  double *out = coef_out + (size_t)s * N;
  if (!dead) for (int i = lane; i < N; i += 32) out[i] = b[i];
  if (lane == 0 && flag) atomicOr(flags, flag);
}

}  // namespace

extern "C" int fscl_set_error(int code, const char *msg);     // fscl_host.cu
static int tfail(int code, const std::string &m) { return fscl_set_error(code, m.c_str()); }

#define TCK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return tfail(FSCL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace {
struct DeviceBlock {               // one allocation for the whole build, released on every path
  char *p = nullptr;
  ~DeviceBlock() { if (p) cudaFree(p); }
};
}  // namespace

// compute_sweep_model_tables (sm-spline.c:486-523) + the class-indexed null model
// (scan-chromosome.c:23-37) for all depths; outputs use the class layout of fscl_problem_t.
// One device allocation, five launches (every stage covers all depths), one copy back.
extern "C" int fscl_build_tables(int device, int n_depths, const int32_t *depths, const double *const *fsp,
                                 int spline_pts, int asc_depth, int asc_min_freq, int ascbias_background_only,
                                 int include_invariant, double *coef_out, double *null_out, int32_t *flags_out) {
  if (n_depths <= 0 || !depths || !fsp || spline_pts < 2 || !coef_out || !null_out)
    return tfail(FSCL_ERR_ARG, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return tfail(FSCL_ERR_NODEVICE, "no CUDA device: fscl_b200 has no CPU path"); }
  TCK(cudaSetDevice(device));
  int max_n = 0;
  for (int d = 0; d < n_depths; d++) {
    if (depths[d] < 2) return tfail(FSCL_ERR_ARG, "sample depth below 2");
    max_n = std::max(max_n, depths[d]);
  }
  const int lf_n = std::max(max_n, asc_depth) + 2;
  std::vector<double> lf(lf_n + 1, 0.0);                 // log_fact's running sum, sm-spline.c:30-36
  for (int i = 1; i <= lf_n; i++) lf[i] = lf[i - 1] + std::log((double)i);
  const int K = spline_pts, P = K + 1, N = 4 * K;
  std::vector<DepthDesc> desc(n_depths);
  long long vec = 0, sq = 0;
  int n_splines = 0;
  for (int d = 0; d < n_depths; d++) {
    const int n = depths[d];
    desc[d] = DepthDesc{n, n_splines, vec, sq};
    vec += n + 1; sq += (long long)(n + 1) * (n + 1);
    n_splines += (n + 1) + (n / 2 + 1);
  }
  std::vector<double> fsp_all((size_t)vec);
  for (int d = 0; d < n_depths; d++) memcpy(fsp_all.data() + desc[d].off_vec, fsp[d], sizeof(double) * (size_t)(depths[d] + 1));
  const size_t per_spline = (size_t)(N + 8) * kBW + 2 * (size_t)(N + 8);
  // carve one block: descriptors | lf | fsp | bg | asc | pjh | pbk | y | coef | scratch | flags
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_desc = carve(sizeof(DepthDesc) * (size_t)n_depths), o_lf = carve(sizeof(double) * lf.size());
  const size_t o_fsp = carve(sizeof(double) * (size_t)vec), o_bg = carve(sizeof(double) * (size_t)vec);
  const size_t o_asc = carve(sizeof(double) * (size_t)vec);
  const size_t o_pjh = carve(sizeof(double) * (size_t)sq), o_pbk = carve(sizeof(double) * (size_t)sq);
  const size_t o_y = carve(sizeof(double) * (size_t)n_splines * P), o_coef = carve(sizeof(double) * (size_t)n_splines * N);
  const size_t o_scr = carve(sizeof(double) * (size_t)n_splines * per_spline), o_flags = carve(sizeof(int));
  DeviceBlock blk;
  TCK(cudaMalloc((void **)&blk.p, off));
  DepthDesc *d_desc = (DepthDesc *)(blk.p + o_desc);
  double *d_lf = (double *)(blk.p + o_lf), *d_fsp = (double *)(blk.p + o_fsp), *d_bg = (double *)(blk.p + o_bg);
  double *d_asc = (double *)(blk.p + o_asc), *d_pjh = (double *)(blk.p + o_pjh), *d_pbk = (double *)(blk.p + o_pbk);
  double *d_y = (double *)(blk.p + o_y), *d_coef = (double *)(blk.p + o_coef), *d_scratch = (double *)(blk.p + o_scr);
  int *d_flags = (int *)(blk.p + o_flags);
  TCK(cudaMemsetAsync(d_flags, 0, sizeof(int)));
  TCK(cudaMemsetAsync(d_pjh, 0, sizeof(double) * (size_t)sq));
  TCK(cudaMemcpyAsync(d_desc, desc.data(), sizeof(DepthDesc) * (size_t)n_depths, cudaMemcpyHostToDevice));
  TCK(cudaMemcpyAsync(d_lf, lf.data(), sizeof(double) * lf.size(), cudaMemcpyHostToDevice));
  TCK(cudaMemcpyAsync(d_fsp, fsp_all.data(), sizeof(double) * (size_t)vec, cudaMemcpyHostToDevice));
  const double *bg = d_fsp;
  if (asc_depth > 0) {                                    // sm-spline.c:500-505
    adjust_background_kernel<<<n_depths, 128>>>(d_lf, d_desc, d_fsp, asc_depth, asc_min_freq, d_asc, d_bg);
    bg = d_bg;
  }
  const unsigned gx = (unsigned)(max_n + 1 + 127) / 128;
  pjh_kernel<<<dim3(gx, max_n + 1, n_depths), 128>>>(d_lf, d_desc, bg, d_pjh);
  pbk_kernel<<<dim3(gx, max_n + 1, n_depths), 128>>>(d_desc, d_pjh, d_pbk);
  knots_kernel<<<dim3(P, n_depths), 128, sizeof(double) * 3 * (max_n + 1)>>>(d_lf, d_desc, bg, d_pbk, K, asc_depth, asc_min_freq,
                                                                            ascbias_background_only, include_invariant, d_y);
  spline_kernel<<<(n_splines + 3) / 4, 128>>>(d_y, n_splines, K, d_scratch, d_coef, d_flags);
  TCK(cudaGetLastError());
  TCK(cudaMemcpy(coef_out, d_coef, sizeof(double) * (size_t)n_splines * N, cudaMemcpyDeviceToHost));
  int flags = 0;
  TCK(cudaMemcpy(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost));
  // null model per class with the host's libm, exactly as compute_snp_null_model
  size_t base = 0;
  for (int d = 0; d < n_depths; d++) {
    const int n = depths[d];
    const double *f = fsp[d];
    for (int q = 0; q <= n; q++) null_out[base + q] = std::log(f[q]);
    for (int q = 0; q <= n / 2; q++)
      null_out[base + n + 1 + q] = (q != n - q) ? std::log(f[q] + f[n - q]) : std::log(f[q]);
    base += (size_t)(n + 1) + (n / 2 + 1);
  }
  if (flags_out) *flags_out = flags;
  if (flags & 1) return tfail(FSCL_ERR_ARG, "ill conditioned spline system (sm-spline.c:78)");
  return FSCL_OK;
}
