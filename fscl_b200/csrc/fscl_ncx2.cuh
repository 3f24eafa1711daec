// fscl_ncx2.cuh -- survival function of the non-central chi-square, shared by the host entry points
// (fscl_ncx2_sf, fscl_project_pvalues) and the device projection kernel (project_kernel).
// The p-value projection is described at README.md:11 of the reference and never implemented there
// (SURVEY.md, fact 3): this is a definition, see DESIGN.md 3.6.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define FSCL_HD __host__ __device__
#else
#define FSCL_HD
#endif

namespace fscl {

// regularised upper incomplete gamma Q(a, x), a > 0, x >= 0 (series below a+1, Lentz continued fraction above)
FSCL_HD inline double gamma_q(double a, double x) {
  if (x <= 0.0) return 1.0;
  const double lg = lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 10000; n++) {
      ap += 1.0; del *= x / ap; sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return 1.0 - sum * exp(-x + a * log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b; if (fabs(d) < tiny) d = tiny;
    c = b + an / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  return exp(-x + a * log(x) - lg) * h;
}

// survival function of the non-central chi-square: Poisson(lambda/2) mixture of central ones,
// summed outwards from the Poisson mode so that tiny tails keep their relative accuracy
FSCL_HD inline double ncx2_sf(double x, double df, double lambda) {
  if (x <= 0.0) return 1.0;
  if (lambda <= 0.0) return gamma_q(0.5 * df, 0.5 * x);
  const double mu = 0.5 * lambda;
  const int mode = (int)mu;
  double total = 0.0;
  double w = exp(-mu + mode * log(mu) - lgamma(mode + 1.0));
  double wi = w;
  for (int i = mode; i < mode + 100000; i++) {          // upwards
    const double term = wi * gamma_q(0.5 * df + i, 0.5 * x);
    total += term;
    if (i > mode && wi < 1e-18 * w && term < 1e-17 * total) break;
    wi *= mu / (i + 1.0);
  }
  wi = w;
  for (int i = mode - 1; i >= 0; i--) {                  // downwards
    wi *= (i + 1.0) / mu;
    const double term = wi * gamma_q(0.5 * df + i, 0.5 * x);
    total += term;
    if (wi < 1e-18 * w && term < 1e-17 * total) break;
  }
  return total > 1.0 ? 1.0 : total;
}

}  // namespace fscl
