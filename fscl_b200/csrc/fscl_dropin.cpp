// fscl_dropin.cpp -- reference-signature entry points over the C-ABI of libfscl_b200.so.
// See include/fscl_dropin.h and INTEGRATION.md.  Error convention of the reference
// (logmsg.c:34, kmacros.h:13-20): message on stderr, then exit(1).
#include "fscl_dropin.h"
#include "fscl_b200.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/time.h>
#include <thread>
#include <unistd.h>
#include <vector>

extern "C" void logmsg(int priority, volatile char *s, ...) __attribute__((weak));   // reference logmsg.o, if linked
// globals of fscl.c:33-34 that the reference's scan-chromosome.c imports for its SIGINT dump
// (scan-chromosome.c:550-552); weak, so the library also loads without the reference's main()
extern "C" char *output_fname __attribute__((weak));
extern "C" char *prepend_label __attribute__((weak));
extern "C" void scan_output(char *output_fname, scan_t *so, int maximum_only, int n_permute, char *prepend_label);

// the reference's host table build when its sm-spline.o is linked (fscl_b200/host/Makefile renames the symbol)
extern "C" sm_ptable_t *ref_compute_sweep_model_tables(scan_t *, double **, int, int, int, int) __attribute__((weak));
extern "C" int spline_pts __attribute__((weak));       // fscl.c:34

namespace {

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "\nfscl_b200: %s: %s\n", what, fscl_last_error());
  exit(1);
}

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// One resident problem.  scan_chromosome() always rebuilds it (it runs once per scan object; on the ms
// path the reference frees and re-allocates the scan_t per block, fscl.c:298-311, and malloc hands the
// same address back, so the address alone says nothing); scan_permute() / search_maxalpha() reuse it
// only when every input they were bound with is unchanged.
struct Bound {
  const scan_t *scan = nullptr;
  const snp_t *snps = nullptr;
  const sm_ptable_t *sm_p = nullptr;
  int n_snps = 0, eval_range = 0, bp_resl = 0;
  unsigned long long digest = 0;                 // of (pos, class) at a few hundred sites
  std::vector<void *> ctx;                       // one per GPU (FSCL_B200_GPUS), ctx[0] answers
  int large_grid_sp = 0;
  // results of the last permutation test, for scan_output's companion file
  std::vector<double> proj_lambda, proj_p;
  std::vector<float> clr_packed;                 // stored permuted CLRs, cell i at clr_off[i]
  float *clr_dense = nullptr;                    // or one FSCL_NULL_SAVE row per cell (sequential streams, checkpoints)
} g_bound;

void release() {
  for (void *c : g_bound.ctx) if (c) fscl_destroy(c);
  free(g_bound.clr_dense);
  g_bound = Bound();
}

unsigned long long digest_of(const scan_t *so) {
  unsigned long long h = 1469598103934665603ull;
  const int n = so->n_snps, step = n > 512 ? n / 512 : 1;
  for (int i = 0; i < n; i += step) {
    const snp_t &s = so->snps[i];
    const unsigned long long v[3] = {(unsigned long long)(unsigned)s.pos, (unsigned long long)(unsigned)s.obs_freq,
                                     ((unsigned long long)(unsigned)s.depth_p << 1) | (unsigned)(s.folded != 0)};
    for (unsigned long long x : v) { h ^= x; h *= 1099511628211ull; }
  }
  return h;
}

bool bound_matches(const scan_t *so, const sm_ptable_t *sm_p, int eval_range, int bp_resl) {
  return !g_bound.ctx.empty() && g_bound.scan == so && g_bound.snps == so->snps && g_bound.n_snps == so->n_snps &&
         g_bound.sm_p == sm_p && g_bound.eval_range == eval_range && g_bound.bp_resl == bp_resl &&
         g_bound.digest == digest_of(so);
}

// tables the drop-in built itself on the device (compute_sweep_model_tables below): flat, class-major
struct DeviceTables {
  const sm_ptable_t *sm_p = nullptr;
  std::vector<double> coef;
  int n_cls = 0, K = 0;
} g_dev_tables;

// sm_ptable_t pointer forest + snp_t array -> the flattened problem of fscl_b200.h
void bind(scan_t *so, sm_ptable_t *sm_p, int eval_range, int bp_resl, bool force) {
  if (!force && bound_matches(so, sm_p, eval_range, bp_resl)) return;
  release();
  const int D = so->n_depths, N = so->n_snps;
  std::vector<long> base(D + 1, 0);
  for (int d = 0; d < D; d++) base[d + 1] = base[d] + (so->sample_depths[d] + 1) + (so->sample_depths[d] / 2 + 1);
  const int n_cls = (int)base[D];
  const int K = sm_p[0].spline_func[0]->n;
  std::vector<double> coef_local;
  const double *coef = nullptr;
  if (g_dev_tables.sm_p == sm_p && g_dev_tables.n_cls == n_cls && g_dev_tables.K == K) {
    coef = g_dev_tables.coef.data();              // already flat
  } else {
    coef_local.resize((size_t)n_cls * K * 4);
    for (int d = 0; d < D; d++) {
      const int n = so->sample_depths[d];
      for (int f = 0; f <= n; f++)
        memcpy(&coef_local[(size_t)(base[d] + f) * K * 4], sm_p[d].spline_func[f]->coef[0], sizeof(double) * K * 4);
      for (int f = 0; f <= n / 2; f++)
        memcpy(&coef_local[(size_t)(base[d] + n + 1 + f) * K * 4], sm_p[d].fspline_func[f]->coef[0], sizeof(double) * K * 4);
    }
    coef = coef_local.data();
  }
  std::vector<double> null_tab(n_cls, 0.0);
  std::vector<int32_t> chr(N), pos(N), cls(N);
  for (int i = 0; i < N; i++) {
    const snp_t &s = so->snps[i];
    chr[i] = s.chr; pos[i] = s.pos;
    const int n = so->sample_depths[s.depth_p];
    const int c = (int)base[s.depth_p] + (s.folded ? n + 1 + s.obs_freq : s.obs_freq);
    cls[i] = c;
    null_tab[c] = s.null_logl;          // a function of the class (scan-chromosome.c:23-37)
  }
  const int C = so->n_chromosomes;
  std::vector<int32_t> cs(C), cn(C), cp(C), cb(C);
  for (int c = 0; c < C; c++) {
    cs[c] = so->chr_limits[c].start_index; cn[c] = so->chr_limits[c].n_snps;
    cp[c] = so->chr_limits[c].start_pos; cb[c] = so->chr_limits[c].bp_length;
  }
  fscl_problem_t p;
  memset(&p, 0, sizeof p);
  p.n_snps = N; p.chr = chr.data(); p.pos = pos.data(); p.cls = cls.data();
  p.n_chr = C; p.chr_start = cs.data(); p.chr_nsnps = cn.data(); p.chr_startpos = cp.data(); p.chr_bplen = cb.data();
  p.n_cls = n_cls; p.n_knots = K; p.step = (4.0 - (-20.0)) / (K + 1.);      // sm-spline.c:325
  p.coef = coef; p.null_tab = null_tab.data(); p.log_table = nullptr;
  p.eval_range = eval_range; p.bp_resl = bp_resl;
  // FSCL_B200_GPUS = n: one context per device FSCL_B200_DEVICE .. +n-1, joined by an NCCL communicator;
  // cells of the forward scan and permutations of the test shard over them (needs the counter-based stream)
  const int first_dev = env_int("FSCL_B200_DEVICE", 0);
  int n_gpus = env_int("FSCL_B200_GPUS", 1);
  if (n_gpus < 1) n_gpus = 1;
  if (n_gpus > fscl_device_count() - first_dev && fscl_device_count() > 0) n_gpus = fscl_device_count() - first_dev;
  if (n_gpus < 1) n_gpus = 1;
  g_bound.ctx.assign(n_gpus, nullptr);
  for (int g = 0; g < n_gpus; g++) {
    p.device = first_dev + g;
    if (fscl_create(&p, &g_bound.ctx[g]) != FSCL_OK) die("cannot set up the GPU scan");
  }
  if (n_gpus > 1 && fscl_comm_init_all(g_bound.ctx.data(), n_gpus) != FSCL_OK) die("cannot set up NCCL over the GPUs");
  g_bound.scan = so; g_bound.snps = so->snps; g_bound.sm_p = sm_p; g_bound.n_snps = N;
  g_bound.eval_range = eval_range; g_bound.bp_resl = bp_resl; g_bound.digest = digest_of(so);
}

// run f(rank) on every context: ranks >= 1 on their own host threads (collectives need all of them)
template <class F> void on_all_ranks(F f) {
  std::vector<std::thread> th;
  for (size_t g = 1; g < g_bound.ctx.size(); g++) th.emplace_back([=]() { f((int)g); });
  f(0);
  for (auto &t : th) t.join();
}

// ---- SIGINT: "current output at any time", a second one within the grace period aborts
// (scan-chromosome.c:557-569, :601-609) ----
volatile sig_atomic_t g_sigint = 0;
struct timeval g_watch;
double g_grace = 10.0;

double since_watch() {
  struct timeval now;
  gettimeofday(&now, nullptr);
  return (now.tv_sec - g_watch.tv_sec) + 1e-6 * (now.tv_usec - g_watch.tv_usec);
}

void on_sigint(int) {
  if (since_watch() < g_grace) {
    static const char msg[] = "\nanother interrupt signal received, aborting permutation\n";
    if (write(2, msg, sizeof msg - 1) < 0) {}
    _exit(255);                                  // the reference calls exit(-1)
  }
  g_sigint = 1;                                  // served between two rounds
  gettimeofday(&g_watch, nullptr);               // a second signal from now on aborts, as after the reference's dump
}

// <out>-nulldist, scan-chromosome.c:753-796 (sorts the stored CLRs of every point in place, as the reference does)
void output_null_distribution(const char *fname, scan_t *so) {
  std::string f2 = std::string(fname) + "-nulldist";
  FILE *f = fopen(f2.c_str(), "w");
  if (!f) { fprintf(stderr, "Can't open output file for CLR null distribution (%s)\n", strerror(errno)); return; }
  fprintf(f, "chr\tpos\tCLR\talpha\tp\tn");
  for (int j = 0; j < FSCL_NULL_SAVE; j++) fprintf(f, "\t%1.4f", j / (double)FSCL_NULL_SAVE);
  fprintf(f, "\n");
  for (int i = 0; i < so->n_scan_pts; i++) {
    scan_pt_t *p = so->scan_pts + i;
    const int m = std::min(p->permute_n, FSCL_NULL_SAVE);
    std::vector<float> v(p->permute_clr, p->permute_clr + (p->permute_clr ? m : 0));
    std::sort(v.begin(), v.end());
    fprintf(f, "%s\t%d\t%1.3f\t%1.3e\t%d\t%d", so->chr_limits[p->chr].name, p->sweep_pos, p->clr, exp(p->lalpha),
            p->permute_p, p->permute_n);
    for (float x : v) fprintf(f, "\t%1.2f", (double)x);
    fprintf(f, "\n");
  }
  fclose(f);
}

// ---- checkpoint of a permutation test (counter-based stream): header, counters, stored CLRs ----
// The header pins everything the counts depend on: a resume with other permutation parameters, other
// observed CLRs (another table build) or another input is refused instead of mixing two null distributions.
struct CkHeader {
  char magic[8];                 // "FSCLCKP2": layout version, little-endian host
  int32_t n_pts, t_next, n_permute, rng_kind;
  uint64_t seed;
  int32_t large_grid_sp, has_clr, n_snps, reserved;
  double permute_nbp, scan_width_mb;
  uint64_t obs_digest;           // of the observed (chr, sweep_pos, clr) of every point
  uint64_t input_digest;         // of the SNP data (digest_of)
};

uint64_t obs_digest(const std::vector<fscl_pt_t> &pts) {
  uint64_t h = 1469598103934665603ull;
  for (const fscl_pt_t &p : pts) {
    uint64_t bits;
    memcpy(&bits, &p.clr, 8);
    const uint64_t v[3] = {(uint64_t)(uint32_t)p.chr, (uint64_t)(uint32_t)p.sweep_pos, bits};
    for (uint64_t x : v) { h ^= x; h *= 1099511628211ull; }
  }
  return h;
}

CkHeader make_header(const std::vector<fscl_pt_t> &pts, int t_next, const fscl_permute_opts_t &o, bool has_clr) {
  CkHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, "FSCLCKP2", 8);
  h.n_pts = (int32_t)pts.size(); h.t_next = t_next; h.n_permute = o.n_permute; h.rng_kind = o.rng_kind;
  h.seed = o.seed; h.large_grid_sp = o.large_grid_sp; h.has_clr = has_clr; h.n_snps = g_bound.n_snps;
  h.permute_nbp = o.permute_nbp; h.scan_width_mb = o.scan_width_mb;
  h.obs_digest = obs_digest(pts); h.input_digest = g_bound.digest;
  return h;
}

bool write_checkpoint(const char *path, const std::vector<fscl_pt_t> &pts, const float *store, int t_next,
                      const fscl_permute_opts_t &o) {
  std::string tmp = std::string(path) + ".tmp";
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  const CkHeader h = make_header(pts, t_next, o, store != nullptr);
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  for (size_t i = 0; ok && i < pts.size(); i++) {
    const int32_t rec[5] = {pts[i].chr, pts[i].sweep_pos, pts[i].permute_p, pts[i].permute_n, pts[i].permute_finished};
    ok = fwrite(rec, sizeof rec, 1, f) == 1;
    const int m = std::min(pts[i].permute_n, FSCL_NULL_SAVE);
    if (ok && store && m > 0) ok = fwrite(store + i * (size_t)FSCL_NULL_SAVE, sizeof(float), m, f) == (size_t)m;
  }
  ok = ok && fflush(f) == 0 && fsync(fileno(f)) == 0;      // on disk before it can replace a good file
  ok = (fclose(f) == 0) && ok;
  return ok && rename(tmp.c_str(), path) == 0;
}

// returns t_next, or -1 when the file does not describe this test
int read_checkpoint(const char *path, std::vector<fscl_pt_t> &pts, float *store, const fscl_permute_opts_t &o) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  CkHeader h;
  const CkHeader want = make_header(pts, 0, o, true);
  bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, want.magic, 8) == 0 && h.n_pts == want.n_pts &&
            h.rng_kind == want.rng_kind && h.seed == want.seed && h.large_grid_sp == want.large_grid_sp && h.t_next >= 0 &&
            h.n_permute == want.n_permute && h.n_snps == want.n_snps && h.permute_nbp == want.permute_nbp &&
            h.scan_width_mb == want.scan_width_mb && h.obs_digest == want.obs_digest && h.input_digest == want.input_digest;
  for (size_t i = 0; ok && i < pts.size(); i++) {
    int32_t rec[5];
    ok = fread(rec, sizeof rec, 1, f) == 1 && rec[0] == pts[i].chr && rec[1] == pts[i].sweep_pos && rec[3] >= 0;
    if (!ok) break;
    pts[i].permute_p = rec[2]; pts[i].permute_n = rec[3]; pts[i].permute_finished = rec[4];
    const int m = std::min(rec[3], FSCL_NULL_SAVE);
    if (h.has_clr && m > 0) {
      std::vector<float> row(m);
      ok = fread(row.data(), sizeof(float), m, f) == (size_t)m;
      if (ok && store) memcpy(store + i * (size_t)FSCL_NULL_SAVE, row.data(), sizeof(float) * m);
    }
  }
  fclose(f);
  return ok ? h.t_next : -1;
}

// <out>-projected: the reference's permutation record plus the projected p-value it describes at README.md:11
// and never computes (definition: DESIGN.md 3.6).  The records of scan_output stay byte-compatible.
void output_projection(const char *fname, scan_t *so, const char *label) {
  if ((int)g_bound.proj_p.size() != so->n_scan_pts) return;
  std::string f2 = std::string(fname) + "-projected";
  FILE *f = fopen(f2.c_str(), "w");
  if (!f) { fprintf(stderr, "Can't open output file for projected p-values (%s)\n", strerror(errno)); return; }
  fprintf(f, "%schr\tpos\tCLR\talpha\tp\tn\tneglog10_p_empirical\tlambda\tneglog10_p_projected\n", label ? "label\t" : "");
  for (int i = 0; i < so->n_scan_pts; i++) {
    const scan_pt_t *p = so->scan_pts + i;
    const double pv = p->permute_p < 2 ? 1.0 / p->permute_n : (p->permute_p - 1.0) / (double)(p->permute_n - 1.0);
    const double pp = g_bound.proj_p[i];
    if (label) fprintf(f, "%s\t", label);
    fprintf(f, "%s\t%d\t%1.2f\t%1.3e\t%d\t%d\t%1.3f\t%1.4f\t%1.3f\n", so->chr_limits[p->chr].name, p->sweep_pos, p->clr,
            exp(p->lalpha), p->permute_p, p->permute_n, -log10(pv), g_bound.proj_lambda[i], pp > 0 ? -log10(pp) : 999.0);
  }
  fclose(f);
}

}  // namespace

extern "C" void init_log_table(void) {}      // the table is built inside fscl_create (sm-search.c:14-26)

// compute_sweep_model_tables (sm-spline.c:486).  With the reference's sm-spline.o linked (the product
// executable renames its entry point to ref_compute_sweep_model_tables) the default is the reference's own
// host build, bit-exact tables; FSCL_B200_DEVICE_TABLES=1 -- or no host build linked -- builds them on the
// GPU (fscl_build_tables: ascertainment terms, pjh / pbk / p_kescape mixtures and the spline fits as CUDA
// kernels; spline values within 5e-11 of the host build).
extern "C" sm_ptable_t *compute_sweep_model_tables(scan_t *so, double **fsp, int asc_depth, int asc_min_freq,
                                                   int ascbias_background_only, int include_invariant) {
  const bool want_device = env_int("FSCL_B200_DEVICE_TABLES", 0) != 0;
  if (!want_device && ref_compute_sweep_model_tables)
    return ref_compute_sweep_model_tables(so, fsp, asc_depth, asc_min_freq, ascbias_background_only, include_invariant);
  const int D = so->n_depths;
  const int K = (&spline_pts && spline_pts > 0) ? spline_pts : 200;
  std::vector<long> base(D + 1, 0);
  std::vector<int32_t> depths(D);
  for (int d = 0; d < D; d++) {
    depths[d] = so->sample_depths[d];
    base[d + 1] = base[d] + (depths[d] + 1) + (depths[d] / 2 + 1);
  }
  const int n_cls = (int)base[D];
  g_dev_tables.coef.assign((size_t)n_cls * K * 4, 0.0);
  std::vector<double> null_tab(n_cls);
  int32_t flags = 0;
  if (fscl_build_tables(env_int("FSCL_B200_DEVICE", 0), D, depths.data(), fsp, K, asc_depth, asc_min_freq,
                        ascbias_background_only, include_invariant, g_dev_tables.coef.data(), null_tab.data(), &flags) != FSCL_OK)
    die("device table build");
  if (logmsg) logmsg(3, (volatile char *)"Sweep model tables built on the GPU (%d depths, %d splines).\n", D, n_cls);
  // the pointer forest the reference's interface promises (fscl.h:64-76), over the flat table
  sm_ptable_t *sm_p = (sm_ptable_t *)calloc((size_t)D, sizeof(sm_ptable_t));
  if (!sm_p) abort();
  for (int d = 0; d < D; d++) {
    const int n = depths[d];
    sm_p[d].sample_size = n;
    sm_p[d].spline_func = (spline_t **)calloc((size_t)n + 1, sizeof(spline_t *));
    sm_p[d].fspline_func = (spline_t **)calloc((size_t)n / 2 + 1, sizeof(spline_t *));
    auto make = [&](long cls) {
      spline_t *sp = (spline_t *)calloc(1, sizeof(spline_t));
      sp->n = K;
      sp->knot_points = (double *)malloc(sizeof(double) * (size_t)(K + 1));
      for (int i = 0; i <= K; i++) sp->knot_points[i] = -20.0 + i * (24.0 / (K + 1.));
      sp->coef = (double **)malloc(sizeof(double *) * (size_t)K);
      for (int i = 0; i < K; i++) sp->coef[i] = g_dev_tables.coef.data() + ((size_t)cls * K + i) * 4;
      return sp;
    };
    for (int f = 0; f <= n; f++) sm_p[d].spline_func[f] = make(base[d] + f);
    for (int f = 0; f <= n / 2; f++) sm_p[d].fspline_func[f] = make(base[d] + n + 1 + f);
  }
  g_dev_tables.sm_p = sm_p; g_dev_tables.n_cls = n_cls; g_dev_tables.K = K;
  return sm_p;
}

extern "C" void compute_snp_null_model(scan_t *so, double **fsp) {
  for (int i = 0; i < so->n_snps; i++) {
    snp_t &s = so->snps[i];
    const int depth = so->sample_depths[s.depth_p];
    const double *f = fsp[s.depth_p];
    if (s.folded && s.obs_freq != depth - s.obs_freq) s.null_logl = log(f[s.obs_freq] + f[depth - s.obs_freq]);
    else s.null_logl = log(f[s.obs_freq]);
  }
}

extern "C" void scan_chromosome(scan_t *so, sm_ptable_t *sm_p, int eval_range, int bp_resl, int large_grid_sp,
                                int n_threads) {
  (void)n_threads;                           // the GPU replaces the pthread pool
  bind(so, sm_p, eval_range, bp_resl, true);
  int n = fscl_enum_cells(g_bound.ctx[0], large_grid_sp, nullptr, nullptr, nullptr, 0);
  if (n < 0) die("cell enumeration");
  std::vector<fscl_pt_t> pts(n > 0 ? n : 1);
  // with several GPUs the cells shard over them; every rank ends with all records
  std::vector<std::vector<fscl_pt_t>> scratch(g_bound.ctx.size());
  std::vector<int> rcs(g_bound.ctx.size(), FSCL_OK);
  int n_out = 0;
  on_all_ranks([&](int g) {
    int m = 0;
    fscl_pt_t *dst = pts.data();
    if (g > 0) { scratch[g].resize(n > 0 ? n : 1); dst = scratch[g].data(); }
    rcs[g] = fscl_scan_chromosome(g_bound.ctx[g], large_grid_sp, dst, n, &m);
    if (g == 0) n_out = m;
  });
  for (int rc : rcs) if (rc != FSCL_OK) die("scan_chromosome");
  so->scan_pts = (scan_pt_t *)calloc((size_t)(n_out > 0 ? n_out : 1), sizeof(scan_pt_t));
  if (!so->scan_pts) { fprintf(stderr, "\nout of memory\n"); abort(); }
  for (int i = 0; i < n_out; i++) {
    scan_pt_t &d = so->scan_pts[i];
    const fscl_pt_t &s = pts[i];
    d.chr = s.chr; d.nearest_snp = s.nearest_snp; d.sweep_pos = s.sweep_pos; d.n_snps = s.n_snps;
    d.window_start = s.window_start; d.window_end = s.window_end; d.lalpha = s.lalpha;
    d.null_logl = s.null_logl; d.sm_logl = s.sm_logl; d.clr = s.clr;
  }
  so->n_scan_pts = n_out;
  g_bound.large_grid_sp = large_grid_sp;
  if (logmsg) logmsg(3, (volatile char *)"\nInitial scan finished.\n");
}

extern "C" void scan_permute(scan_t *so, sm_ptable_t *sm_p, int n_permute, double permute_nbp, double alpha_factor,
                             int n_threads, int eval_range, int bp_resl, int large_grid_sp, double scan_width_mb) {
  (void)alpha_factor; (void)n_threads;       // unused by the reference too
  bind(so, sm_p, eval_range, bp_resl, false);
  const int n = so->n_scan_pts;
  const int n_gpus = (int)g_bound.ctx.size();
  std::vector<fscl_pt_t> pts(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) {
    const scan_pt_t &s = so->scan_pts[i];
    fscl_pt_t &d = pts[i];
    memset(&d, 0, sizeof d);
    d.chr = s.chr; d.nearest_snp = s.nearest_snp; d.sweep_pos = s.sweep_pos; d.n_snps = s.n_snps;
    d.window_start = s.window_start; d.window_end = s.window_end; d.lalpha = s.lalpha; d.clr = s.clr;
    d.permute_n = s.permute_n; d.permute_p = s.permute_p; d.permute_finished = s.permute_finished;
  }
  fscl_permute_opts_t o;
  memset(&o, 0, sizeof o);
  o.n_permute = n_permute; o.permute_nbp = permute_nbp; o.large_grid_sp = large_grid_sp;
  o.scan_width_mb = scan_width_mb; o.prune = 1; o.first_perm = 0; o.perm_stride = 1;
  // default: the process's libc rand() stream, exactly the one the reference consumes (fscl.c:135);
  // FSCL_B200_RNG=philox selects the counter-based stream (device-resident pipelined rounds, shardable;
  // the one to use for long tests -- a sequential stream forces rounds of 20 - max(permute_p) permutations)
  const char *rng = getenv("FSCL_B200_RNG");
  if ((rng && strcmp(rng, "philox") == 0) || (!rng && n_gpus > 1)) {
    o.rng_kind = FSCL_RNG_PHILOX;
    const char *seed = getenv("FSCL_B200_SEED");
    o.seed = seed ? strtoull(seed, nullptr, 0) : 0xFD821A6ull;
  } else o.rng_kind = FSCL_RNG_LIBC;
  if (n_gpus > 1 && o.rng_kind != FSCL_RNG_PHILOX) {
    fprintf(stderr, "\nfscl_b200: FSCL_B200_GPUS > 1 needs the counter-based stream (FSCL_B200_RNG=philox)\n");
    exit(1);
  }
  const char *ck_path = getenv("FSCL_B200_CHECKPOINT"), *resume = getenv("FSCL_B200_RESUME");
  const char *ck_every = getenv("FSCL_B200_CHECKPOINT_SECONDS"), *grace = getenv("FSCL_B200_SIGINT_GRACE");
  const char *stop_after = getenv("FSCL_B200_STOP_AFTER_ROUNDS");            // test hook: leave after a checkpoint
  const char *round_perms = getenv("FSCL_B200_ROUND_PERMS");                 // permutations per round (default: sized by the library)
  const char *round_delay = getenv("FSCL_B200_ROUND_DELAY_MS");              // test hook: pause between rounds
  if (round_perms && atoi(round_perms) > 0) o.round_perms = atoi(round_perms);
  const double ck_seconds = ck_every ? atof(ck_every) : 60.0;
  g_grace = grace ? atof(grace) : 10.0;
  if ((ck_path || resume) && o.rng_kind != FSCL_RNG_PHILOX) {
    fprintf(stderr, "\nfscl_b200: checkpoints need the counter-based stream (FSCL_B200_RNG=philox)\n");
    exit(1);
  }
  if ((ck_path || resume) && n_gpus > 1) {
    fprintf(stderr, "\nfscl_b200: checkpoints are single-GPU (unset FSCL_B200_GPUS)\n");
    exit(1);
  }
  // Stored permuted CLRs (scan_pt_t.permute_clr, scan-chromosome.c:496-497).  The reference allocates 40 KB per
  // cell up front (D11).  Counter-based stream: they stay on the device and come back packed, one gather + one
  // copy (fscl_perm_fetch_clr).  Sequential streams and checkpointed runs keep one row per cell in zero pages
  // that only become resident where a cell writes.
  const bool packed = o.rng_kind == FSCL_RNG_PHILOX && !ck_path && !resume && env_int("FSCL_B200_HOST_ROUNDS", 0) == 0;
  free(g_bound.clr_dense);
  g_bound.clr_dense = nullptr;
  g_bound.clr_packed.clear();
  g_bound.proj_p.clear(); g_bound.proj_lambda.clear();
  std::vector<int64_t> clr_off(n + 1, 0);
  float *store = nullptr;
  if (!packed) {
    store = g_bound.clr_dense = (float *)calloc((size_t)(n > 0 ? n : 1) * FSCL_NULL_SAVE, sizeof(float));
    if (!store) { fprintf(stderr, "\nout of memory\n"); abort(); }
  } else o.save_clr = 1;
  // The test runs round by round so that between two rounds the reference's SIGINT behaviour
  // (current output + null distribution; a second signal within 10 s aborts) can be served and,
  // with the counter-based stream, a checkpoint written (FSCL_B200_CHECKPOINT=file, every
  // FSCL_B200_CHECKPOINT_SECONDS, default 60) or a previous one continued (FSCL_B200_RESUME=file).
  auto fetch_packed = [&]() {
    if (!packed) return;
    if (fscl_perm_fetch_clr(g_bound.ctx[0], clr_off.data(), nullptr, 0) != FSCL_OK) die("stored CLRs");
    g_bound.clr_packed.assign((size_t)std::max<int64_t>(clr_off[n], 1), 0.0f);
    if (fscl_perm_fetch_clr(g_bound.ctx[0], clr_off.data(), g_bound.clr_packed.data(), (int64_t)g_bound.clr_packed.size()) != FSCL_OK)
      die("stored CLRs");
  };
  auto publish = [&]() {
    for (int i = 0; i < n; i++) {
      scan_pt_t &d = so->scan_pts[i];
      d.permute_n = pts[i].permute_n; d.permute_p = pts[i].permute_p; d.permute_finished = pts[i].permute_finished;
      d.permute_clr = packed ? g_bound.clr_packed.data() + clr_off[i] : store + (size_t)i * FSCL_NULL_SAVE;
    }
  };
  int t_start = 0;
  if (resume) {
    t_start = read_checkpoint(resume, pts, store, o);
    if (t_start < 0) { fprintf(stderr, "\nfscl_b200: checkpoint \"%s\" does not belong to this scan\n", resume); exit(1); }
  }
  gettimeofday(&g_watch, nullptr);
  struct sigaction sa, old_sa;
  memset(&sa, 0, sizeof sa);
  sa.sa_handler = on_sigint;
  sigemptyset(&sa.sa_mask);
  g_sigint = 0;
  sigaction(SIGINT, &sa, &old_sa);
  fscl_stats_t st;
  memset(&st, 0, sizeof st);
  // ranks >= 1 of a multi-GPU run: the same rounds on their own contexts (private copies of the points);
  // the all-gather inside every round keeps them in step with rank 0, which serves signals and output
  std::vector<std::vector<fscl_pt_t>> rank_pts(n_gpus);
  std::vector<std::thread> workers;
  for (int g = 1; g < n_gpus; g++) {
    rank_pts[g] = pts;
    workers.emplace_back([&, g]() {
      void *ctx = g_bound.ctx[g];
      if (fscl_perm_begin(ctx, rank_pts[g].data(), n, &o, nullptr) != FSCL_OK) die("scan_permute (worker)");
      for (;;) {
        int R = 0;
        if (fscl_perm_round_score(ctx, nullptr, 0, &R, nullptr) != FSCL_OK) die("scan_permute (worker)");
        if (R == 0) break;
        if (fscl_perm_round_commit(ctx, nullptr) != FSCL_OK) die("scan_permute (worker)");
      }
      if (fscl_perm_end(ctx, nullptr) != FSCL_OK) die("scan_permute (worker)");
    });
  }
  void *ctx0 = g_bound.ctx[0];
  if (fscl_perm_begin(ctx0, pts.data(), n, &o, store) != FSCL_OK) die("scan_permute");
  if (t_start > 0 && fscl_perm_seek(ctx0, t_start) != FSCL_OK) die("resume");
  struct timeval last_ck;
  gettimeofday(&last_ck, nullptr);
  for (int rounds = 0;; rounds++) {
    int R = 0;
    if (fscl_perm_round_score(ctx0, nullptr, 0, &R, nullptr) != FSCL_OK) die("scan_permute");
    if (R == 0) break;
    if (fscl_perm_round_commit(ctx0, nullptr) != FSCL_OK) die("scan_permute");
    if (round_delay) usleep((useconds_t)(1000.0 * atof(round_delay)));
    struct timeval now;
    gettimeofday(&now, nullptr);
    const bool ck_due = ck_path && ((now.tv_sec - last_ck.tv_sec) + 1e-6 * (now.tv_usec - last_ck.tv_usec) >= ck_seconds ||
                                    (stop_after && rounds + 1 >= atoi(stop_after)));
    if (g_sigint || ck_due) {
      int32_t t_next = 0;
      if (fscl_perm_snapshot(ctx0, &t_next) != FSCL_OK) die("snapshot");
      if (ck_path && !write_checkpoint(ck_path, pts, store, t_next, o))
        fprintf(stderr, "\nfscl_b200: cannot write checkpoint \"%s\" (%s)\n", ck_path, strerror(errno));
      last_ck = now;
      if (g_sigint) {
        fetch_packed();
        publish();
        if (&output_fname && output_fname) {
          scan_output(output_fname, so, 0, n_permute, &prepend_label ? prepend_label : nullptr);   // n_permute: this call's argument = fscl.c's global
          output_null_distribution(output_fname, so);
        }
        g_sigint = 0;
        gettimeofday(&g_watch, nullptr);
      }
      if (ck_due && stop_after && rounds + 1 >= atoi(stop_after)) {
        fscl_perm_end(ctx0, &st);
        fprintf(stderr, "\nfscl_b200: stopped after %d round(s), checkpoint at permutation %d\n", rounds + 1, t_next);
        exit(3);
      }
    }
  }
  if (fscl_perm_end(ctx0, &st) != FSCL_OK) die("scan_permute");
  for (auto &w : workers) w.join();
  sigaction(SIGINT, &old_sa, nullptr);
  fetch_packed();
  publish();
  // projected p-values, once at the end (README.md:11): on the device from the CLRs kept there, or on the host
  // from the stored rows of a sequential-stream run
  if (env_int("FSCL_B200_PROJECT", 1) != 0 && n > 0) {
    const double df = getenv("FSCL_B200_PROJECT_DF") ? atof(getenv("FSCL_B200_PROJECT_DF")) : 2.0;
    g_bound.proj_lambda.assign(n, 0.0); g_bound.proj_p.assign(n, 1.0);
    const int rc = packed ? fscl_perm_project(ctx0, df, g_bound.proj_lambda.data(), g_bound.proj_p.data())
                          : fscl_project_pvalues(n, pts.data(), store, df, g_bound.proj_lambda.data(), g_bound.proj_p.data());
    if (rc != FSCL_OK) die("p-value projection");
  }
  if (st.d7_hits)
    fprintf(stderr, "\nfscl_b200: %d permutation block(s) overshot the end of the SNP array "
            "(undefined behaviour in the reference, scan-chromosome.c:360-363; shifted to end at the array)\n", st.d7_hits);
  if (env_int("FSCL_B200_STATS", 0))
    fprintf(stderr, "\nfscl_b200: %d GPU(s), %lld rounds, %lld permutations scored, %lld searches, scan kernel %.1f ms\n", n_gpus,
            (long long)st.n_rounds, (long long)st.n_scored, (long long)st.n_tasks, st.score_kernel_ms);
  if (logmsg) logmsg(3, (volatile char *)"\rScanning snp block permutations... finished.\n");
}

// scan-chromosome.c:666-750, same records byte for byte
extern "C" void scan_output(char *output_fname, scan_t *so, int maximum_only, int n_permute, char *prepend_label) {
  FILE *f = stdout;
  if (output_fname) {
    f = fopen(output_fname, "w");
    if (!f) { fprintf(stderr, "Can't open output file \"%s\" (%s)", output_fname, strerror(errno)); return; }
  }
  const scan_pt_t *best = so->scan_pts;
  double max_clr = so->scan_pts[0].clr;
  for (int i = 1; i < so->n_scan_pts; i++)
    if (so->scan_pts[i].clr > max_clr) { max_clr = so->scan_pts[i].clr; best = so->scan_pts + i; }
  if (logmsg) {
    char where[96];
    const char *nm = so->chr_limits[best->chr].name;
    if (best->sweep_pos > 1000000) snprintf(where, sizeof where, "chromosome %s %1.2f Mb", nm, best->sweep_pos / 1e6);
    else if (best->sweep_pos > 2000) snprintf(where, sizeof where, "chromosome %s %1.2f Kb", nm, best->sweep_pos / 1e3);
    else snprintf(where, sizeof where, "chromosome %s %d bp", nm, best->sweep_pos);
    logmsg(3, (volatile char *)"\rOutput complete -- maximum CLR of %g at %s (alpha = %g)\n", max_clr, where,
           exp(best->lalpha));
  }
  if (maximum_only) {
    if (prepend_label) fprintf(f, "%s\t", prepend_label);
    fprintf(f, "%s\t%d\t%1.2f\t%1.3e\t%d\t%d\t%d\n", so->chr_limits[best->chr].name, best->sweep_pos, max_clr,
            exp(best->lalpha), best->n_snps, so->snps[best->window_start].pos, so->snps[best->window_end].pos);
    return;                                   // the reference leaves the file open here as well (:716)
  }
  for (int i = 0; i < so->n_scan_pts; i++) {
    const scan_pt_t *p = so->scan_pts + i;
    if (prepend_label) fprintf(f, "%s\t", prepend_label);
    if (n_permute > 0) {
      const double pv = p->permute_p < 2 ? 1.0 / p->permute_n : (p->permute_p - 1.0) / (double)(p->permute_n - 1.0);
      fprintf(f, "%s\t%d\t%1.2f\t%1.3e\t%d\t%d\t%1.3f\n", so->chr_limits[p->chr].name, p->sweep_pos, p->clr,
              exp(p->lalpha), p->permute_p, p->permute_n, -log10(pv));
    } else {
      fprintf(f, "%s\t%d\t%1.2f\t%1.3e\t%d\t%d\t%d\n", so->chr_limits[p->chr].name, p->sweep_pos, p->clr,
              exp(p->lalpha), p->n_snps, so->snps[p->window_start].pos, so->snps[p->window_end].pos);
    }
  }
  if (output_fname) fclose(f);
  // companion files of a permutation test: projected p-values (always, unless FSCL_B200_PROJECT=0) and, on
  // request, the null distribution the reference only writes from its SIGINT handler (FSCL_B200_NULLDIST=1)
  if (output_fname && n_permute > 0 && so == g_bound.scan) {
    output_projection(output_fname, so, prepend_label);
    if (env_int("FSCL_B200_NULLDIST", 0)) output_null_distribution(output_fname, so);
  }
}

// Single-point evaluation outside a bound scan is not part of main()'s path (only
// scan-chromosome.c called it); it is served for the problem scan_chromosome bound, at the point's
// sweep_pos: nearest site, window and null sum are re-derived from the position exactly as
// init_scan_result does (scan-chromosome.c:58-101), so a caller-made window is not honoured.
extern "C" void search_maxalpha(scan_pt_t *pt, snp_t *snps, sm_ptable_t *sm_p) {
  if (g_bound.ctx.empty() || !g_bound.scan || g_bound.scan->snps != snps || g_bound.sm_p != sm_p ||
      g_bound.digest != digest_of(g_bound.scan)) {
    fprintf(stderr, "\nfscl_b200: search_maxalpha needs the problem bound by scan_chromosome first\n");
    exit(1);
  }
  fscl_pt_t out;
  int32_t chr = pt->chr, pos = pt->sweep_pos;
  if (fscl_grid_eval(g_bound.ctx[0], 1, &chr, &pos, &out) != FSCL_OK) die("search_maxalpha");
  pt->lalpha = out.lalpha; pt->sm_logl = out.sm_logl; pt->clr = out.clr;
}
