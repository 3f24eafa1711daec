// fscl_dropin.cpp -- reference-signature entry points over the C-ABI of libfscl_b200.so.
// See include/fscl_dropin.h and INTEGRATION.md.  Error convention of the reference
// (logmsg.c:34, kmacros.h:13-20): message on stderr, then exit(1).
#include "fscl_dropin.h"
#include "fscl_b200.h"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" void logmsg(int priority, volatile char *s, ...) __attribute__((weak));   // reference logmsg.o, if linked

namespace {

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "\nfscl_b200: %s: %s\n", what, fscl_last_error());
  exit(1);
}

struct Bound {                 // one resident problem, keyed by the caller's scan_t
  const scan_t *scan = nullptr;
  void *ctx = nullptr;
  std::vector<int32_t> cls;
  int large_grid_sp = 0;
} g_bound;

void release() {
  if (g_bound.ctx) fscl_destroy(g_bound.ctx);
  g_bound = Bound();
}

// sm_ptable_t pointer forest + snp_t array -> the flattened problem of fscl_b200.h
void bind(scan_t *so, sm_ptable_t *sm_p, int eval_range, int bp_resl) {
  if (g_bound.ctx && g_bound.scan == so) return;
  release();
  const int D = so->n_depths, N = so->n_snps;
  std::vector<long> base(D + 1, 0);
  for (int d = 0; d < D; d++) base[d + 1] = base[d] + (so->sample_depths[d] + 1) + (so->sample_depths[d] / 2 + 1);
  const int n_cls = (int)base[D];
  const int K = sm_p[0].spline_func[0]->n;
  std::vector<double> coef((size_t)n_cls * K * 4), null_tab(n_cls, 0.0);
  for (int d = 0; d < D; d++) {
    const int n = so->sample_depths[d];
    for (int f = 0; f <= n; f++)
      memcpy(&coef[(size_t)(base[d] + f) * K * 4], sm_p[d].spline_func[f]->coef[0], sizeof(double) * K * 4);
    for (int f = 0; f <= n / 2; f++)
      memcpy(&coef[(size_t)(base[d] + n + 1 + f) * K * 4], sm_p[d].fspline_func[f]->coef[0], sizeof(double) * K * 4);
  }
  std::vector<int32_t> chr(N), pos(N);
  g_bound.cls.resize(N);
  for (int i = 0; i < N; i++) {
    const snp_t &s = so->snps[i];
    chr[i] = s.chr; pos[i] = s.pos;
    const int n = so->sample_depths[s.depth_p];
    const int c = (int)base[s.depth_p] + (s.folded ? n + 1 + s.obs_freq : s.obs_freq);
    g_bound.cls[i] = c;
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
  p.n_snps = N; p.chr = chr.data(); p.pos = pos.data(); p.cls = g_bound.cls.data();
  p.n_chr = C; p.chr_start = cs.data(); p.chr_nsnps = cn.data(); p.chr_startpos = cp.data(); p.chr_bplen = cb.data();
  p.n_cls = n_cls; p.n_knots = K; p.step = (4.0 - (-20.0)) / (K + 1.);      // sm-spline.c:325
  p.coef = coef.data(); p.null_tab = null_tab.data(); p.log_table = nullptr;
  p.eval_range = eval_range; p.bp_resl = bp_resl;
  const char *dev = getenv("FSCL_B200_DEVICE");
  p.device = dev ? atoi(dev) : 0;
  if (fscl_create(&p, &g_bound.ctx) != FSCL_OK) die("cannot set up the GPU scan");
  g_bound.scan = so;
}

}  // namespace

extern "C" void init_log_table(void) {}      // the table is built inside fscl_create (sm-search.c:14-26)

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
  bind(so, sm_p, eval_range, bp_resl);
  int n = fscl_enum_cells(g_bound.ctx, large_grid_sp, nullptr, nullptr, nullptr, 0);
  if (n < 0) die("cell enumeration");
  std::vector<fscl_pt_t> pts(n > 0 ? n : 1);
  int n_out = 0;
  if (fscl_scan_chromosome(g_bound.ctx, large_grid_sp, pts.data(), n, &n_out) != FSCL_OK) die("scan_chromosome");
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
  bind(so, sm_p, eval_range, bp_resl);
  const int n = so->n_scan_pts;
  std::vector<fscl_pt_t> pts(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) {
    const scan_pt_t &s = so->scan_pts[i];
    fscl_pt_t &d = pts[i];
    memset(&d, 0, sizeof d);
    d.chr = s.chr; d.nearest_snp = s.nearest_snp; d.sweep_pos = s.sweep_pos; d.n_snps = s.n_snps;
    d.window_start = s.window_start; d.window_end = s.window_end; d.lalpha = s.lalpha; d.clr = s.clr;
    d.permute_n = s.permute_n; d.permute_p = s.permute_p; d.permute_finished = s.permute_finished;
  }
  float *store = (float *)malloc(sizeof(float) * FSCL_NULL_SAVE * (size_t)(n > 0 ? n : 1));
  if (!store) { fprintf(stderr, "\nout of memory\n"); abort(); }
  fscl_permute_opts_t o;
  memset(&o, 0, sizeof o);
  o.n_permute = n_permute; o.permute_nbp = permute_nbp; o.large_grid_sp = large_grid_sp;
  o.scan_width_mb = scan_width_mb; o.prune = 1; o.first_perm = 0; o.perm_stride = 1;
  // default: the process's libc rand() stream, exactly the one the reference consumes (fscl.c:135);
  // FSCL_B200_RNG=philox selects the counter-based stream (shardable, batched rounds)
  const char *rng = getenv("FSCL_B200_RNG");
  if (rng && strcmp(rng, "philox") == 0) {
    o.rng_kind = FSCL_RNG_PHILOX;
    const char *seed = getenv("FSCL_B200_SEED");
    o.seed = seed ? strtoull(seed, nullptr, 0) : 0xFD821A6ull;
  } else o.rng_kind = FSCL_RNG_LIBC;
  fscl_stats_t st;
  if (fscl_scan_permute(g_bound.ctx, pts.data(), n, &o, store, &st) != FSCL_OK) die("scan_permute");
  for (int i = 0; i < n; i++) {
    scan_pt_t &d = so->scan_pts[i];
    d.permute_n = pts[i].permute_n; d.permute_p = pts[i].permute_p; d.permute_finished = pts[i].permute_finished;
    d.permute_clr = store + (size_t)i * FSCL_NULL_SAVE;
  }
  if (st.d7_hits)
    fprintf(stderr, "\nfscl_b200: %d permutation block(s) overshot the end of the SNP array "
            "(undefined behaviour in the reference, scan-chromosome.c:360-363; shifted to end at the array)\n", st.d7_hits);
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
}

// Single-point evaluation outside a bound scan is not part of main()'s path (only
// scan-chromosome.c called it); it needs the scan context, so it is served for the
// bound problem only.
extern "C" void search_maxalpha(scan_pt_t *pt, snp_t *snps, sm_ptable_t *sm_p) {
  (void)sm_p;
  if (!g_bound.ctx || !g_bound.scan || g_bound.scan->snps != snps) {
    fprintf(stderr, "\nfscl_b200: search_maxalpha needs the problem bound by scan_chromosome first\n");
    exit(1);
  }
  fscl_pt_t out;
  int32_t chr = pt->chr, pos = pt->sweep_pos;
  if (fscl_grid_eval(g_bound.ctx, 1, &chr, &pos, &out) != FSCL_OK) die("search_maxalpha");
  pt->lalpha = out.lalpha; pt->sm_logl = out.sm_logl; pt->clr = out.clr;
}
