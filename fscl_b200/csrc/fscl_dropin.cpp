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
#include <unistd.h>
#include <vector>

extern "C" void logmsg(int priority, volatile char *s, ...) __attribute__((weak));   // reference logmsg.o, if linked
// globals of fscl.c:33-34 that the reference's scan-chromosome.c imports for its SIGINT dump
// (scan-chromosome.c:550-552); weak, so the library also loads without the reference's main()
extern "C" char *output_fname __attribute__((weak));
extern "C" char *prepend_label __attribute__((weak));
extern "C" void scan_output(char *output_fname, scan_t *so, int maximum_only, int n_permute, char *prepend_label);

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
struct CkHeader {
  char magic[8];
  int32_t n_pts, t_next, n_permute, rng_kind;
  uint64_t seed;
  int32_t large_grid_sp, has_clr;
};

bool write_checkpoint(const char *path, const std::vector<fscl_pt_t> &pts, const float *store, int t_next,
                      const fscl_permute_opts_t &o) {
  std::string tmp = std::string(path) + ".tmp";
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  CkHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, "FSCLCKP1", 8);
  h.n_pts = (int32_t)pts.size(); h.t_next = t_next; h.n_permute = o.n_permute; h.rng_kind = o.rng_kind;
  h.seed = o.seed; h.large_grid_sp = o.large_grid_sp; h.has_clr = store != nullptr;
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  for (size_t i = 0; ok && i < pts.size(); i++) {
    const int32_t rec[5] = {pts[i].chr, pts[i].sweep_pos, pts[i].permute_p, pts[i].permute_n, pts[i].permute_finished};
    ok = fwrite(rec, sizeof rec, 1, f) == 1;
    const int m = std::min(pts[i].permute_n, FSCL_NULL_SAVE);
    if (ok && store && m > 0) ok = fwrite(store + i * (size_t)FSCL_NULL_SAVE, sizeof(float), m, f) == (size_t)m;
  }
  ok = (fclose(f) == 0) && ok;
  return ok && rename(tmp.c_str(), path) == 0;         // a half-written file never replaces a good one
}

// returns t_next, or -1 when the file does not describe this test
int read_checkpoint(const char *path, std::vector<fscl_pt_t> &pts, float *store, const fscl_permute_opts_t &o) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  CkHeader h;
  bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "FSCLCKP1", 8) == 0 && h.n_pts == (int32_t)pts.size() &&
            h.rng_kind == o.rng_kind && h.seed == o.seed && h.large_grid_sp == o.large_grid_sp && h.t_next >= 0;
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
  // The test runs round by round so that between two rounds the reference's SIGINT behaviour
  // (current output + null distribution; a second signal within 10 s aborts) can be served and,
  // with the counter-based stream, a checkpoint written (FSCL_B200_CHECKPOINT=file, every
  // FSCL_B200_CHECKPOINT_SECONDS, default 60) or a previous one continued (FSCL_B200_RESUME=file).
  auto publish = [&]() {
    for (int i = 0; i < n; i++) {
      scan_pt_t &d = so->scan_pts[i];
      d.permute_n = pts[i].permute_n; d.permute_p = pts[i].permute_p; d.permute_finished = pts[i].permute_finished;
      d.permute_clr = store + (size_t)i * FSCL_NULL_SAVE;
    }
  };
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
  if (fscl_perm_begin(g_bound.ctx, pts.data(), n, &o, store) != FSCL_OK) die("scan_permute");
  if (t_start > 0 && fscl_perm_seek(g_bound.ctx, t_start) != FSCL_OK) die("resume");
  struct timeval last_ck;
  gettimeofday(&last_ck, nullptr);
  for (int rounds = 0;; rounds++) {
    int R = 0;
    if (fscl_perm_round_score(g_bound.ctx, nullptr, 0, &R, nullptr) != FSCL_OK) die("scan_permute");
    if (R == 0) break;
    if (fscl_perm_round_commit(g_bound.ctx, nullptr) != FSCL_OK) die("scan_permute");
    if (round_delay) usleep((useconds_t)(1000.0 * atof(round_delay)));
    struct timeval now;
    gettimeofday(&now, nullptr);
    const bool ck_due = ck_path && ((now.tv_sec - last_ck.tv_sec) + 1e-6 * (now.tv_usec - last_ck.tv_usec) >= ck_seconds ||
                                    (stop_after && rounds + 1 >= atoi(stop_after)));
    if (g_sigint || ck_due) {
      int32_t t_next = 0;
      if (fscl_perm_snapshot(g_bound.ctx, &t_next) != FSCL_OK) die("snapshot");
      if (ck_path && !write_checkpoint(ck_path, pts, store, t_next, o))
        fprintf(stderr, "\nfscl_b200: cannot write checkpoint \"%s\" (%s)\n", ck_path, strerror(errno));
      last_ck = now;
      if (g_sigint) {
        publish();
        if (&output_fname && output_fname) {
          scan_output(output_fname, so, 0, n_permute, &prepend_label ? prepend_label : nullptr);   // n_permute: this call's argument = fscl.c's global
          output_null_distribution(output_fname, so);
        }
        g_sigint = 0;
        gettimeofday(&g_watch, nullptr);
      }
      if (ck_due && stop_after && rounds + 1 >= atoi(stop_after)) {
        fscl_perm_end(g_bound.ctx, &st);
        fprintf(stderr, "\nfscl_b200: stopped after %d round(s), checkpoint at permutation %d\n", rounds + 1, t_next);
        exit(3);
      }
    }
  }
  if (fscl_perm_end(g_bound.ctx, &st) != FSCL_OK) die("scan_permute");
  sigaction(SIGINT, &old_sa, nullptr);
  publish();
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
