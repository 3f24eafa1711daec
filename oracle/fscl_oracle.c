/* fscl_oracle.c -- TEST INFRASTRUCTURE ONLY (see fscl_oracle.h).
 *
 * Plain-C restatement of the FSCL CLR scan + block-permutation hot path.
 * Citations are to the read-only reference tree (slowkoni/fscl).  Nothing in
 * the product (fscl_b200/, include/) may include, link or call this file.
 * Pinned against the compiled, unmodified reference by tests/test_oracle_vs_ref.py.
 */
#include "fscl_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ logt */
/* sm-search.c:14-26  64 Ki-entry ln LUT, entry 0 := 0 (distance 0 == 1 bp) */
void fo_init_log_table(double *T) {
  T[0] = 0.0;
  for (int i = 1; i <= 0xFFFF; i++) T[i] = log((double)i);
}

/* sm-search.c:40-46 (D2: the 2^16 branch adds ln 2^17, kept as is) */
double fo_logt(const double *T, int d) {
  if (d < 0) d = -d;
  if (d > 0xFFFFFF) return 11.783502069519070 + T[d >> 16];
  if (d > 0xFFFF) return 5.545177444479562 + T[d >> 8];
  return T[d];
}

/* largest distance d >= 0 for which !(logt(d)+lalpha > LOG_AD_MAX); -1 if none.
 * logt is monotone non-decreasing in d, so the reference's "break at first
 * exceed" walks (sm-search.c:122-142) select exactly the sites with d <= dmax. */
int fo_dmax(const double *T, double lalpha) {
  if (fo_logt(T, 0) + lalpha > FO_LOG_AD_MAX) return -1;
  long lo = 0, hi = 0x7FFFFFFFL;               /* invariant: lo passes */
  if (!(fo_logt(T, (int)hi) + lalpha > FO_LOG_AD_MAX)) return (int)hi;
  while (hi - lo > 1) {
    long m = (lo + hi) / 2;
    if (fo_logt(T, (int)m) + lalpha > FO_LOG_AD_MAX) hi = m; else lo = m;
  }
  return (int)lo;
}

/* ---------------------------------------------------------- spline eval */
/* sm-spline.c:48-60: knot by division + truncation, clamp, global-basis cubic */
double fo_spline_eval(const double *row, int n_knots, double step, double x) {
  int i = (int)((x - FO_LOG_AD_MIN) / step);
  if (i >= n_knots) i = n_knots - 1;
  if (i < 0) i = 0;
  const double *c = row + 4 * (size_t)i;
  return x * (c[0] * x * x + c[1] * x + c[2]) + c[3];
}

/* ------------------------------------------------------ alpha candidates */
/* sm-search.c:276-296.  coarse[11] (FP-accumulated += 2.4); fine list for the
 * coarse index which_max; returns the number of fine candidates (14 or 15). */
int fo_alpha_lists(double *coarse, int which_max, double *fine) {
  double step = (FO_LOG_AD_MAX - FO_LOG_AD_MIN) / 10.0;
  int n = 0;
  for (double la = FO_LOG_AD_MIN; la <= FO_LOG_AD_MAX; la += step) coarse[n++] = la;
  if (which_max < 0 || !fine) return n;
  double m = coarse[which_max];
  double l = m - step; if (l < FO_LOG_AD_MIN) l = FO_LOG_AD_MIN;
  double r = m + step; if (r > FO_LOG_AD_MAX) r = FO_LOG_AD_MAX;
  double s = (r - l) / 15.;
  int k = 0;
  for (double la = l + s; la < r; la += s) fine[k++] = la;
  return k;
}

/* --------------------------------------------------------- sm_likelihood */
static inline double term(const fo_ctx_t *c, int cl, double log_ad) {
  /* snp_likelihood sm-search.c:85-103 */
  return fo_spline_eval(c->coef + (size_t)cl * c->n_knots * 4, c->n_knots, c->step, log_ad)
         - c->null_tab[cl];
}

/* sm-search.c:105-150: nearest first (early return), walk left, walk right */
void fo_sm_likelihood(const fo_ctx_t *c, const int *cls, fo_pt_t *pt) {
  const int *pos = c->pos;
  const double *T = c->logT;
  double la = pt->lalpha;
  double log_ad = fo_logt(T, abs(pt->sweep_pos - pos[pt->nearest])) + la;
  if (c->mode == 0) {
    double s = pt->null_logl;
    pt->sm_logl = s;
    if (log_ad > FO_LOG_AD_MAX) return;
    s += term(c, cls[pt->nearest], log_ad);
    for (int i = pt->nearest - 1; i >= pt->wstart; i--) {
      log_ad = fo_logt(T, pt->sweep_pos - pos[i]) + la;
      if (log_ad > FO_LOG_AD_MAX) break;
      s += term(c, cls[i], log_ad);
    }
    for (int i = pt->nearest + 1; i <= pt->wend; i++) {
      log_ad = fo_logt(T, abs(pos[i] - pt->sweep_pos)) + la;
      if (log_ad > FO_LOG_AD_MAX) break;
      s += term(c, cls[i], log_ad);
    }
    pt->sm_logl = s;
  } else {
    long double s = 0.0L;
    pt->sm_logl = 0.0;
    if (log_ad > FO_LOG_AD_MAX) return;
    s += term(c, cls[pt->nearest], log_ad);
    for (int i = pt->nearest - 1; i >= pt->wstart; i--) {
      log_ad = fo_logt(T, pt->sweep_pos - pos[i]) + la;
      if (log_ad > FO_LOG_AD_MAX) break;
      s += term(c, cls[i], log_ad);
    }
    for (int i = pt->nearest + 1; i <= pt->wend; i++) {
      log_ad = fo_logt(T, abs(pos[i] - pt->sweep_pos)) + la;
      if (log_ad > FO_LOG_AD_MAX) break;
      s += term(c, cls[i], log_ad);
    }
    pt->sm_logl = (double)s;
    if (c->mode == 2) {      /* absorb into the null sum the way the reference's accumulator does */
      volatile double t = pt->null_logl + (double)s;
      pt->sm_logl = t - pt->null_logl;
    }
  }
}

long fo_count_terms(const fo_ctx_t *c, const fo_pt_t *pt, double la) {
  const int *pos = c->pos;
  const double *T = c->logT;
  long n = 0;
  if (fo_logt(T, abs(pt->sweep_pos - pos[pt->nearest])) + la > FO_LOG_AD_MAX) return 0;
  n++;
  for (int i = pt->nearest - 1; i >= pt->wstart; i--) {
    if (fo_logt(T, pt->sweep_pos - pos[i]) + la > FO_LOG_AD_MAX) break;
    n++;
  }
  for (int i = pt->nearest + 1; i <= pt->wend; i++) {
    if (fo_logt(T, abs(pos[i] - pt->sweep_pos)) + la > FO_LOG_AD_MAX) break;
    n++;
  }
  return n;
}

/* sm-search.c:269-300 (live #else branch) */
static void maxalpha_impl(const fo_ctx_t *c, const int *cls, fo_pt_t *result, long *terms) {
  fo_pt_t tmp = *result, mx = *result;
  double base = (c->mode == 0) ? result->null_logl : 0.0;
  mx.sm_logl = -DBL_MAX;
  double step = (FO_LOG_AD_MAX - FO_LOG_AD_MIN) / 10.0;
  for (double la = FO_LOG_AD_MIN; la <= FO_LOG_AD_MAX; la += step) {
    tmp.lalpha = la;
    fo_sm_likelihood(c, cls, &tmp);
    if (terms) *terms += fo_count_terms(c, &tmp, la);
    if (tmp.sm_logl > mx.sm_logl) mx = tmp;
  }
  double l = mx.lalpha - step; if (l < FO_LOG_AD_MIN) l = FO_LOG_AD_MIN;
  double r = mx.lalpha + step; if (r > FO_LOG_AD_MAX) r = FO_LOG_AD_MAX;
  step = (r - l) / 15.;
  for (double la = l + step; la < r; la += step) {
    tmp.lalpha = la;
    fo_sm_likelihood(c, cls, &tmp);
    if (terms) *terms += fo_count_terms(c, &tmp, la);
    if (tmp.sm_logl > mx.sm_logl) mx = tmp;
  }
  mx.clr = 2.0 * (mx.sm_logl - base);
  *result = mx;
}

void fo_search_maxalpha(const fo_ctx_t *c, const int *cls, fo_pt_t *pt) {
  maxalpha_impl(c, cls, pt, NULL);
}

long fo_count_terms_maxalpha(const fo_ctx_t *c, const int *cls, const fo_pt_t *pt) {
  fo_pt_t t = *pt; long n = 0;
  maxalpha_impl(c, cls, &t, &n);
  return n;
}

/* ---------------------------------------------------------- scan points */
/* scan-chromosome.c:39-56 */
int fo_search_snppos(const int *pos, int n, double sweep_pos) {
  int i = 0, j = n;
  while (j - i > 1) {
    int m = (i + j) / 2;
    if (pos[m] < sweep_pos) i = m; else j = m;
  }
  if (j == n) return n - 1;
  if ((sweep_pos - pos[i]) < (pos[j] - sweep_pos)) return i;
  return j;
}

/* scan-chromosome.c:58-101 */
void fo_init_scan_result(const fo_ctx_t *c, const int *cls, fo_pt_t *pt, int chr, int pos) {
  int start = c->chr_start[chr], n = c->chr_nsnps[chr];
  int er = c->eval_range;
  pt->chr = chr;
  pt->nearest = start + fo_search_snppos(c->pos + start, n, (double)pos);
  /* D3: global index compared with the per-chromosome count, kept literally */
  int i = pt->nearest;
  while (i < n && c->pos[i] == pos) { i++; pos++; }
  pt->sweep_pos = pos;
  int chm_start = start, chm_stop = start + n - 1;
  if (pt->nearest - er < chm_start) {
    pt->wstart = chm_start;
    pt->wend = chm_start + er * 2;
    if (pt->wend > chm_stop) pt->wend = chm_stop;
  } else if (pt->nearest + er > chm_stop) {
    pt->wend = chm_stop;
    pt->wstart = chm_stop - er * 2;
    if (pt->wstart < chm_start) pt->wstart = chm_start;
  } else {
    pt->wstart = pt->nearest - er;
    pt->wend = pt->nearest + er;
  }
  pt->n_snps = pt->wend - pt->wstart + 1;
  double s = 0.;
  for (i = pt->wstart; i <= pt->wend; i++) s += c->null_tab[cls[i]];
  pt->null_logl = s;
  pt->sm_logl = -DBL_MAX;
  pt->lalpha = FO_LOG_AD_MAX;
  pt->clr = 0.;
  pt->permute_n = pt->permute_p = pt->finished = 0;
  pt->pad_ = 0;
}

/* scan-chromosome.c:103-139 (recursion unrolled into a loop) */
void fo_search_maxpos(const fo_ctx_t *c, const int *cls, int chr, int start_pos, int end_pos,
                      fo_pt_t *out, long *n_evals) {
  fo_pt_t a, b, m;
  fo_init_scan_result(c, cls, &a, chr, start_pos);
  fo_search_maxalpha(c, cls, &a);
  fo_init_scan_result(c, cls, &b, chr, end_pos);
  fo_search_maxalpha(c, cls, &b);
  long ne = 2;
  while (b.sweep_pos - a.sweep_pos > c->bp_resl) {
    fo_init_scan_result(c, cls, &m, chr, (a.sweep_pos + b.sweep_pos) / 2);
    fo_search_maxalpha(c, cls, &m);
    ne++;
    if ((a.clr + m.clr) >= (b.clr + m.clr)) b = m; else a = m;
  }
  *out = (a.clr > b.clr) ? a : b;
  if (n_evals) *n_evals += ne;
}

/* cell enumeration: scan-chromosome.c:178-205 (single-thread order) */
int fo_enum_cells(const fo_ctx_t *c, int G, int *cell_chr, int *cell_start, int *cell_end) {
  int chm = 0, n = 0;
  int pos = c->chr_startpos[0];
  for (;;) {
    if (chm == c->n_chr) break;
    if (pos >= c->chr_bplen[chm]) {
      chm++;
      if (chm == c->n_chr) break;
      pos = c->chr_startpos[chm];
    }
    int s = pos, e = pos + G;
    pos += G;
    if (e > c->chr_bplen[chm]) e = c->chr_bplen[chm];
    if (cell_chr) { cell_chr[n] = chm; cell_start[n] = s; cell_end[n] = e; }
    n++;
  }
  return n;
}
int fo_count_cells(const fo_ctx_t *c, int G) { return fo_enum_cells(c, G, NULL, NULL, NULL); }

static int pt_cmp(const fo_pt_t *a, const fo_pt_t *b) {
  if (a->chr != b->chr) return a->chr < b->chr ? -1 : 1;
  if (a->sweep_pos != b->sweep_pos) return a->sweep_pos < b->sweep_pos ? -1 : 1;
  return 0;
}

/* scan_chromosome, scan-chromosome.c:228-266, single-threaded; stable sort by
 * (chr, sweep_pos) (equal keys only arise for identical records) */
int fo_scan_cells(const fo_ctx_t *c, const int *cls, int G, fo_pt_t *out, long *n_evals) {
  int n = fo_count_cells(c, G);
  int *cc = malloc(sizeof(int) * 3 * (size_t)(n > 0 ? n : 1));
  fo_enum_cells(c, G, cc, cc + n, cc + 2 * n);
  for (int i = 0; i < n; i++)
    fo_search_maxpos(c, cls, cc[i], cc[n + i], cc[2 * n + i], out + i, n_evals);
  free(cc);
  for (int i = 1; i < n; i++) {            /* insertion sort: input is almost sorted */
    fo_pt_t t = out[i];
    int j = i - 1;
    while (j >= 0 && pt_cmp(out + j, &t) > 0) { out[j + 1] = out[j]; j--; }
    out[j + 1] = t;
  }
  return n;
}

/* ------------------------------------------------------------------ RNG */
/* glibc 2.39 random_r TYPE_3 (degree 31, separation 3): srandom_r seeds r[0..30]
 * with the Lehmer generator 16807 mod 2^31-1, discards 310 outputs; each output
 * is (r[i-31] + r[i-3]) >> 1.  Call sites fscl.c:135, scan-chromosome.c:350-351,
 * 440, 492. */
void fo_rng_seed_glibc(fo_rng_t *g, unsigned seed) {
  memset(g, 0, sizeof *g);
  g->kind = 0;
  int32_t r[344];
  if (seed == 0) seed = 1;
  r[0] = (int32_t)seed;
  for (int i = 1; i < 31; i++) {
    long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    r[i] = (int32_t)w;
  }
  for (int i = 31; i < 34; i++) r[i] = r[i - 31];
  for (int i = 34; i < 344; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
  /* keep the last 34 words as a ring: slot (i % 34) */
  for (int i = 310; i < 344; i++) g->r[i % 34] = r[i];
  g->k = 344;
}

int32_t fo_rng_next(fo_rng_t *g) {
  g->n_draws++;
  if (g->kind == 0) {
    int k = g->k;
    uint32_t v = (uint32_t)g->r[(k - 31) % 34] + (uint32_t)g->r[(k - 3) % 34];
    g->r[k % 34] = (int32_t)v;
    g->k = k + 1;
    if (g->k >= 34 * 1000000) g->k -= 34 * 999999;   /* keep k bounded, same residues */
    return (int32_t)(v >> 1);
  }
  return fo_philox_draw(g->key0, g->key1, g->perm, 0, g->ctr++);
}

/* Philox4x32-10 (Salmon et al. 2011, Random123): counter-based mode for
 * sharded permutations.  Not in the reference. */
void fo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* one 31-bit draw addressed by (permutation t, stream, index):
 * stream 0 = block draws (idx = draw number inside permutation t),
 * stream 1 = pruning uniform (idx = scan point index). */
int32_t fo_philox_draw(uint32_t key0, uint32_t key1, uint32_t t, uint32_t stream, uint32_t idx) {
  uint32_t ctr[4] = { idx, stream, t, 0u }, key[2] = { key0, key1 }, out[4];
  fo_philox4x32_10(ctr, key, out);
  return (int32_t)(out[0] >> 1);
}

void fo_rng_seed_philox(fo_rng_t *g, uint64_t seed) {
  memset(g, 0, sizeof *g);
  g->kind = 1;
  g->key0 = (uint32_t)seed;
  g->key1 = (uint32_t)(seed >> 32);
}
void fo_rng_set_perm(fo_rng_t *g, uint32_t t) { g->perm = t; g->ctr = 0; }

/* ---------------------------------------------------- block permutation */
/* scan-chromosome.c:336-389.  Only the class (obs_freq, depth_p, folded,
 * null_logl) moves; chr/pos stay.  D7 (k > n_snps, reference indexes the heap
 * below the array) is DEFINED here as "shift the block so it ends at n_snps"
 * and reported in *flags bit 0. */
int fo_block_permute(const fo_ctx_t *c, fo_rng_t *g, double permute_nbp, double scan_width_mb,
                     const int *cls_in, int *cls_out, int *blocks, int max_blocks, int *flags) {
  int n = c->n_snps, nb = 0;
  const int *chr = c->chr, *pos = c->pos;
  memcpy(cls_out, cls_in, sizeof(int) * (size_t)n);
  int i = 0;
  while (i < n) {
    int j = (int)(fo_rng_next(g) / (2147483647 + 1.0) * n);
    int k = j + (int)(-1.0 / permute_nbp * log(fo_rng_next(g) / (2147483647 + 1.0)));
    while (k < n && chr[k] == chr[j] && pos[k] - pos[j] < scan_width_mb * 1e6) k++;
    if (i + (k - j) >= n) k = n;
    if (k > n) {
      if (flags) *flags |= 1;
      j = n - (k - j);
      k = n;
    }
    if (blocks && nb < max_blocks) { blocks[3 * nb] = i; blocks[3 * nb + 1] = j; blocks[3 * nb + 2] = k; }
    nb++;
    while (j < k && i < n && j < n) {
      int t = cls_out[i]; cls_out[i] = cls_out[j]; cls_out[j] = t;
      i++; j++;
    }
  }
  return nb;
}

/* scan_permute / scan_permute_thread, scan-chromosome.c:412-546 at n_threads=1.
 * kind 0: one draw is burnt first (usleep :440), block and pruning draws share
 * the stream.  kind 1: permutation t draws blocks from stream 0 of t and the
 * pruning uniform of scan point s from (t, stream 1, s).  Returns the number
 * of permutations scored (n_permute+1 unless every point finished, D5). */
int fo_scan_permute(const fo_ctx_t *c, const int *cls, fo_pt_t *pts, int n_pts, int n_permute,
                    double permute_nbp, int G, double scan_width_mb, fo_rng_t *g, int prune,
                    float *permute_clr, int *flags, long *n_evals) {
  int n = c->n_snps;
  int *pcls = malloc(sizeof(int) * (size_t)n);
  int *active = malloc(sizeof(int) * (size_t)(n_pts > 0 ? n_pts : 1));
  int n_active = n_pts, t = -1, scored = 0;
  for (int i = 0; i < n_pts; i++) active[i] = i;
  if (g->kind == 0) (void)fo_rng_next(g);
  for (;;) {
    if (g->kind == 1) fo_rng_set_perm(g, (uint32_t)(t + 1));
    fo_block_permute(c, g, permute_nbp, scan_width_mb, cls, pcls, NULL, 0, flags);
    t++;
    int k = 0;
    for (int i = 0; i < n_active; i++)
      if (!pts[active[i]].finished) active[k++] = active[i];
    n_active = k;
    if (n_active == 0) break;
    if (t > n_permute) break;
    scored++;
    for (int a = 0; a < n_active; a++) {
      int s = active[a];
      fo_pt_t mx;
      int start = pts[s].sweep_pos - (pts[s].sweep_pos % G);          /* D4 */
      fo_search_maxpos(c, pcls, pts[s].chr, start, start + G, &mx, n_evals);
      if (mx.clr >= pts[s].clr) {
        pts[s].permute_p++;
        if (prune && pts[s].permute_p >= 20) {                         /* D6 */
          int32_t r = (g->kind == 0) ? fo_rng_next(g)
                                     : fo_philox_draw(g->key0, g->key1, (uint32_t)t, 1, (uint32_t)s);
          if (pts[s].permute_p / (double)pts[s].permute_n >= r / (2147483647 + 1.0))
            pts[s].finished = 1;
        }
      }
      if (permute_clr && pts[s].permute_n < FO_NULL_SAVE)
        permute_clr[(size_t)s * FO_NULL_SAVE + pts[s].permute_n] = (float)mx.clr;
      pts[s].permute_n++;
    }
  }
  free(pcls);
  free(active);
  return scored;
}

/* scan_output records, scan-chromosome.c:718-746 */
int fo_format_record(char *buf, int cap, const char *label, const char *chr_name, const fo_pt_t *pt,
                     int n_permute, int pos_wstart, int pos_wend) {
  int o = 0;
  if (label) o += snprintf(buf + o, cap - o, "%s\t", label);
  if (n_permute > 0) {
    double pv;
    if (pt->permute_p < 2) pv = 1.0 / pt->permute_n;
    else pv = (pt->permute_p - 1.0) / (double)(pt->permute_n - 1.0);
    o += snprintf(buf + o, cap - o, "%s\t%d\t%1.2f\t%1.3e\t%d\t%d\t%1.3f\n", chr_name, pt->sweep_pos,
                  pt->clr, exp(pt->lalpha), pt->permute_p, pt->permute_n, -log10(pv));
  } else {
    o += snprintf(buf + o, cap - o, "%s\t%d\t%1.2f\t%1.3e\t%d\t%d\t%d\n", chr_name, pt->sweep_pos,
                  pt->clr, exp(pt->lalpha), pt->n_snps, pos_wstart, pos_wend);
  }
  return o;
}

/* --------------------------------------------------------------- tables */
/* log_fact / lchoose, sm-spline.c:18-46 (table built by the same running sum) */
static double *lf_tab = NULL;
static int lf_max = 0;
static double fo_log_fact(int n) {
  if (n < 0) return -DBL_MAX;
  if (n == 0 || n == 1) return 0.;
  if (n > lf_max) {
    lf_tab = realloc(lf_tab, sizeof(double) * (size_t)(n + 1));
    if (lf_max == 0) { lf_tab[0] = 0; lf_max = 1; }
    for (int i = lf_max; i <= n; i++) lf_tab[i] = lf_tab[i - 1] + log((double)i);
    lf_max = n;
  }
  return lf_tab[n];
}
double fo_lchoose(int n, int k) {
  if (n == 0 && k == 0) return 0.;
  if (k > n || n == 0) return -DBL_MAX;
  return fo_log_fact(n) - fo_log_fact(k) - fo_log_fact(n - k);
}

/* compute_snp_null_model, scan-chromosome.c:23-37, tabulated per class */
void fo_null_model(const double *fsp, int depth, double *null_unf, double *null_fold) {
  for (int f = 0; f <= depth; f++) null_unf[f] = log(fsp[f]);
  for (int f = 0; f <= depth / 2; f++) {
    if (f != depth - f) null_fold[f] = log(fsp[f] + fsp[depth - f]);
    else null_fold[f] = log(fsp[f]);
  }
}

/* background_fsp, background-fsp.c:182-270 (+ neutral_spectra :19-51).
 * D1 (unfolded sites land on depth-obs_freq) is inherited literally. */
static void downsample(double *d, const double *fsp, int n, int N, int include_invariant) {
  /* hypergeometric_downsample_fsp background-fsp.c:72-88 */
  if (include_invariant) {
    for (int m = 0; m <= N; m++)
      for (int k = 0; k <= m && k <= n; k++)
        d[k] += exp(fo_lchoose(m, k) + fo_lchoose(N - m, n - k) - fo_lchoose(N, n)) * fsp[m];
  } else {
    for (int m = 1; m <= N; m++)
      for (int k = 1; k <= m && k < n; k++)
        d[k] += exp(fo_lchoose(m, k) + fo_lchoose(N - m, n - k) - fo_lchoose(N, n)) * fsp[m];
  }
}

void fo_background_fsp(int n_snps, const int *obs_freq, const int *depth_p, const int *folded,
                       int n_depths, const int *depths, int force_neutral, int include_invariant,
                       double **fsp) {
  if (force_neutral) {
    int n_inv = 0, n_fix = 0;
    for (int i = 0; i < n_snps; i++) {
      if (obs_freq[i] == 0) n_inv++;
      if (obs_freq[i] == depths[depth_p[i]]) n_fix++;
    }
    for (int i = 0; i < n_depths; i++) {
      int m = depths[i];
      fsp[i][0] = n_inv; fsp[i][m] = n_fix;
      double ss = 0.;
      for (int k = 1; k < m; k++) ss += 1 / (double)k;
      int n_seg = n_snps - n_fix - n_inv;
      for (int k = 1; k < m; k++) fsp[i][k] = (1. / (double)k) / ss * n_seg;
      for (int k = 0; k <= m; k++) fsp[i][k] /= (double)n_snps;
    }
    return;
  }
  int max_depth = -1000;
  for (int m = 0; m < n_depths; m++) {
    for (int k = 0; k <= depths[m]; k++) fsp[m][k] = 0.;
    if (depths[m] > max_depth) max_depth = depths[m];
  }
  fo_log_fact(max_depth + 1);
  double *tmp = calloc((size_t)max_depth + 1, sizeof(double));
  for (int i = 0; i < n_snps; i++) {
    int depth = depths[depth_p[i]];
    double wa, wd;
    if (folded[i]) {
      if (obs_freq[i] == 0) { wa = 1; wd = 0; }
      else if (obs_freq[i] == depth) { wa = 0; wd = 1; }
      else { wa = 1. / obs_freq[i]; wd = 1. / (depth - obs_freq[i]); }
    } else { wd = 1.; wa = 0.; }
    if (depth == max_depth) {
      tmp[obs_freq[i]] += wa / (wa + wd);
      tmp[depth - obs_freq[i]] += wd / (wa + wd);
    }
  }
  double s = 0.;
  for (int k = 0; k <= max_depth; k++) s += tmp[k];
  for (int k = 0; k <= max_depth; k++) tmp[k] /= s;
  for (int m = 0; m < n_depths; m++) {
    int depth = depths[m];
    downsample(fsp[m], tmp, depth, max_depth, include_invariant);
    double ss = 0.;
    for (int k = 0; k <= depth; k++) ss += fsp[m][k];
    for (int k = 0; k <= depth; k++) fsp[m][k] /= ss;
  }
  free(tmp);
}

/* asc-bias.c:12-25 */
static double ascprob(int k, int d, int min_obs, int n) {
  double no_asc = 0.;
  for (int i = 0; i < min_obs; i++)
    no_asc += exp(fo_lchoose(k, d - i) + fo_lchoose(n - k, i)) + exp(fo_lchoose(n - k, d - i) + fo_lchoose(k, i));
  no_asc /= exp(fo_lchoose(n, d));
  return 1.0 - no_asc;
}

/* asc-bias.c:27-52 */
void fo_ascbias_adjust_background(const double *bsf, int n, int asc_depth, int min_obs, double *adj) {
  double *asc = calloc((size_t)n + 1, sizeof(double));
  double asc_sum = 0., adj_sum = 0.;
  for (int i = 1; i < n; i++) { asc[i] = ascprob(i, asc_depth, min_obs, n); asc_sum += asc[i]; }
  for (int i = 1; i < n; i++) asc[i] /= asc_sum;
  adj[0] = adj[n] = 0.;
  for (int i = 1; i < n; i++) { adj[i] = bsf[i] / asc[i]; adj_sum += adj[i]; }
  for (int i = 1; i < n; i++) adj[i] /= adj_sum;
  free(asc);
}

/* asc-bias.c:97-109 */
static void ascbias_adjust_expect(double *p, int n, int min_obs, int d) {
  double s = 0;
  for (int i = 0; i <= n; i++) s += p[i] * ascprob(i, d, min_obs, n);
  for (int i = 0; i <= n; i++) p[i] = p[i] * ascprob(i, d, min_obs, n) / s;
}

/* solve_linear_system sm-spline.c:63-118: unpivoted elimination restricted to a
 * band of 8, rows normalised, back substitution.  (The rescue branch for a
 * vanishing pivot :67-88 is restated too.) */
static long fo_rescues = 0, fo_zeroed = 0;
long fo_rescue_count(void) { return fo_rescues; }
long fo_zeroed_count(void) { return fo_zeroed; }
static void solve_band(double *b, double **m, double *v, int n) {
  for (int i = 0; i < n; i++) {
    if (fabs(m[i][i]) < 1e-20) {
      fo_rescues++;
      int mx = i;
      for (int j = i + 1; j < n; j++)
        if (fabs(m[j][i]) > 0 && (mx == i || fabs(fabs(m[j][i]) - 1) < fabs(fabs(m[mx][i]) - 1))) mx = j;
      if (mx == i) { fprintf(stderr, "oracle: ill conditioned spline system\n"); exit(1); }
      for (int k = 0; k < n; k++) m[i][k] += m[mx][k];
      v[i] += v[mx];
    }
    double f = m[i][i];
    for (int k = i; k < i + 8 && k < n; k++) m[i][k] /= f;
    v[i] /= f;
    for (int j = i + 1; j < i + 8 && j < n; j++) {
      f = m[j][i];
      for (int k = i; k < i + 8 && k < n; k++) m[j][k] = m[j][k] - m[i][k] * f;
      v[j] = v[j] - v[i] * f;
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    if (fabs(m[i][i]) < 1e-10) { b[i] = 0; fo_zeroed++; continue; }
    b[i] = v[i];
    for (int k = i + 1; k < i + 8 && k < n; k++) b[i] -= m[i][k] * b[k];
  }
}

/* estimate_spline sm-spline.c:120-220: natural cubic spline through n+1 points
 * posed as a 4n x 4n system in the global monomial basis. */
void fo_estimate_spline(const double *x, const double *y, int n, double *coef) {
  int N = 4 * (n + 1);
  double **m = malloc(sizeof(double *) * (size_t)N);
  double *v = calloc((size_t)N, sizeof(double)), *b = calloc((size_t)N, sizeof(double));
  for (int i = 0; i < N; i++) m[i] = calloc((size_t)N, sizeof(double));
  m[0][0] = 6 * x[0];
  m[0][1] = 2;
  int i, j, k;
  for (i = 1, j = 0, k = 0; k < n - 1; i += 4, j += 4, k++) {
    m[i][j] = x[k] * x[k] * x[k]; m[i][j + 1] = x[k] * x[k]; m[i][j + 2] = x[k]; m[i][j + 3] = 1.;
    v[i] = y[k];
    m[i + 1][j] = x[k + 1] * x[k + 1] * x[k + 1]; m[i + 1][j + 1] = x[k + 1] * x[k + 1];
    m[i + 1][j + 2] = x[k + 1]; m[i + 1][j + 3] = 1;
    v[i + 1] = y[k + 1];
    m[i + 2][j] = 3 * x[k + 1] * x[k + 1]; m[i + 2][j + 1] = 2 * x[k + 1]; m[i + 2][j + 2] = 1; m[i + 2][j + 3] = 0;
    m[i + 2][j + 4] = -3 * x[k + 1] * x[k + 1]; m[i + 2][j + 5] = -2 * x[k + 1]; m[i + 2][j + 6] = -1; m[i + 2][j + 7] = 0;
    v[i + 2] = 0;
    m[i + 3][j] = 6 * x[k + 1]; m[i + 3][j + 1] = 2; m[i + 3][j + 2] = 0; m[i + 3][j + 3] = 0;
    m[i + 3][j + 4] = -6 * x[k + 1]; m[i + 3][j + 5] = -2; m[i + 3][j + 6] = 0; m[i + 3][j + 7] = 0;
    v[i + 3] = 0;
  }
  m[i][j] = x[k] * x[k] * x[k]; m[i][j + 1] = x[k] * x[k]; m[i][j + 2] = x[k]; m[i][j + 3] = 1;
  v[i] = y[k];
  m[i + 1][j] = x[n] * x[n] * x[n]; m[i + 1][j + 1] = x[n] * x[n]; m[i + 1][j + 2] = x[n]; m[i + 1][j + 3] = 1;
  v[i + 1] = y[n];
  m[i + 2][j] = 6 * x[n]; m[i + 2][j + 1] = 2; m[i + 2][j + 2] = 0; m[i + 2][j + 3] = 0;
  v[i + 2] = 0;
  solve_band(b, m, v, 4 * n);
  for (i = 0; i < 4 * n; i++) coef[i] = b[i];
  for (i = 0; i < N; i++) free(m[i]);
  free(m); free(v); free(b);
}

/* p_kescape sm-spline.c:236-240 */
static double p_kescape(int k, int n, double ad) {
  if (k == 0) return exp(-n * ad);
  return exp(fo_lchoose(n, k) + k * log(1.0 - exp(-ad)) - (n - k) * ad);
}

/* compute_sweep_model_fsp sm-spline.c:316-484 (exact-pjh branch; the spline-pjh
 * branch is dead, :327).  fsp here is what compute_sweep_model_tables passes:
 * the background, ascertainment-adjusted when asc_depth > 0 (:500-505). */
void fo_sweep_model_tables(const double *fsp_in, int n, int spline_pts, int asc_depth, int asc_min_freq,
                           int ascbias_background_only, int include_invariant,
                           double *coef_unf, double *coef_fold) {
  double step = (FO_LOG_AD_MAX - FO_LOG_AD_MIN) / (spline_pts + 1.);
  double *fsp = malloc(sizeof(double) * (size_t)(n + 1));
  if (asc_depth > 0) fo_ascbias_adjust_background(fsp_in, n, asc_depth, asc_min_freq, fsp);
  else memcpy(fsp, fsp_in, sizeof(double) * (size_t)(n + 1));
  size_t W = (size_t)n + 1;
  double *pjh = malloc(sizeof(double) * W * W), *pbk = malloc(sizeof(double) * W * W);
  for (int j = 0; j <= n; j++)
    for (int h = 0; h <= n; h++) {
      double s = 0.;
      for (int i = j; i <= n; i++)
        s += fsp[i] * exp(fo_lchoose(i, j) + fo_lchoose(n - i, h - j) - fo_lchoose(n, h));
      pjh[j * W + h] = s;
    }
  for (int b = 0; b <= n; b++)
    for (int k = 0; k < n; k++) {
      double s = 0.;
      int q = b - (n - k) + 1;
      if (q > 0) s += pjh[q * W + k + 1] * (q / (double)(k + 1));
      if (b < k + 1) s += pjh[b * W + k + 1] * ((k + 1 - b) / (double)(k + 1));
      pbk[b * W + k] = s;
    }
  int P = spline_pts + 1, nf = n / 2 + 1;
  double *x = malloc(sizeof(double) * (size_t)P);
  double *y = malloc(sizeof(double) * W * (size_t)P), *fy = malloc(sizeof(double) * (size_t)nf * (size_t)P);
  double *p = malloc(sizeof(double) * W);
  for (int i = 0; i <= spline_pts; i++) {
    double log_ad = FO_LOG_AD_MIN + i * step;
    double ad = exp(log_ad);
    double p_sum = 0.;
    for (int f = 0; f <= n; f++) {
      p[f] = p_kescape(n, n, ad) * fsp[f];
      for (int k = 0; k < n; k++) p[f] += p_kescape(k, n, ad) * pbk[f * W + k];
      p_sum += p[f];
    }
    if (!include_invariant) {
      p_sum -= p[0] + p[n];
      p[0] = p[n] = 0.;
    }
    for (int f = 0; f <= n; f++) p[f] /= p_sum;
    if (asc_depth > 0 && ascbias_background_only == 0) ascbias_adjust_expect(p, n, asc_min_freq, asc_depth);
    for (int f = 0; f <= n; f++) y[(size_t)f * P + i] = (p[f] == 0.) ? log(DBL_MIN) : log(p[f]);
    int f;
    for (f = 0; f < n - f; f++)
      fy[(size_t)f * P + i] = (p[f] + p[n - f] == 0.) ? log(DBL_MIN) : log(p[f] + p[n - f]);
    if (f == n - f) fy[(size_t)f * P + i] = (p[f] == 0.) ? log(DBL_MIN) : log(p[f]);
    x[i] = log_ad;
  }
  for (int f = 0; f <= n; f++) fo_estimate_spline(x, y + (size_t)f * P, spline_pts, coef_unf + (size_t)f * spline_pts * 4);
  for (int f = 0; f <= n - f; f++) fo_estimate_spline(x, fy + (size_t)f * P, spline_pts, coef_fold + (size_t)f * spline_pts * 4);
  free(fsp); free(pjh); free(pbk); free(x); free(y); free(fy); free(p);
}
