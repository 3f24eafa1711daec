/* fscl_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, single-threaded CPU restatement of the FSCL hot path
 * (slowkoni/fscl): CLR grid-eval, alpha maximisation, per-cell bisection,
 * contiguous-block permutation, exceedance counting with stochastic pruning,
 * the sweep-model / spline / ascertainment tables and the output records.
 * Every function cites the reference file:line it follows.
 *
 * Pinning: the reference has no tests or golden vectors, so this
 * oracle is pinned against the UNMODIFIED reference compiled here as
 * oracle/_ref/libfscl_ref.so + oracle/_ref/fscl_ref (tests/test_oracle_vs_ref.py)
 * and against fixtures under tests/golden/ generated from that binary.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.
 */
#ifndef FSCL_ORACLE_H
#define FSCL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_LOG_AD_MIN (-20.0)
#define FO_LOG_AD_MAX (4.0)
#define FO_NULL_SAVE 10000

/* Flattened data model.  A SNP "class" is the row of the spline table a site
 * selects: (depth_p, folded, obs_freq)  [sm-search.c:92-100].  null_tab[cls]
 * is the site's null log-likelihood [scan-chromosome.c:23-37]; it is a pure
 * function of the class, which is why only cls[] has to be permuted
 * (scan-chromosome.c:365-379 swaps obs_freq, depth_p, folded, null_logl). */
typedef struct {
  int n_snps;
  const int *chr;      /* [n_snps] chromosome index                */
  const int *pos;      /* [n_snps] sorted within chromosome        */
  int n_chr;
  const int *chr_start;   /* chr_limits_t.start_index */
  const int *chr_nsnps;   /* chr_limits_t.n_snps      */
  const int *chr_startpos;/* chr_limits_t.start_pos   */
  const int *chr_bplen;   /* chr_limits_t.bp_length   */
  int n_cls;
  int n_knots;            /* spline_t.n  (= --splines, 200) */
  double step;            /* log_ad_step = 24/(splines+1)   */
  const double *coef;     /* [n_cls][n_knots][4]            */
  const double *null_tab; /* [n_cls]                        */
  const double *logT;     /* [65536] integer-log LUT        */
  int eval_range;         /* 81920 */
  int bp_resl;            /* 128   */
  int mode;               /* 0: reference arithmetic (sm_logl = null + sum, sequential)
                             1: delta arithmetic (long-double sum of the per-site terms,
                                sm_logl := that sum, clr = 2*sum)
                             2: as 1, then quantised to the reference accumulator's grid:
                                sm_logl := (null_logl + sum) - null_logl, so that sums below
                                ulp(null_logl)/2 tie exactly as they do in the reference.
                                This is the arithmetic the CUDA path implements. */
} fo_ctx_t;

typedef struct {
  int chr, nearest, sweep_pos, n_snps, wstart, wend;
  double lalpha, null_logl, sm_logl, clr;
  int permute_n, permute_p, finished;
  int pad_;
} fo_pt_t;

/* RNG: kind 0 = glibc TYPE_3 rand() restatement; kind 1 = Philox4x32-10 counter mode */
typedef struct {
  int kind;
  int32_t r[34]; int k;             /* glibc state: ring of the last 34 words */
  uint32_t key0, key1;              /* philox key (seed)               */
  uint32_t perm;                    /* philox: permutation index t     */
  uint32_t ctr;                     /* philox: running draw counter    */
  uint64_t n_draws;
} fo_rng_t;

void   fo_init_log_table(double *T);
double fo_logt(const double *T, int d);
double fo_spline_eval(const double *coef_row, int n_knots, double step, double x);
int    fo_alpha_lists(double *coarse, int which_max, double *fine);
int    fo_dmax(const double *T, double lalpha);
void   fo_sm_likelihood(const fo_ctx_t *c, const int *cls, fo_pt_t *pt);
void   fo_search_maxalpha(const fo_ctx_t *c, const int *cls, fo_pt_t *pt);
int    fo_search_snppos(const int *pos, int n, double sweep_pos);
void   fo_init_scan_result(const fo_ctx_t *c, const int *cls, fo_pt_t *pt, int chr, int pos);
void   fo_search_maxpos(const fo_ctx_t *c, const int *cls, int chr, int start_pos, int end_pos,
                        fo_pt_t *out, long *n_evals);
int    fo_count_cells(const fo_ctx_t *c, int G);
int    fo_enum_cells(const fo_ctx_t *c, int G, int *cell_chr, int *cell_start, int *cell_end);
int    fo_scan_cells(const fo_ctx_t *c, const int *cls, int G, fo_pt_t *out, long *n_evals);
long   fo_count_terms(const fo_ctx_t *c, const fo_pt_t *pt, double lalpha);
long   fo_count_terms_maxalpha(const fo_ctx_t *c, const int *cls, const fo_pt_t *pt);

void     fo_rng_seed_glibc(fo_rng_t *g, unsigned seed);
void     fo_rng_seed_philox(fo_rng_t *g, uint64_t seed);
void     fo_rng_set_perm(fo_rng_t *g, uint32_t t);
int32_t  fo_rng_next(fo_rng_t *g);                 /* 31-bit, like rand() */
int32_t  fo_philox_draw(uint32_t key0, uint32_t key1, uint32_t t, uint32_t stream, uint32_t idx);
void     fo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* blocks[3*b+0..2] = (i, j, k) for every block applied (after clamping);
 * returns #blocks; flags bit0 = D7 hit (defined here as "shift block to end") */
int  fo_block_permute(const fo_ctx_t *c, fo_rng_t *g, double permute_nbp, double scan_width_mb,
                      const int *cls_in, int *cls_out, int *blocks, int max_blocks, int *flags);
int  fo_scan_permute(const fo_ctx_t *c, const int *cls, fo_pt_t *pts, int n_pts, int n_permute,
                     double permute_nbp, int G, double scan_width_mb, fo_rng_t *g, int prune,
                     float *permute_clr /* [n_pts][FO_NULL_SAVE] or NULL */, int *flags,
                     long *n_evals);
int  fo_format_record(char *buf, int cap, const char *label, const char *chr_name, const fo_pt_t *pt,
                      int n_permute, int pos_wstart, int pos_wend);

/* tables */
double fo_lchoose(int n, int k);
void fo_null_model(const double *fsp_d, int depth, double *null_unf /* [depth+1] */,
                   double *null_fold /* [depth/2+1] */);
void fo_background_fsp(int n_snps, const int *obs_freq, const int *depth_p, const int *folded,
                       int n_depths, const int *depths, int force_neutral, int include_invariant,
                       double **fsp /* [n_depths] -> [depth+1] */);
void fo_ascbias_adjust_background(const double *bsf, int n, int asc_depth, int min_obs, double *adj);
void fo_sweep_model_tables(const double *fsp, int n, int spline_pts, int asc_depth, int asc_min_freq,
                           int ascbias_background_only, int include_invariant,
                           double *coef_unf /* [n+1][spline_pts][4] */,
                           double *coef_fold /* [n/2+1][spline_pts][4] */);
void fo_estimate_spline(const double *x, const double *y, int n, double *coef /* [n][4] */);
long fo_rescue_count(void);   /* times the vanishing-pivot branch (sm-spline.c:67-88) ran */
long fo_zeroed_count(void);   /* coefficients zeroed in back substitution (sm-spline.c:107-111) */

#ifdef __cplusplus
}
#endif
#endif
