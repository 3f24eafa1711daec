/* fscl_b200.h -- C-ABI of the B200-native CLR sweep scan + block-permutation test.
 *
 * This library replaces the hot path of slowkoni/fscl (reference tree cited as
 * file:line below).  The reference has no plugin/FFI layer; its boundary is the
 * internal C API of fscl.h:103-116 called from main() (fscl.c:316-335).  Two
 * layers are exported:
 *
 *  1. fscl_*  (this header): plain pointers + sizes, flattened tables.  Host
 *     buffers in, host buffers out; all device memory is owned by fscl_ctx.
 *  2. the reference's own entry points with the reference's struct layouts
 *     (include/fscl_dropin.h, libfscl_dropin.so), implemented on top of 1.
 *
 * There is no CPU fallback: every entry point fails (non-zero return,
 * fscl_last_error()) when no sm_100 device is usable.
 */
#ifndef FSCL_B200_H
#define FSCL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FSCL_OK 0
#define FSCL_ERR_CUDA 1
#define FSCL_ERR_ARG 2
#define FSCL_ERR_NODEVICE 3

#define FSCL_NULL_SAVE 10000          /* CLR_NULL_DIST_SAVE scan-chromosome.c:227 */
#define FSCL_PRUNE_MIN_P 20           /* scan-chromosome.c:490 */

/* Flattened problem.  A site's "class" selects one spline of the sweep-model
 * table: (depth_p, folded, obs_freq) -> sm_p[depth_p].(f)spline_func[obs_freq]
 * (sm-search.c:92-100).  null_tab[cls] is snp_t.null_logl
 * (scan-chromosome.c:23-37), a function of the class only. */
typedef struct {
  int32_t n_snps;
  const int32_t *chr;          /* [n_snps] snp_t.chr  (fscl.h:8)  sorted (chr,pos) */
  const int32_t *pos;          /* [n_snps] snp_t.pos  (fscl.h:9)                   */
  const int32_t *cls;          /* [n_snps] class index, 0 <= cls < n_cls           */
  int32_t n_chr;
  const int32_t *chr_start;    /* chr_limits_t.start_index (fscl.h:29) */
  const int32_t *chr_nsnps;    /* chr_limits_t.n_snps      (fscl.h:30) */
  const int32_t *chr_startpos; /* chr_limits_t.start_pos   (fscl.h:31) */
  const int32_t *chr_bplen;    /* chr_limits_t.bp_length   (fscl.h:32) */
  int32_t n_cls;
  int32_t n_knots;             /* spline_t.n (fscl.h:65) = --splines */
  double step;                 /* log_ad_step sm-spline.c:325 */
  const double *coef;          /* [n_cls][n_knots][4]  spline_t.coef (fscl.h:67) */
  const double *null_tab;      /* [n_cls] */
  const double *log_table;     /* [65536] init_log_table sm-search.c:14-26, or NULL: built with libm */
  int32_t eval_range;          /* 81920  fscl.c:175 */
  int32_t bp_resl;             /* 128    fscl.c:174 */
  int32_t device;              /* CUDA device ordinal */
  int32_t max_round_perms;     /* 0 = default; permutation slots resident at once */
} fscl_problem_t;

/* One scan point; field meaning = scan_pt_t fscl.h:35-51 (without the
 * permute_clr pointer and scan_running). sm_logl/null_logl are filled by the
 * forward scan only. clr = 2 * sum of per-site log-likelihood differences
 * (sm-search.c:298 evaluated without the null_logl cancellation). */
typedef struct {
  int32_t chr, nearest_snp, sweep_pos, n_snps, window_start, window_end;
  double lalpha, null_logl, sm_logl, clr;
  int32_t permute_n, permute_p, permute_finished, reserved;
} fscl_pt_t;

enum { FSCL_RNG_GLIBC = 0,   /* own restatement of glibc TYPE_3 rand(), seeded by .seed */
       FSCL_RNG_PHILOX = 1,  /* Philox4x32-10 keyed by .seed, counter (idx, stream, t, 0) */
       FSCL_RNG_LIBC = 2 };  /* the process's libc rand() (drop-in: stream seeded at fscl.c:135) */

typedef struct {
  int32_t n_permute;        /* scores n_permute+1 permutations (scan-chromosome.c:444,466) */
  double permute_nbp;       /* --permute-nbp */
  int32_t large_grid_sp;    /* -G */
  double scan_width_mb;     /* -w */
  int32_t rng_kind;
  uint64_t seed;
  int32_t prune;            /* 1 = stochastic pruning as scan-chromosome.c:488-494; 0 = score everything */
  int32_t first_perm;       /* sharding: score permutations t with t % perm_stride == first_perm */
  int32_t perm_stride;      /* 1 on a single GPU */
  int32_t round_perms;      /* 0 = auto */
  int32_t save_clr;         /* 1 = keep the permuted CLRs on the device (fscl_perm_fetch_clr, fscl_perm_project)
                               even when no dense permute_clr buffer is given */
} fscl_permute_opts_t;

typedef struct {
  int64_t n_scored;         /* permutations scored                         */
  int64_t n_tasks;          /* (permutation, cell) searches launched       */
  int64_t n_grid_evals;     /* fscl_grid_eval points + cell searches of permutation tests (a search is 2 + bisection-steps
                               search_maxalpha calls, 12 at -G 100000; neighbouring cells share an end point, see below) */
  int64_t n_rounds;
  int64_t n_blocks;         /* permutation blocks applied                  */
  int32_t d7_hits;          /* blocks that overshot the array end (defined behaviour) */
  int32_t bound_failures;   /* coarse passes repeated exactly because a bounded candidate could not be excluded
                               (filled by fscl_get_stats; counted since fscl_create / fscl_reset_stats) */
  double score_kernel_ms;   /* CUDA-event time of the scan kernel, summed  */
  double permute_kernel_ms; /* CUDA-event time of the block-permute kernel */
  int64_t score_launches, permute_launches;
  int64_t kernel_launches;  /* every kernel of this library launched for the context */
  /* Arithmetic contract: sums are formed in a fixed tree and returned as (N + sum) - N with N the window null
   * log-likelihood (the reference accumulates onto N, sm-search.c:113), so decisions -- the alpha choice, the bisection
   * branch, the exceedance test -- equal the reference's except where two candidates differ by a few ulp(N) without being
   * equal.  These count such decisions (filled by fscl_get_stats; since fscl_create / fscl_reset_stats):            */
  int64_t near_ties_alpha;     /* grid-evals whose best and second-best alpha candidate differ by 0 < gap < 8 ulp(N)  */
  int64_t near_ties_position;  /* bisection steps whose two end points differ by 0 < gap < 32 ulp(N)                  */
  int64_t fine_bound_failures; /* fine passes repeated exactly because a bounded fine candidate could not be excluded  */
} fscl_stats_t;

const char *fscl_last_error(void);
int fscl_device_count(void);

int  fscl_create(const fscl_problem_t *p, void **ctx_out);
void fscl_destroy(void *ctx);

/* cell enumeration of scan_thread (scan-chromosome.c:178-205); returns the count */
int fscl_enum_cells(void *ctx, int large_grid_sp, int32_t *cell_chr, int32_t *cell_start,
                    int32_t *cell_end, int cap);
/* init_scan_result + search_maxalpha at n given positions (scan-chromosome.c:58, sm-search.c:269) */
int fscl_grid_eval(void *ctx, int n, const int32_t *chr, const int32_t *pos, fscl_pt_t *out);
/* search_maxpos over n cells (scan-chromosome.c:126) on the observed classes */
int fscl_search_maxpos(void *ctx, int n, const int32_t *chr, const int32_t *start_pos,
                       const int32_t *end_pos, fscl_pt_t *out);
/* scan_chromosome (scan-chromosome.c:228): all cells, sorted by (chr, sweep_pos) */
int fscl_scan_chromosome(void *ctx, int large_grid_sp, fscl_pt_t *out, int cap, int *n_out);
/* snp_block_permute (scan-chromosome.c:336) for permutation index t of the
 * stream described by opts; writes the permuted classes and, if blocks != NULL,
 * up to max_blocks (i, j, len) triples. Returns #blocks or <0. Test hook. */
int fscl_block_permute(void *ctx, const fscl_permute_opts_t *opts, int t, int32_t *cls_out,
                       int32_t *blocks, int max_blocks);
/* The same draw phase executed by the device kernel (counter-based stream only); test hook.
 * *flags: 1 = D7 hit, 2 = a length too close to an integer to trust the device log (the
 * library then redraws on the host), 4 = block list overflow. */
int fscl_device_gen_blocks(void *ctx, const fscl_permute_opts_t *opts, int t, int32_t *blocks,
                           int max_blocks, int32_t *flags);
/* scan_permute (scan-chromosome.c:582): pts in/out (permute_p/n/finished updated);
 * permute_clr = [n_pts][FSCL_NULL_SAVE] floats or NULL */
int fscl_scan_permute(void *ctx, fscl_pt_t *pts, int n_pts, const fscl_permute_opts_t *opts,
                      float *permute_clr, fscl_stats_t *stats);
/* The same test as rounds, for sharding permutations over GPUs:
 *   begin; loop { round_score -> [caller: allreduce(sum) of the matrix over ranks]
 *                 -> round_commit } until *R_out == 0; end.
 * Rank g of n sets opts->first_perm = g, opts->perm_stride = n (Philox only).
 * d_matrix is a DEVICE buffer of cap doubles laid out [R][n_active] (rows of
 * other ranks are zero-filled) or NULL to use an internal one (single GPU).
 * pts / permute_clr must stay valid until fscl_perm_end. */
int fscl_perm_begin(void *ctx, fscl_pt_t *pts, int n_pts, const fscl_permute_opts_t *opts,
                    float *permute_clr);
int fscl_perm_round_score(void *ctx, double *d_matrix, int64_t cap, int *R_out, int *n_active_out);
int fscl_perm_round_commit(void *ctx, const double *d_matrix);
int fscl_perm_end(void *ctx, fscl_stats_t *stats);
/* Between rounds: bring pts / permute_clr up to date without ending the test (what the reference's
 * SIGINT handler dumps, scan-chromosome.c:557-569; also the content of a checkpoint).
 * *t_next_out = index of the next permutation to be scored. */
int fscl_perm_snapshot(void *ctx, int32_t *t_next_out);
/* Right after fscl_perm_begin: continue at permutation t_next with the counters, finished flags
 * and stored CLRs found in pts / permute_clr (a checkpoint).  Counter-based stream only: the
 * result equals an uninterrupted run bit for bit. */
int fscl_perm_seek(void *ctx, int32_t t_next);
/* Stored permuted CLRs (scan_pt_t.permute_clr, scan-chromosome.c:496-497) without the 40 KB-per-cell dense
 * table of the reference: after fscl_perm_end / fscl_perm_snapshot of a test begun with
 * permute_clr != NULL or opts.save_clr, cell i holds offsets[i+1] - offsets[i] = min(permute_n, FSCL_NULL_SAVE)
 * values at packed + offsets[i].  packed == NULL fills offsets[n_pts + 1] only.  One gather kernel + one copy. */
int fscl_perm_fetch_clr(void *ctx, int64_t *offsets, float *packed, int64_t cap);
/* Projected p-values (README.md:11 of the reference; never implemented there -- definition in DESIGN.md 3.6)
 * computed on the device from the CLRs kept there: lambda = max(mean - df, 0), p = P[ncx2(df, lambda) >= clr]. */
int fscl_perm_project(void *ctx, double df, double *lambda_out, double *p_out);
/* Multi-GPU (the intent of fscl-master.c:1-18).  A context with a communicator shards
 * fscl_scan_chromosome over cells and fscl_scan_permute / the round API over permutations
 * (rank g scores t = g mod world; one ncclAllGather of the round's CLR rows per round on the library's
 * stream; every rank replays the same counting rule, so results are identical on all ranks and equal to
 * the single-GPU ones).  Every rank must make the same calls with the same arguments.
 *   one process per GPU : rank 0 calls fscl_comm_unique_id, ships the 128 bytes to the others (any
 *                         channel), every rank calls fscl_comm_init_rank;
 *   one process, n GPUs : fscl_comm_init_all on n contexts (one per device), then one host thread per
 *                         context.
 * libnccl.so.2 is loaded on first use (FSCL_B200_NCCL overrides the path). */
int fscl_comm_unique_id(char *id128);
int fscl_comm_init_rank(void *ctx, int world, int rank, const char *id128);
int fscl_comm_init_all(void **ctxs, int n);
int fscl_comm_destroy(void *ctx);
int fscl_comm_move(void *dst_ctx, void *src_ctx);   /* same device: a new problem on the same rank keeps the communicator */
int fscl_comm_info(void *ctx, int *world, int *rank);
/* Host-side pieces of the path, callable without a device (same code the GPU
 * path runs on the host): alpha candidate tables [12][15] (row 11 = coarse list,
 * sm-search.c:276-296, with the per-alpha distance cut), cell enumeration, the
 * draw phase of snp_block_permute, one round of counting/pruning, RNG draws. */
int fscl_host_alpha_tables(const double *log_table, double *la, int32_t *dmax, int32_t *n);
/* The range-maximum table behind the candidate bounds, as fscl_create builds it: out =
 * [levels][n_knots+2][n_cls], entry [l][k][c] >= spline_c(x) - null_c for every x whose knot index is in k .. k+2^l-1. */
int fscl_host_piece_maxima(int n_cls, int n_knots, double step, const double *coef, const double *null_tab,
                           double *out, int64_t cap, int32_t *levels_out);
int fscl_host_enum_cells(int n_chr, const int32_t *chr_startpos, const int32_t *chr_bplen, int large_grid_sp,
                         int32_t *cell_chr, int32_t *cell_start, int32_t *cell_end, int cap);
int fscl_host_gen_blocks(int n_snps, const int32_t *chr, const int32_t *pos, const fscl_permute_opts_t *opts,
                         int t, int32_t *blocks, int max_blocks, int32_t *d7_hits);
int fscl_host_replay_round(fscl_pt_t *pts, const int32_t *active, int n_active, const double *matrix,
                           int R, int t_first, const fscl_permute_opts_t *opts, float *permute_clr);
/* the same replay for rows laid out as the library's per-round all-gather delivers them: rank g's rows
 * [q][n_active] at gathered + g * chunk, rank g owning the permutations t = g (mod world) */
int fscl_host_replay_gathered(fscl_pt_t *pts, const int32_t *active, int n_active, const double *gathered, int world,
                              int64_t chunk, int R, int t_first, const fscl_permute_opts_t *opts, float *permute_clr);
int fscl_host_rng_draws(int rng_kind, uint64_t seed, uint32_t t, uint32_t stream, int n, int32_t *out);
/* Sweep-model tables built on the device: compute_sweep_model_tables (sm-spline.c:486) with
 * ascbias_adjust_background / ascbias_adjust_expect (asc-bias.c:27,97), pjh/pbk/p_kescape
 * (sm-spline.c:337-445) and the natural-cubic-spline fit (sm-spline.c:63-220), plus the null model
 * per class (scan-chromosome.c:23-37).  fsp[d] = background spectrum of depths[d] (depth+1 doubles,
 * from background_fsp).  coef_out = [n_cls][spline_pts][4], null_out = [n_cls], class layout of
 * fscl_problem_t.  *flags_out: 2 = vanishing-pivot rescue ran, 4 = a coefficient was zeroed
 * (sm-spline.c:67-88, :107-111). */
int fscl_build_tables(int device, int n_depths, const int32_t *depths, const double *const *fsp,
                      int spline_pts, int asc_depth, int asc_min_freq, int ascbias_background_only,
                      int include_invariant, double *coef_out, double *null_out, int32_t *flags_out);
int fscl_set_error(int code, const char *msg);
/* p-value projection described at README.md:11 of the reference and never implemented there; host-only.
 * DEFINITION: lambda = max(mean(stored permuted CLRs) - df, 0), p = P[ncx2(df, lambda) >= clr].
 * permute_clr = [n_pts][FSCL_NULL_SAVE] as filled by fscl_scan_permute; lambda_out may be NULL. */
int fscl_project_pvalues(int n_pts, const fscl_pt_t *pts, const float *permute_clr, double df,
                         double *lambda_out, double *p_out);
int fscl_ncx2_sf(int n, const double *x, double df, const double *lambda, double *out);
/* device-resident variants used by bench.py: time whole scans with inputs in HBM */
int fscl_bench_scan(void *ctx, int large_grid_sp, int n_perm_slots, int count_terms,
                    int64_t *terms_out, fscl_stats_t *stats);
/* search_maxalpha's coarse list (sm-search.c:276-282): by default a candidate whose sum is provably
 * below an exactly evaluated one (upper bound from per-tile class counts) is not
 * evaluated site by site -- the selected alpha, CLR and every later decision are unchanged.
 * on = 1 evaluates every candidate exactly (A/B tests, reference-equivalent term counts). */
int fscl_set_exact_candidates(void *ctx, int on);
/* search_maxpos (scan-chromosome.c:103-139) evaluates both end points of every cell, so a scan evaluates each interior
 * grid point twice.  By default a cell that starts where the previous task of the same genome ended takes that end
 * point's result over instead of evaluating it again (same position, same sites, same order: bit-identical).
 * on = 0 makes every cell evaluate both of its end points itself (A/B tests). */
int fscl_set_endpoint_sharing(void *ctx, int on);
/* Debug hook: the 16 device counters of the context ([0] SNP-terms of COUNT launches, [1] coarse passes
 * repeated exactly; [2..15] phase latencies in cycles, only filled by a library built with -DFSCL_TIMING). */
int fscl_debug_counters(void *ctx, unsigned long long *out16);
/* Libraries built with -DFSCL_HIST: SNP-terms of the term-counting launches by [pass][code path][candidate] (128 values). */
int fscl_debug_histogram(void *ctx, unsigned long long *out128);
int fscl_get_stats(void *ctx, fscl_stats_t *stats);
int fscl_reset_stats(void *ctx);
/* FP64 FMA peak of the device (dependent-chain microbenchmark), in TFLOP/s */
int fscl_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
