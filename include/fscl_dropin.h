/* fscl_dropin.h -- the reference's own scan entry points, implemented on the GPU.
 *
 * libfscl_dropin.so exports the symbols that slowkoni/fscl's scan-chromosome.c and
 * sm-search.c provide to main() (prototypes at /root/reference fscl.h:103-116), with
 * the struct layouts of fscl.h:7-76, so that the reference's unmodified fscl.o,
 * snp-input.o, background-fsp.o, sm-spline.o, asc-bias.o, cmdline-utils.o and
 * logmsg.o link against it in place of scan-chromosome.o + sm-search.o
 * (INTEGRATION.md).  The declarations below are an ABI mirror: field order,
 * types and padding must equal the reference's.
 */
#ifndef FSCL_DROPIN_H
#define FSCL_DROPIN_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {            /* fscl.h:7-14, 32 bytes */
  int chr, pos;
  double null_logl;
  int obs_freq, depth_p, folded;
} snp_t;

typedef struct {            /* fscl.h:26-33 */
  int chr;
  char *name;
  int start_index, n_snps, start_pos, bp_length;
} chr_limits_t;

typedef struct {            /* fscl.h:35-51 */
  int chr, nearest_snp, sweep_pos, n_snps, window_start, window_end;
  double lalpha, null_logl, sm_logl, clr;
  int permute_n, permute_p, permute_finished, scan_running;
  float *permute_clr;
} scan_pt_t;

typedef struct {            /* fscl.h:53-62 */
  int n_snps;
  snp_t *snps;
  int n_depths;
  int *sample_depths;
  int n_scan_pts;
  scan_pt_t *scan_pts;
  chr_limits_t *chr_limits;
  int n_chromosomes;
} scan_t;

typedef struct {            /* fscl.h:64-68 */
  int n;
  double *knot_points;
  double **coef;
} spline_t;

typedef struct {            /* fscl.h:70-76 */
  spline_t **spline_func;
  spline_t **fspline_func;
  int sample_size;
  double **pbk;
  double *fsp;
} sm_ptable_t;

void init_log_table(void);                                                        /* fscl.h:104 <- sm-search.c:14 */
void compute_snp_null_model(scan_t *scan_obj, double **fsp);                      /* fscl.h:108 <- scan-chromosome.c:23 */
void scan_chromosome(scan_t *scan_obj, sm_ptable_t *sm_p, int eval_range, int bp_resl,
                     int large_grid_sp, int n_threads);                           /* fscl.h:109 <- scan-chromosome.c:228 */
void scan_permute(scan_t *scan_obj, sm_ptable_t *sm_p, int n_permute, double permute_nbp,
                  double alpha_factor, int n_threads, int eval_range, int bp_resl,
                  int large_grid_sp, double scan_width_mb);                       /* fscl.h:111 <- scan-chromosome.c:582 */
void scan_output(char *output_fname, scan_t *scan_obj, int maximum_only, int n_permute,
                 char *prepend_label);                                            /* fscl.h:115 <- scan-chromosome.c:666 */
void search_maxalpha(scan_pt_t *scan_pt, snp_t *snps, sm_ptable_t *sm_p);         /* fscl.h:105 <- sm-search.c:269 */

#ifdef __cplusplus
}
#endif
#endif
