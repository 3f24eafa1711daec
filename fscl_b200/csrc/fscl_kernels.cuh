// fscl_kernels.cuh -- device side of the CLR scan + block permutation (sm_100a).
//
// Path replaced (reference slowkoni/fscl, cited file:line):
//   logt                    sm-search.c:40-46
//   spline_interpolate      sm-spline.c:48-60
//   snp_likelihood          sm-search.c:85-103
//   sm_likelihood           sm-search.c:105-150
//   search_maxalpha         sm-search.c:269-300
//   init_scan_result        scan-chromosome.c:58-101
//   search_maxpos(_rec.)    scan-chromosome.c:103-139
//   snp_block_permute       scan-chromosome.c:336-389 (swap phase)
//
// Not a dense contraction: the bound is the FP64 FMA pipe fed by 32-byte
// coefficient gathers that stay in L1/L2; tensor cores are deliberately unused.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fscl {

constexpr int kCoarse = 11;      // coarse alpha candidates (sm-search.c:276-282)
constexpr int kMaxFine = 15;     // fine alpha candidates   (sm-search.c:284-296)
constexpr int kRows = kCoarse + 1;  // fine lists for each coarse argmax + the coarse list (row 11)
#ifndef FSCL_WARPS
#define FSCL_WARPS 8
#endif
#ifndef FSCL_MINBLK
#define FSCL_MINBLK 2
#endif
#ifndef FSCL_TILE
#define FSCL_TILE 1024
#endif
#ifndef FSCL_KB
#define FSCL_KB 4
#endif
constexpr int kScanThreads = FSCL_WARPS * 32;
constexpr int kScanMinBlocks = FSCL_MINBLK;   // up to 128 registers: batch state + 15 fine-alpha accumulators stay in registers
constexpr int kTile = FSCL_TILE;         // sites per class-bucketed tile (<= 4096: 12-bit index in the key)
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kPermThreads = 1024;
#ifndef FSCL_LIM_LAST
#define FSCL_LIM_LAST 5
#endif
#ifndef FSCL_LIM_MID
#define FSCL_LIM_MID 7
#endif
#ifndef FSCL_NB
#define FSCL_NB 4
#endif
constexpr int kNB = FSCL_NB;     // leading coarse candidates a bound may stand for (dev_cell_bounds)
constexpr int kLimLast = FSCL_LIM_LAST, kLimMid = FSCL_LIM_MID;   // extra knots a bound may span (dev_cell_bounds)
#ifndef FSCL_NBF
#define FSCL_NBF 4
#endif
#ifndef FSCL_LIM_FINE
#define FSCL_LIM_FINE 4
#endif
constexpr int kNBF = FSCL_NBF;           // leading FINE candidates a bound may stand for (dev_phase, fine pass)
constexpr int kLimFine = FSCL_LIM_FINE;  // extra knots such a bound may span
#ifndef FSCL_DIRECT
#define FSCL_DIRECT 128
#endif
#ifndef FSCL_CHAIN
#define FSCL_CHAIN 8
#endif
#ifndef FSCL_CHAIN_RAMP
#define FSCL_CHAIN_RAMP 8192
#endif
constexpr int kChainRamp = FSCL_CHAIN_RAMP;   // chains shorten over the last kChainRamp tasks of a launch (about 28 tasks per resident CTA slot)
constexpr int kChain = FSCL_CHAIN;       // neighbouring cells one CTA may take in a row, each starting from the evaluated end point of the last
constexpr int kDirect = FSCL_DIRECT;     // sites either side of the nearest one that narrow-cut candidates are summed from directly

struct Params {
  int n_snps, n_pad, n_chr, n_cls, n_knots, eval_range, bp_resl;
  double inv_step, off_magic;      // knot = low word of x*inv_step + (20/step - 0.5 + 2^52 + 2^51)
  const int *pos;                  // [n_snps]
  const int *chr_start, *chr_nsnps;
  const double *coefC;             // [n_cls][n_knots+2][4], c3 already minus null_tab[cls]; the two
                                   // extra rows repeat the last knot so x == 4 needs no clamp
  int row_doubles;                 // (n_knots + 2) * 4
  const int *tile_lo, *tile_hi;    // [n_pad / kTile] smallest / largest position inside each tile
  const double *logT;              // [65536]
  const double *null_tab;          // [n_cls]
  const double *pmax;              // [levels][n_knots+2][n_cls] entry [l][k][c]: upper bound of (spline piece - null) of class c
                                   // over the knot intervals k .. k+2^l-1 (sparse table for range maxima; class fastest:
                                   // the lanes of a warp walk class runs)
};

// One genome instance (observed, or the permutation slots) as class-bucketed tiles.
struct TileSet {
  int *key, *pos;                  // [slots][n_pad]  class | index_in_tile << 20 ; position
  double *tnull;                   // [slots][n_pad / kTile]  null log-likelihood of the tile
  int *run;                        // [slots][n_pad]  per tile: class | count << 20 for every run of equal classes
  int *nrun;                       // [slots][n_pad / kTile]  runs in the tile
  const int *cls;                  // classes in index order, slot s at cls + s * cls_stride (not owned)
  size_t cls_stride;
};

// device counters of a context (fscl_debug_counters): [0] SNP-terms of COUNT launches, [1] coarse passes repeated exactly,
// [2..15] phase latencies (-DFSCL_TIMING), [16..143] term histogram (-DFSCL_HIST), then the near-tie counters
constexpr int kCtrNearAlpha = 144, kCtrNearPos = 145, kCtrFineRedo = 146, kDebugCounters = 147;

struct Task { int slot, chr, start_pos, end_pos, dst; };   // dst: index into the clr matrix

// Row r of a round that starts at permutation t_first, sharded over `stride` ranks (rank g scores the permutations
// t with t % stride == g as its local rows q = 0, 1, ...): which rank scored it, and as which of its rows.
// One function for the device replay and the host statement of it (tests run the latter over gloo).
__host__ __device__ inline void round_row_owner(int t_first, int stride, int r, int &g, int &q) {
  const int t0m = t_first % stride;
  g = (t0m + r) % stride;
  const int first_r = (g - t0m + stride) % stride;       // first row of the round that rank g scored
  q = (r - first_r) / stride;
}

struct OutPt {
  int chr, nearest, sweep_pos, n_snps, wstart, wend;
  double lalpha, clr, null_logl;
};

// alpha candidate tables, identical for every problem (depend on the LUT only):
// row 11 = coarse list; row m<11 = fine list when the coarse argmax is m.
struct AlphaTables {
  double la[kRows][kMaxFine];
  int dmax[kRows][kMaxFine];     // largest distance with !(logt(d)+la > 4); -1 if none
  int n[kRows];
};

void upload_alpha_tables(const AlphaTables &t);

void launch_scan(const Params &P, const Task *tasks, OutPt *out, const TileSet &ts, int n_tasks, int flags,
                 unsigned long long *term_counter, double *clr_matrix, cudaStream_t s,
                 const int *n_tasks_dev = nullptr, int *queue = nullptr, int n_sm = 0);
void launch_bucket(int n, int n_pad, int n_cls, const int *cls_all, size_t cls_stride, const int *pos,
                   const double *null_tab, const TileSet &ts, int n_slots, cudaStream_t s,
                   const int *skip_if_zero = nullptr);
void launch_permute(int n, const int *cls0, int *pcls_all, int *scratch_all, const int4 *blocks,
                    const int *blk_start, const int *blk_cnt, int n_slots, cudaStream_t s,
                    const int *skip_if_zero = nullptr);
void launch_gen_blocks(int n, const int *pos, const int *chr_end, unsigned long long seed, const int *t_list,
                       double nbp, double width_bp, int4 *blocks, int stride, int *counts, int *flags, int *d7,
                       int n_slots, cudaStream_t s, const int *skip_if_zero = nullptr);
void launch_taskgen(int na_cap, int ns, const int *active, const int *count_dev, const int *pt_chr, const int *pt_sweep,
                    int G, Task *tasks, int *n_tasks_dev, unsigned long long *task_total, cudaStream_t s);
void launch_replay(int na_cap, const int *count_dev, int R, int t_first, int stride, size_t chunk, const int *active,
                   const double *matrix, const double *obs_clr, int *pp, int *pn, int *pf, float *permute_clr,
                   int clr_cap, unsigned long long seed, int prune, int *scored, cudaStream_t s);
void launch_compact(int na_cap, const int *count_in, const int *active_in, const int *pf, int *active_out,
                    int *count_out, int *scored, int R, unsigned long long *scored_total, cudaStream_t s);
void launch_pack_clr(int n_pts, const int *pn, const long long *off, int clr_cap, float *table, float *packed,
                     int to_table, cudaStream_t s);
void launch_project(int n_pts, const int *pn, const double *obs_clr, const float *table, int clr_cap, double df,
                    double *lambda_out, double *p_out, cudaStream_t s);
void launch_fma_peak(double *out, int iters, int grid, cudaStream_t s);

enum { kFlagSingle = 1,      // evaluate start_pos only (search_maxalpha at a point)
       kFlagNullSum = 2,     // also sum null_tab over the final window
       kFlagCount = 4,       // count the SNP-terms the kernel evaluates (roofline numerator)
       kFlagNoPrune = 8,     // evaluate every candidate of the coarse list exactly (no pruning by bounds)
       kFlagNoChain = 16 };  // every cell evaluates both of its end points itself (no sharing between neighbouring cells)

}  // namespace fscl
