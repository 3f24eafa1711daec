// fscl_host.cu -- host side of libfscl_b200.so: context, block-permutation
// generator, permutation rounds, the C-ABI of include/fscl_b200.h.
//
// Reference map (slowkoni/fscl):
//   fscl_scan_chromosome  <- scan_chromosome      scan-chromosome.c:228-266
//   fscl_scan_permute     <- scan_permute(_thread) scan-chromosome.c:412-652
//   BlockGen::generate    <- snp_block_permute    scan-chromosome.c:345-363 (draw phase)
//   GlibcRand             <- glibc 2.39 random_r TYPE_3 (srand fscl.c:135)
#include "../../include/fscl_b200.h"
#include "fscl_kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

using namespace fscl;

static thread_local std::string g_err;
static int fail(int code, const std::string &m) { g_err = m; return code; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(FSCL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

extern "C" const char *fscl_last_error(void) { return g_err.c_str(); }
extern "C" int fscl_set_error(int code, const char *msg) { return fail(code, msg ? msg : ""); }

extern "C" int fscl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// --------------------------------------------------------------------------- RNG
namespace {

struct GlibcRand {            // r[i] = r[i-31] + r[i-3]; output >> 1; 310 warm-up draws
  int32_t r[34];
  long k;
  void seed(unsigned s) {
    int32_t w[344];
    if (s == 0) s = 1;
    w[0] = (int32_t)s;
    for (int i = 1; i < 31; i++) {
      long hi = w[i - 1] / 127773, lo = w[i - 1] % 127773;
      long v = 16807 * lo - 2836 * hi;
      if (v < 0) v += 2147483647;
      w[i] = (int32_t)v;
    }
    for (int i = 31; i < 34; i++) w[i] = w[i - 31];
    for (int i = 34; i < 344; i++) w[i] = (int32_t)((uint32_t)w[i - 31] + (uint32_t)w[i - 3]);
    for (int i = 310; i < 344; i++) r[i % 34] = w[i];
    k = 344;
  }
  int32_t next() {
    uint32_t v = (uint32_t)r[(k - 31) % 34] + (uint32_t)r[(k - 3) % 34];
    r[k % 34] = (int32_t)v;
    k++;
    if (k >= 34L * 1000000) k -= 34L * 999999;
    return (int32_t)(v >> 1);
  }
};

inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 31-bit draw, same range as rand(): counter = (idx, stream, t, 0), key = seed
inline int32_t philox_draw(uint64_t seed, uint32_t t, uint32_t stream, uint32_t idx) {
  uint32_t o[4];
  philox4x32_10(idx, stream, t, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return (int32_t)(o[0] >> 1);
}

struct DrawSource {
  int kind = FSCL_RNG_PHILOX;
  uint64_t seed = 0;
  GlibcRand g;
  uint32_t t = 0, ctr = 0;
  void init(int kind_, uint64_t seed_) {
    kind = kind_; seed = seed_;
    if (kind == FSCL_RNG_GLIBC) g.seed((unsigned)seed);
  }
  void set_perm(uint32_t t_) { t = t_; ctr = 0; }
  int32_t block_draw() {
    if (kind == FSCL_RNG_GLIBC) return g.next();
    if (kind == FSCL_RNG_LIBC) return (int32_t)rand();
    return philox_draw(seed, t, 0, ctr++);
  }
  int32_t prune_draw(uint32_t perm, uint32_t scan_pt) {
    if (kind == FSCL_RNG_GLIBC) return g.next();
    if (kind == FSCL_RNG_LIBC) return (int32_t)rand();
    return philox_draw(seed, perm, 1, scan_pt);
  }
  bool sequential() const { return kind != FSCL_RNG_PHILOX; }
};

struct PermState {
  bool open = false;
  fscl_permute_opts_t o{};
  int stride = 1;
  DrawSource src;
  fscl_pt_t *pts = nullptr;
  int n_pts = 0;
  float *permute_clr = nullptr;
  std::vector<int> active, my_rows, off;
  std::vector<int4> blocks;
  std::vector<double> h_matrix;
  double *d_matrix = nullptr;
  int t_next = 0, R = 0;
  // device-resident rounds (Philox): counters, active list and pruning stay on the GPU
  bool on_device = false;
  int n_active = 0, cur = 0;           // d_active[cur] / d_count[cur] hold the active list and its length
  // pipelined engine: the host runs about one round ahead of the GPU and learns the active count two rounds late
  long round = 0, collected = 0;       // rounds enqueued so far; rounds whose timings the host has read
  int64_t rounds_at_begin = 0;
  int na_round = 0, ns_round = 0;      // launch bound (stale active count) and local permutations of the round in flight
  size_t chunk = 0;                    // doubles per rank in the gathered matrix of the round in flight
  int clr_cap = 0;                     // stored permuted CLRs per cell on the device: min(n_permute + 1, FSCL_NULL_SAVE)
  bool keep_clr = false;
};

// Device memory comes from the stream-ordered pool of the device with an unlimited release
// threshold: creating and destroying a context (one per end-to-end call of the Python layer)
// reuses cached blocks instead of paying cudaMalloc / cudaFree (device-wide synchronisation, tens
// of ms each once other libraries hold mappings).  Callers free only memory no stream still uses.
constexpr int kMaxDevices = 64;
cudaStream_t g_alloc_stream[kMaxDevices] = {};
std::mutex g_alloc_lock;

// FSCL_B200_POOL_KEEP_MB: how much freed memory the device's pool keeps cached (default: all of it; 0 gives
// every block back to the driver as soon as it is free).
cudaStream_t alloc_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;   // null stream: still correct
  std::lock_guard<std::mutex> guard(g_alloc_lock);
  if (!g_alloc_stream[dev]) {
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      if (const char *mb = getenv("FSCL_B200_POOL_KEEP_MB")) keep = strtoull(mb, nullptr, 10) << 20;
      if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
    } else cudaGetLastError();
    g_alloc_stream[dev] = s;
  }
  return g_alloc_stream[dev];
}

template <class T>
cudaError_t dev_malloc(T **p, size_t bytes) {
  cudaStream_t s = alloc_stream();
  cudaError_t e = cudaMallocAsync((void **)p, bytes ? bytes : 1, s);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

// `user` = the stream that used the memory last: waited for (idle in practice: the entry points
// synchronise before they return), then the block goes back on the allocation stream, so the next
// dev_malloc reuses it by plain stream order (a free on another stream is only reused
// opportunistically, and a miss makes the pool map fresh memory: 30-130 ms per context measured)
void dev_free(void *p, cudaStream_t user) {
  if (!p) return;
  cudaStreamSynchronize(user);
  cudaFreeAsync(p, alloc_stream());
}

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  Params P{};
  int n = 0, n_chr = 0, n_cls = 0;
  std::vector<int> h_chr, h_pos, h_cls, chr_start, chr_nsnps, chr_startpos, chr_bplen;
  std::vector<int> chr_end_index;         // per SNP: one past the last index of its chromosome
  // device
  int *d_pos = nullptr, *d_cls0 = nullptr, *d_chr_start = nullptr, *d_chr_nsnps = nullptr;
  double *d_coefC = nullptr, *d_logT = nullptr, *d_null = nullptr;
  int n_pad = 0;
  TileSet ts0{}, ts{};                            // observed genome / permutation slots, class-bucketed
  double *d_pmax = nullptr;                       // piece maxima for the candidate bounds
  bool no_prune = false;                          // evaluate every coarse candidate exactly
  int block_cap = 0;
  size_t slot_memory = (size_t)12 << 30;          // device memory the two slot sets may take together (FSCL_B200_SLOT_GB)
  bool round_log = false; int log_R[4] = {0, 0, 0, 0}, log_na[4] = {0, 0, 0, 0};   // FSCL_B200_ROUND_LOG
  bool no_chain = false;                          // every cell evaluates both end points itself (fscl_set_endpoint_sharing)
  int *d_tile_lo = nullptr, *d_tile_hi = nullptr;
  int max_slots = 0;
  int *d_pcls = nullptr, *d_scratch = nullptr;
  int4 *d_blocks = nullptr; size_t blocks_cap = 0;
  int *d_blkoff = nullptr, *d_blkcnt = nullptr;
  int *d_chr_end = nullptr;                        // per site: one past the last index of its chromosome
  // device-resident permutation test state
  int *d_active[2] = {nullptr, nullptr}, *d_pt_chr = nullptr, *d_pt_sweep = nullptr;
  int *d_pp = nullptr, *d_pn = nullptr, *d_pf = nullptr;
  double *d_pt_clr = nullptr;
  float *d_pclr = nullptr;
  int pts_cap = 0; size_t pclr_floats = 0;
  int *d_small = nullptr;                          // [5 * kSmall]: t_list, rows, flags, d7, scored ; + count
  int small_cap = 0;
  Task *d_tasks = nullptr; OutPt *d_out = nullptr; size_t tasks_cap = 0;
  unsigned long long *d_terms = nullptr;
  int *d_queue = nullptr; int n_sm = 0;           // persistent scan: per-SM task counters (FSCL_B200_PERSISTENT=1)
  std::vector<Task> h_tasks;
  std::vector<OutPt> h_out;
  double *d_matrix = nullptr; size_t matrix_cap = 0;
  // second permutation slot set + what the pipelined rounds need (see device_round_score)
  TileSet ts1{};
  int *d_pcls1 = nullptr, *d_scratch1 = nullptr, *d_blkoff1 = nullptr, *d_blkcnt1 = nullptr;
  int4 *d_blocks1 = nullptr; size_t blocks_cap1 = 0; int max_slots1 = 0;
  cudaStream_t prep = nullptr;                    // block draws, swaps and bucketing of the NEXT round
  cudaEvent_t ev_gen[2] = {}, ev_prep[2] = {}, ev_scan[2] = {};                                 // per slot set
  bool set_used[2] = {false, false};
  cudaEvent_t ev_round[4] = {}, ev_k0[4] = {}, ev_k1[4] = {};                                   // ring by round
  int *h_pin = nullptr;                           // pinned: [4] counts, then per set [3 * kPinSlots] flags / d7 / block counts
  int *d_eng = nullptr;                           // [2] active counts, [1] task count of the round, then per set t_list/flags/d7
  int *d_scored = nullptr; int scored_cap = 0;
  unsigned long long *d_totals = nullptr;         // [0] searches launched, [1] permutations scored
  long long *d_off = nullptr; float *d_pack = nullptr; size_t pack_cap = 0; int off_cap = 0;
  // multi-GPU: NCCL communicator (one rank per context)
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  PermState perm;
  fscl_stats_t stats{};
};

constexpr int kPinSlots = 1024;                    // permutations per rank and round the pinned flag area holds

struct SlotRef {
  TileSet &ts; int *&pcls; int *&scratch; int4 *&blocks; size_t &blocks_cap; int *&blkoff; int *&blkcnt; int &max_slots;
};
SlotRef slot_set(Ctx *c, int k) {
  if (k == 0) return SlotRef{c->ts, c->d_pcls, c->d_scratch, c->d_blocks, c->blocks_cap, c->d_blkoff, c->d_blkcnt, c->max_slots};
  return SlotRef{c->ts1, c->d_pcls1, c->d_scratch1, c->d_blocks1, c->blocks_cap1, c->d_blkoff1, c->d_blkcnt1, c->max_slots1};
}

int ensure_tasks(Ctx *c, size_t n) {
  if (n <= c->tasks_cap) return FSCL_OK;
  size_t cap = std::max<size_t>(n, c->tasks_cap * 2);
  if (c->d_tasks) dev_free(c->d_tasks, c->stream);
  if (c->d_out) dev_free(c->d_out, c->stream);
  c->d_tasks = nullptr; c->d_out = nullptr; c->tasks_cap = 0;
  CK(dev_malloc(&c->d_tasks, cap * sizeof(Task)));
  CK(dev_malloc(&c->d_out, cap * sizeof(OutPt)));
  c->tasks_cap = cap;
  return FSCL_OK;
}

// Upper bound of every spline piece (null already folded into c3) over the x interval its knot
// index stands for on the device: knot k < K covers [-20 + k step, -20 + (k+1) step), the two
// clamp rows K, K+1 repeat piece K-1 (sm-spline.c:53) up to x = 4 and one step beyond.  End
// points, interior critical points, plus a margin far above the rounding of a Horner evaluation.
// A piece with a non-finite coefficient (D14) gets +inf: its candidates are never pruned.
void build_piece_maxima(const double *coefC, int NC, int K, double step, std::vector<double> &ptab) {
  std::vector<double> pmax((size_t)NC * (K + 2), 0.0);
  for (int cl = 0; cl < NC; cl++)
    for (int k = 0; k < K + 2; k++) {
      const double *q = coefC + ((size_t)cl * (K + 2) + k) * 4;
      const double c0 = q[0], c1 = q[1], c2 = q[2], c3 = q[3];
      double &out = pmax[(size_t)cl * (K + 2) + k];
      if (!(std::isfinite(c0) && std::isfinite(c1) && std::isfinite(c2) && std::isfinite(c3))) {
        out = std::numeric_limits<double>::infinity();
        continue;
      }
      const double a = -20.0 + k * step - 1e-9, b = -20.0 + (k + 1) * step + 1e-9;
      auto f = [&](double x) { return x * (x * (x * c0 + c1) + c2) + c3; };
      double m = std::max(f(a), f(b));
      const double qa = 3.0 * c0, qb = 2.0 * c1, qc = c2;          // f'(x) = qa x^2 + qb x + qc
      if (qa != 0.0) {
        const double disc = qb * qb - 4.0 * qa * qc;
        if (disc >= 0.0) {
          const double sq = std::sqrt(disc);
          for (int sg = -1; sg <= 1; sg += 2) {
            const double r = (-qb + sg * sq) / (2.0 * qa);
            if (r > a && r < b) m = std::max(m, f(r));
          }
        }
      } else if (qb != 0.0) {
        const double r = -qc / qb;
        if (r > a && r < b) m = std::max(m, f(r));
      }
      for (int i = 1; i < 8; i++) m = std::max(m, f(a + (b - a) * i / 8.0));   // belt and braces
      const double xm = std::max(std::fabs(a), std::fabs(b));
      const double mag = std::fabs(c3) + xm * (std::fabs(c2) + xm * (std::fabs(c1) + xm * std::fabs(c0)));
      out = m + 1e-11 * mag + 1e-300;
    }
  // sparse table for range maxima over knots, [level][knot][class], level l = knots k .. k+2^l-1
  // (clipped at the last row); floor(log2(K+2))+1 levels cover every range of the K+2 rows
  const size_t lv = (size_t)(K + 2) * NC;
  int n_lv = 1;
  while ((2 << (n_lv - 1)) <= K + 2) n_lv++;
  ptab.assign((size_t)n_lv * lv, 0.0);
  for (int k = 0; k < K + 2; k++)
    for (int cl = 0; cl < NC; cl++) ptab[(size_t)k * NC + cl] = pmax[(size_t)cl * (K + 2) + k];
  for (int l = 1; l < n_lv; l++)
    for (int k = 0; k < K + 2; k++) {
      const int k2 = std::min(k + (1 << (l - 1)), K + 1);
      for (int cl = 0; cl < NC; cl++)
        ptab[l * lv + (size_t)k * NC + cl] = std::max(ptab[(l - 1) * lv + (size_t)k * NC + cl], ptab[(l - 1) * lv + (size_t)k2 * NC + cl]);
    }
}

void free_tiles(TileSet &t, cudaStream_t s) {
  dev_free(t.key, s); dev_free(t.pos, s); dev_free(t.tnull, s); dev_free(t.run, s); dev_free(t.nrun, s);
  t = TileSet{};
}

int alloc_tiles(TileSet &t, int slots, int n_pad) {
  CK(dev_malloc(&t.tnull, (size_t)slots * (n_pad / kTile) * sizeof(double)));
  CK(dev_malloc(&t.nrun, (size_t)slots * (n_pad / kTile) * sizeof(int)));
  CK(dev_malloc(&t.key, (size_t)slots * n_pad * sizeof(int)));
  CK(dev_malloc(&t.pos, (size_t)slots * n_pad * sizeof(int)));
  CK(dev_malloc(&t.run, (size_t)slots * n_pad * sizeof(int)));
  return FSCL_OK;
}

int ensure_slots(Ctx *c, int slots, int set = 0) {
  SlotRef Q = slot_set(c, set);
  if (slots <= Q.max_slots) return FSCL_OK;
  // growing a set that kernels may still read: wait for both streams first (rare: sets are sized at begin)
  if (c->prep) CK(cudaStreamSynchronize(c->prep));
  if (Q.pcls) dev_free(Q.pcls, c->stream);
  if (Q.scratch) dev_free(Q.scratch, c->stream);
  if (Q.blkoff) dev_free(Q.blkoff, c->stream);
  if (Q.blkcnt) dev_free(Q.blkcnt, c->stream);
  Q.blkcnt = nullptr;
  free_tiles(Q.ts, c->stream);
  Q.pcls = Q.scratch = Q.blkoff = nullptr; Q.max_slots = 0;
  { int rc = alloc_tiles(Q.ts, slots, c->n_pad); if (rc) return rc; }
  CK(dev_malloc(&Q.pcls, (size_t)slots * c->n * sizeof(int)));
  Q.ts.cls = Q.pcls; Q.ts.cls_stride = (size_t)c->n;
  CK(dev_malloc(&Q.scratch, (size_t)slots * c->n * sizeof(int)));
  CK(dev_malloc(&Q.blkoff, (size_t)(slots + 1) * sizeof(int)));
  CK(dev_malloc(&Q.blkcnt, (size_t)(slots + 1) * sizeof(int)));
  Q.max_slots = slots;
  return FSCL_OK;
}

static inline int scan_flags(const Ctx *c) { return (c->no_prune ? kFlagNoPrune : 0) | (c->no_chain ? kFlagNoChain : 0); }

// run the scan kernel over c->h_tasks; results land in c->h_out
int run_tasks(Ctx *c, bool permuted, int flags, bool fetch = true, double *d_clr = nullptr) {
  size_t nt = c->h_tasks.size();
  if (nt == 0) { c->h_out.clear(); return FSCL_OK; }
  int rc = ensure_tasks(c, nt);
  if (rc) return rc;
  CK(cudaMemcpyAsync(c->d_tasks, c->h_tasks.data(), nt * sizeof(Task), cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(c->ev0, c->stream));
  launch_scan(c->P, c->d_tasks, c->d_out, permuted ? c->ts : c->ts0, (int)nt, flags | scan_flags(c),
              c->d_terms, d_clr, c->stream, nullptr, c->d_queue, c->n_sm);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev1, c->stream));
  if (fetch) {
    c->h_out.resize(nt);
    CK(cudaMemcpyAsync(c->h_out.data(), c->d_out, nt * sizeof(OutPt), cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stats.score_kernel_ms += ms;
  c->stats.score_launches++;
  c->stats.n_tasks += (int64_t)nt;
  return FSCL_OK;
}

void to_pt(const OutPt &o, fscl_pt_t &p) {
  p.chr = o.chr; p.nearest_snp = o.nearest; p.sweep_pos = o.sweep_pos; p.n_snps = o.n_snps;
  p.window_start = o.wstart; p.window_end = o.wend; p.lalpha = o.lalpha; p.clr = o.clr;
  p.null_logl = o.null_logl; p.sm_logl = o.null_logl + 0.5 * o.clr;
  p.permute_n = p.permute_p = p.permute_finished = 0; p.reserved = 0;
}

int enum_cells_raw(int n_chr, const int *startpos, const int *bplen, int G, std::vector<Task> *out) {
  // scan_thread's dispenser, scan-chromosome.c:178-205
  int chm = 0, n = 0;
  if (n_chr == 0) return 0;
  int pos = startpos[0];
  for (;;) {
    if (chm == n_chr) break;
    if (pos >= bplen[chm]) {
      chm++;
      if (chm == n_chr) break;
      pos = startpos[chm];
    }
    int s = pos, e = pos + G;
    pos += G;
    if (e > bplen[chm]) e = bplen[chm];
    if (out) out->push_back(Task{0, chm, s, e, 0});
    n++;
  }
  return n;
}
int enum_cells(const Ctx *c, int G, std::vector<Task> *out) {
  return enum_cells_raw(c->n_chr, c->chr_startpos.data(), c->chr_bplen.data(), G, out);
}

// Counting / stochastic pruning of one round, in permutation order
// (scan-chromosome.c:488-498).  matrix is [R][n_active]; live[a] drops to 0
// when a point finishes inside the round.  Returns permutations that scored.
int replay_round(fscl_pt_t *pts, const int *active, size_t na, const double *matrix, int R, int t_first,
                 DrawSource &src, int prune, float *permute_clr) {
  std::vector<char> live(na, 1);
  int scored = 0;
  for (int r = 0; r < R; r++) {
    const int t = t_first + r;
    bool any = false;
    for (size_t a = 0; a < na; a++) {
      if (!live[a]) continue;
      any = true;
      fscl_pt_t &p = pts[active[a]];
      const double clr = matrix[(size_t)r * na + a];
      if (clr >= p.clr) {
        p.permute_p++;
        if (prune && p.permute_p >= FSCL_PRUNE_MIN_P) {                 // D6
          int32_t u = src.prune_draw((uint32_t)t, (uint32_t)active[a]);
          if (p.permute_p / (double)p.permute_n >= u / (2147483647 + 1.0)) { p.permute_finished = 1; live[a] = 0; }
        }
      }
      if (permute_clr && p.permute_n < FSCL_NULL_SAVE)
        permute_clr[(size_t)active[a] * FSCL_NULL_SAVE + p.permute_n] = (float)clr;
      p.permute_n++;
    }
    if (!any) break;
    scored++;
  }
  return scored;
}

// Draw phase of snp_block_permute (scan-chromosome.c:345-363).  Emits
// (i, j, len) with len = number of swaps the reference's loop :365 performs.
// D7 (k > n_snps; the reference then indexes below the array) is DEFINED as
// "shift the block so that it ends at n_snps" and counted.
struct BlockGen {
  int n;
  const int *pos;
  const int *chr_end_index;      // per site: one past the last index of its chromosome
  double nbp, width_bp;
  int generate(DrawSource &src, std::vector<int4> &out, int *d7) const {
    int i = 0, nb = 0;
    while (i < n) {
      int j = (int)(src.block_draw() / (2147483647 + 1.0) * n);
      double len = -1.0 / nbp * std::log(src.block_draw() / (2147483647 + 1.0));
      int k = j + (len < 2147483647.0 ? (int)len : 0x3FFFFFFF);
      if (k < n && k >= 0) {
        // while (k < n && chr[k]==chr[j] && pos[k]-pos[j] < width) k++  -- monotone, so search
        int lim = chr_end_index[j];
        if (k < lim) {
          int lo = k, hi = lim;              // first index in [k, lim] failing pos[.]-pos[j] < width
          while (lo < hi) {
            int m = lo + (hi - lo) / 2;
            if ((double)(pos[m] - pos[j]) < width_bp) lo = m + 1; else hi = m;
          }
          k = lo;
        }
      }
      if (i + (k - j) >= n) k = n;
      if (k > n) { j = n - (k - j); k = n; if (d7) (*d7)++; }
      int L = std::min(k - j, n - i);
      if (L > 0) { out.push_back(make_int4(i, j, L, 0)); nb++; i += L; }
    }
    return nb;
  }
};

}  // namespace

// NCCL, resolved at run time (see the multi-GPU section below)
namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
} g_nccl;
std::mutex g_nccl_lock;

int nccl_load() {
  std::lock_guard<std::mutex> guard(g_nccl_lock);
  if (g_nccl.h) return FSCL_OK;
  void *h = nullptr;
  if (const char *path = getenv("FSCL_B200_NCCL")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(FSCL_ERR_ARG, std::string("cannot load libnccl.so.2: ") + dlerror());
#define FSCL_SYM(field, name)                                                                        \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                                        \
  if (!g_nccl.field) { dlclose(h); return fail(FSCL_ERR_ARG, std::string("libnccl lacks ") + name); }
  FSCL_SYM(GetUniqueId, "ncclGetUniqueId")
  FSCL_SYM(CommInitRank, "ncclCommInitRank")
  FSCL_SYM(CommInitAll, "ncclCommInitAll")
  FSCL_SYM(CommDestroy, "ncclCommDestroy")
  FSCL_SYM(AllGather, "ncclAllGather")
  FSCL_SYM(GetErrorString, "ncclGetErrorString")
  FSCL_SYM(GroupStart, "ncclGroupStart")
  FSCL_SYM(GroupEnd, "ncclGroupEnd")
#undef FSCL_SYM
  g_nccl.h = h;
  return FSCL_OK;
}
}  // namespace

static const char *nccl_error(ncclResult_t r) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"; }
#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) return fail(FSCL_ERR_CUDA, std::string(#call) + ": " + nccl_error(r_)); \
  } while (0)

// in place: rank g's bytes_per_rank bytes sit at buf + g * bytes_per_rank
static int nccl_all_gather(Ctx *c, void *buf, size_t bytes_per_rank) {
  if (!c->comm || c->world <= 1 || bytes_per_rank == 0) return FSCL_OK;
  NK(g_nccl.AllGather((const char *)buf + (size_t)c->rank * bytes_per_rank, buf, bytes_per_rank, ncclChar, c->comm, c->stream));
  return FSCL_OK;
}

// --------------------------------------------------------------------------- create / destroy
static void build_alpha_tables(const double *T, AlphaTables &A) {
  auto logt = [&](long d) -> double {
    if (d > 0xFFFFFF) return 11.783502069519070 + T[d >> 16];
    if (d > 0xFFFF) return 5.545177444479562 + T[d >> 8];
    return T[d];
  };
  auto dmax = [&](double la) -> int {
    if (logt(0) + la > 4.0) return -1;
    long lo = 0, hi = 0x7FFFFFFFL;
    if (!(logt(hi) + la > 4.0)) return 0x7FFFFFFE;   // capped: INT_MAX marks records outside the window
    while (hi - lo > 1) { long m = (lo + hi) / 2; if (logt(m) + la > 4.0) hi = m; else lo = m; }
    return (int)lo;
  };
  memset(&A, 0, sizeof A);
  // coarse list: FP-accumulated exactly as sm-search.c:276-279
  double step = (4.0 - (-20.0)) / 10.0;
  int nc = 0;
  for (double la = -20.0; la <= 4.0; la += step) {
    if (nc < kMaxFine) { A.la[kCoarse][nc] = la; A.dmax[kCoarse][nc] = dmax(la); }
    nc++;
  }
  A.n[kCoarse] = std::min(nc, kCoarse);
  for (int m = 0; m < kCoarse; m++) {                 // sm-search.c:284-296
    double l = A.la[kCoarse][m] - step; if (l < -20.0) l = -20.0;
    double r = A.la[kCoarse][m] + step; if (r > 4.0) r = 4.0;
    double s = (r - l) / 15.;
    int k = 0;
    for (double la = l + s; la < r; la += s) {
      if (k < kMaxFine) { A.la[m][k] = la; A.dmax[m][k] = dmax(la); }
      k++;
    }
    A.n[m] = std::min(k, kMaxFine);
  }
}

extern "C" void fscl_destroy(void *ctx);

extern "C" int fscl_create(const fscl_problem_t *p, void **ctx_out) {
  if (!p || !ctx_out) return fail(FSCL_ERR_ARG, "null argument");
  *ctx_out = nullptr;
  int ndev = fscl_device_count();
  if (ndev <= 0) return fail(FSCL_ERR_NODEVICE, "no CUDA device: fscl_b200 has no CPU path");
  if (p->device < 0 || p->device >= ndev) return fail(FSCL_ERR_ARG, "bad device ordinal");
  if (p->n_snps <= 0 || p->n_chr <= 0 || p->n_cls <= 0 || p->n_knots <= 0)
    return fail(FSCL_ERR_ARG, "empty problem");
  if (p->eval_range < 1 || p->eval_range > 81920)
    return fail(FSCL_ERR_ARG, "eval_range must be in [1, 81920] (fscl.c:175)");
  CK(cudaSetDevice(p->device));
  int cc_major = 0;      // one attribute, not cudaGetDeviceProperties: that call costs 10-200 ms per context
  CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, p->device));
  if (cc_major < 10) return fail(FSCL_ERR_NODEVICE, "device is not sm_100 class");
  Ctx *c = new Ctx();
  struct Guard {          // a failure below releases whatever the context already owns
    Ctx *c;
    ~Guard() { if (c) fscl_destroy(c); }
  } guard{c};
  c->device = p->device;
  c->n = p->n_snps; c->n_chr = p->n_chr; c->n_cls = p->n_cls;
  c->h_chr.assign(p->chr, p->chr + c->n);
  c->h_pos.assign(p->pos, p->pos + c->n);
  c->h_cls.assign(p->cls, p->cls + c->n);
  c->chr_start.assign(p->chr_start, p->chr_start + c->n_chr);
  c->chr_nsnps.assign(p->chr_nsnps, p->chr_nsnps + c->n_chr);
  c->chr_startpos.assign(p->chr_startpos, p->chr_startpos + c->n_chr);
  c->chr_bplen.assign(p->chr_bplen, p->chr_bplen + c->n_chr);
  for (int i = 0; i < c->n; i++)
    if (c->h_cls[i] < 0 || c->h_cls[i] >= c->n_cls) { return fail(FSCL_ERR_ARG, "class index out of range"); }
  c->chr_end_index.resize(c->n);
  for (int i = c->n - 1; i >= 0; i--)
    c->chr_end_index[i] = (i == c->n - 1 || c->h_chr[i + 1] != c->h_chr[i]) ? i + 1 : c->chr_end_index[i + 1];

  const bool timing = getenv("FSCL_B200_TIMING") != nullptr;
  auto now_ms = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
  const double tm0 = now_ms();
  std::vector<double> T(65536);
  if (p->log_table) memcpy(T.data(), p->log_table, sizeof(double) * 65536);
  else { T[0] = 0.0; for (int i = 1; i <= 0xFFFF; i++) T[i] = std::log((double)i); }   // sm-search.c:14-26
  AlphaTables A;
  build_alpha_tables(T.data(), A);

  // class-major table [cls][knot][4] with the null model folded into c3
  const int K = p->n_knots, NC = p->n_cls;
  if (NC >= (1 << 20)) { return fail(FSCL_ERR_ARG, "more than 2^20 classes"); }
  if ((size_t)NC * (K + 2) * 32 >= ((size_t)1 << 32)) { return fail(FSCL_ERR_ARG, "spline table over 4 GiB"); }
  std::vector<double> coefC((size_t)(K + 2) * NC * 4);
  for (int cl = 0; cl < NC; cl++)
    for (int k = 0; k < K + 2; k++) {
      const double *s = p->coef + ((size_t)cl * K + std::min(k, K - 1)) * 4;   // sm-spline.c:53 clamp as rows
      double *d = coefC.data() + ((size_t)cl * (K + 2) + k) * 4;
      d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3] - p->null_tab[cl];
    }
  std::vector<double> pmax;
  build_piece_maxima(coefC.data(), NC, K, p->step, pmax);
  c->no_prune = getenv("FSCL_B200_NO_PRUNE") != nullptr;
  c->no_chain = getenv("FSCL_B200_NO_ENDPOINT_SHARING") != nullptr;
  c->round_log = getenv("FSCL_B200_ROUND_LOG") != nullptr;
  if (const char *e = getenv("FSCL_B200_BLOCK_CAP")) c->block_cap = std::max(1, atoi(e));
  if (const char *e = getenv("FSCL_B200_SLOT_GB")) c->slot_memory = (size_t)std::min(160, std::max(1, atoi(e))) << 30;
  c->n_pad = (c->n + kTile - 1) / kTile * kTile;
  const int n_tiles = c->n_pad / kTile;
  std::vector<int> tlo(n_tiles), thi(n_tiles);
  for (int t = 0; t < n_tiles; t++) {
    int lo = 0x7FFFFFFF, hi = -0x7FFFFFFF;
    for (int i = t * kTile; i < std::min(c->n, (t + 1) * kTile); i++) { lo = std::min(lo, c->h_pos[i]); hi = std::max(hi, c->h_pos[i]); }
    tlo[t] = lo; thi[t] = hi;
  }

  const double tm1 = now_ms();
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(cudaEventCreate(&c->ev2));
  CK(cudaEventCreate(&c->ev3));
  CK(dev_malloc(&c->d_pos, sizeof(int) * c->n));
  CK(dev_malloc(&c->d_cls0, sizeof(int) * c->n));
  CK(dev_malloc(&c->d_chr_start, sizeof(int) * c->n_chr));
  CK(dev_malloc(&c->d_chr_nsnps, sizeof(int) * c->n_chr));
  CK(dev_malloc(&c->d_coefC, sizeof(double) * coefC.size()));
  { int rc = alloc_tiles(c->ts0, 1, c->n_pad); if (rc) return rc; }
  c->ts0.cls = c->d_cls0; c->ts0.cls_stride = 0;
  CK(dev_malloc(&c->d_pmax, sizeof(double) * pmax.size()));
  CK(cudaMemcpy(c->d_pmax, pmax.data(), sizeof(double) * pmax.size(), cudaMemcpyHostToDevice));
  CK(dev_malloc(&c->d_tile_lo, sizeof(int) * n_tiles));
  CK(dev_malloc(&c->d_tile_hi, sizeof(int) * n_tiles));
  CK(cudaMemcpy(c->d_tile_lo, tlo.data(), sizeof(int) * n_tiles, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_tile_hi, thi.data(), sizeof(int) * n_tiles, cudaMemcpyHostToDevice));
  CK(dev_malloc(&c->d_logT, sizeof(double) * 65536));
  CK(dev_malloc(&c->d_null, sizeof(double) * NC));
  CK(dev_malloc(&c->d_terms, kDebugCounters * sizeof(unsigned long long)));   // [0] SNP-terms (COUNT launches), [1] coarse passes repeated exactly
  CK(dev_malloc(&c->d_chr_end, sizeof(int) * c->n));
  {
    // persistent scan (one CTA per SM slot looping over neighbouring cells): measured +4 % on c2, -2 % on c4 -- opt-in
    const char *pe = getenv("FSCL_B200_PERSISTENT");
    if (pe && atoi(pe) != 0) {
      CK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, p->device));
      CK(dev_malloc(&c->d_queue, sizeof(int) * (size_t)std::max(c->n_sm, 1)));
    }
  }
  CK(cudaMemcpy(c->d_chr_end, c->chr_end_index.data(), sizeof(int) * c->n, cudaMemcpyHostToDevice));
  CK(cudaMemset(c->d_terms, 0, kDebugCounters * sizeof(unsigned long long)));
  CK(cudaMemcpy(c->d_pos, c->h_pos.data(), sizeof(int) * c->n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_cls0, c->h_cls.data(), sizeof(int) * c->n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_chr_start, c->chr_start.data(), sizeof(int) * c->n_chr, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_chr_nsnps, c->chr_nsnps.data(), sizeof(int) * c->n_chr, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_coefC, coefC.data(), sizeof(double) * coefC.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_logT, T.data(), sizeof(double) * 65536, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_null, p->null_tab, sizeof(double) * NC, cudaMemcpyHostToDevice));
  const double tm2 = now_ms();
  upload_alpha_tables(A);
  CK(cudaGetLastError());

  Params &P = c->P;
  P.n_snps = c->n; P.n_pad = c->n_pad; P.n_chr = c->n_chr; P.n_cls = NC; P.n_knots = K;
  P.eval_range = p->eval_range; P.bp_resl = p->bp_resl;
  P.inv_step = 1.0 / p->step; P.off_magic = 20.0 / p->step - 0.5 + 6755399441055744.0;
  P.pos = c->d_pos; P.chr_start = c->d_chr_start; P.chr_nsnps = c->d_chr_nsnps;
  P.coefC = c->d_coefC; P.row_doubles = (K + 2) * 4; P.tile_lo = c->d_tile_lo; P.tile_hi = c->d_tile_hi; P.logT = c->d_logT; P.null_tab = c->d_null; P.pmax = c->d_pmax;
  launch_bucket(c->n, c->n_pad, c->P.n_cls, c->d_cls0, 0, c->d_pos, c->d_null, c->ts0, 1, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  if (timing)
    fprintf(stderr, "fscl_create: host tables %.2f ms, stream + alloc + copies %.2f ms, constants + bucket %.2f ms\n", tm1 - tm0,
            tm2 - tm1, now_ms() - tm2);
  guard.c = nullptr;
  *ctx_out = c;
  return FSCL_OK;
}

extern "C" void fscl_destroy(void *ctx) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  dev_free(c->d_pos, c->stream); dev_free(c->d_cls0, c->stream); dev_free(c->d_chr_start, c->stream); dev_free(c->d_chr_nsnps, c->stream);
  dev_free(c->d_coefC, c->stream); free_tiles(c->ts0, c->stream); free_tiles(c->ts, c->stream); dev_free(c->d_pmax, c->stream);
  dev_free(c->d_tile_lo, c->stream); dev_free(c->d_tile_hi, c->stream); dev_free(c->d_logT, c->stream); dev_free(c->d_null, c->stream); dev_free(c->d_terms, c->stream);
  dev_free(c->d_pcls, c->stream); dev_free(c->d_scratch, c->stream); dev_free(c->d_blocks, c->stream); dev_free(c->d_blkoff, c->stream);
  dev_free(c->d_tasks, c->stream); dev_free(c->d_out, c->stream); dev_free(c->d_matrix, c->stream);
  dev_free(c->d_blkcnt, c->stream); dev_free(c->d_chr_end, c->stream); dev_free(c->d_active[0], c->stream); dev_free(c->d_active[1], c->stream);
  dev_free(c->d_pt_chr, c->stream); dev_free(c->d_pt_sweep, c->stream); dev_free(c->d_pp, c->stream); dev_free(c->d_pn, c->stream); dev_free(c->d_pf, c->stream);
  dev_free(c->d_pt_clr, c->stream); dev_free(c->d_pclr, c->stream); dev_free(c->d_small, c->stream);
  dev_free(c->d_queue, c->stream);
  if (c->prep) cudaStreamSynchronize(c->prep);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  free_tiles(c->ts1, c->stream);
  dev_free(c->d_pcls1, c->stream); dev_free(c->d_scratch1, c->stream); dev_free(c->d_blkoff1, c->stream);
  dev_free(c->d_blkcnt1, c->stream); dev_free(c->d_blocks1, c->stream); dev_free(c->d_eng, c->stream);
  dev_free(c->d_scored, c->stream); dev_free(c->d_totals, c->stream); dev_free(c->d_off, c->stream); dev_free(c->d_pack, c->stream);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  for (int k = 0; k < 2; k++) {
    if (c->ev_gen[k]) cudaEventDestroy(c->ev_gen[k]);
    if (c->ev_prep[k]) cudaEventDestroy(c->ev_prep[k]);
    if (c->ev_scan[k]) cudaEventDestroy(c->ev_scan[k]);
  }
  for (int k = 0; k < 4; k++) {
    if (c->ev_round[k]) cudaEventDestroy(c->ev_round[k]);
    if (c->ev_k0[k]) cudaEventDestroy(c->ev_k0[k]);
    if (c->ev_k1[k]) cudaEventDestroy(c->ev_k1[k]);
  }
  if (c->prep) cudaStreamDestroy(c->prep);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev2) cudaEventDestroy(c->ev2);
  if (c->ev3) cudaEventDestroy(c->ev3);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// --------------------------------------------------------------------------- forward scan
extern "C" int fscl_enum_cells(void *ctx, int G, int32_t *cell_chr, int32_t *cell_start, int32_t *cell_end,
                               int cap) {
  Ctx *c = (Ctx *)ctx;
  if (!c || G <= 0) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  std::vector<Task> cells;
  int n = enum_cells(c, G, &cells);
  for (int i = 0; i < n && i < cap; i++) {
    if (cell_chr) cell_chr[i] = cells[i].chr;
    if (cell_start) cell_start[i] = cells[i].start_pos;
    if (cell_end) cell_end[i] = cells[i].end_pos;
  }
  return n;
}

static int eval_list(Ctx *c, int n, const int32_t *chr, const int32_t *start, const int32_t *end,
                     fscl_pt_t *out, int flags) {
  if (!c || n < 0 || (n > 0 && (!chr || !start || !out))) return fail(FSCL_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  c->h_tasks.resize(n);
  for (int i = 0; i < n; i++) {
    if (chr[i] < 0 || chr[i] >= c->n_chr) return fail(FSCL_ERR_ARG, "chromosome index out of range");
    c->h_tasks[i] = Task{0, chr[i], start[i], end ? end[i] : start[i], 0};
  }
  int rc = run_tasks(c, false, flags | kFlagNullSum);
  if (rc) return rc;
  for (int i = 0; i < n; i++) to_pt(c->h_out[i], out[i]);
  c->stats.n_grid_evals += n;
  return FSCL_OK;
}

extern "C" int fscl_grid_eval(void *ctx, int n, const int32_t *chr, const int32_t *pos, fscl_pt_t *out) {
  return eval_list((Ctx *)ctx, n, chr, pos, nullptr, out, kFlagSingle);
}

extern "C" int fscl_search_maxpos(void *ctx, int n, const int32_t *chr, const int32_t *start_pos,
                                  const int32_t *end_pos, fscl_pt_t *out) {
  if (!end_pos) return fail(FSCL_ERR_ARG, "bad argument");
  return eval_list((Ctx *)ctx, n, chr, start_pos, end_pos, out, 0);
}

extern "C" int fscl_scan_chromosome(void *ctx, int G, fscl_pt_t *out, int cap, int *n_out) {
  Ctx *c = (Ctx *)ctx;
  if (!c || G <= 0 || !out || !n_out) return fail(FSCL_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  c->h_tasks.clear();
  int n = enum_cells(c, G, &c->h_tasks);
  *n_out = n;
  if (n > cap) return fail(FSCL_ERR_ARG, "output capacity too small");
  int rc;
  if (c->comm && c->world > 1 && n > 0) {
    // cells shard over the ranks of the communicator (scan_thread's cells are independent,
    // scan-chromosome.c:162-216); one all-gather of the result records makes every rank complete
    const size_t chunk = ((size_t)n + c->world - 1) / c->world;
    const size_t lo = std::min((size_t)n, chunk * c->rank), hi = std::min((size_t)n, lo + chunk);
    rc = ensure_tasks(c, chunk * c->world); if (rc) return rc;
    if (hi > lo) {
      CK(cudaMemcpyAsync(c->d_tasks, c->h_tasks.data() + lo, (hi - lo) * sizeof(Task), cudaMemcpyHostToDevice, c->stream));
      CK(cudaEventRecord(c->ev0, c->stream));
      launch_scan(c->P, c->d_tasks, c->d_out + lo, c->ts0, (int)(hi - lo), kFlagNullSum | scan_flags(c),
                  c->d_terms, nullptr, c->stream, nullptr, c->d_queue, c->n_sm);
      c->stats.kernel_launches++;
      CK(cudaGetLastError());
      CK(cudaEventRecord(c->ev1, c->stream));
    }
    rc = nccl_all_gather(c, c->d_out, chunk * sizeof(OutPt)); if (rc) return rc;
    c->h_out.resize(n);
    CK(cudaMemcpyAsync(c->h_out.data(), c->d_out, (size_t)n * sizeof(OutPt), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (hi > lo) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
      c->stats.score_kernel_ms += ms; c->stats.score_launches++; c->stats.n_tasks += (int64_t)(hi - lo);
    }
  } else {
    rc = run_tasks(c, false, kFlagNullSum);
    if (rc) return rc;
  }
  std::vector<fscl_pt_t> pts(n);
  for (int i = 0; i < n; i++) to_pt(c->h_out[i], pts[i]);
  // qsort by (chr, sweep_pos), scan-chromosome.c:264 (stable here)
  std::stable_sort(pts.begin(), pts.end(), [](const fscl_pt_t &a, const fscl_pt_t &b) {
    return a.chr != b.chr ? a.chr < b.chr : a.sweep_pos < b.sweep_pos;
  });
  memcpy(out, pts.data(), sizeof(fscl_pt_t) * n);
  return FSCL_OK;
}

// --------------------------------------------------------------------------- permutations
static int upload_blocks(Ctx *c, const std::vector<int4> &blocks, const std::vector<int> &off) {
  if (blocks.size() > c->blocks_cap) {
    if (c->d_blocks) dev_free(c->d_blocks, c->stream);
    c->d_blocks = nullptr; c->blocks_cap = 0;
    size_t cap = blocks.size() * 2 + 1024;
    CK(dev_malloc(&c->d_blocks, cap * sizeof(int4)));
    c->blocks_cap = cap;
  }
  if (!blocks.empty())
    CK(cudaMemcpyAsync(c->d_blocks, blocks.data(), blocks.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
  std::vector<int> cnt(off.size(), 0);
  for (size_t q = 0; q + 1 < off.size(); q++) cnt[q] = off[q + 1] - off[q];
  CK(cudaMemcpyAsync(c->d_blkoff, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_blkcnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));          // cnt is a local
  return FSCL_OK;
}

static int apply_blocks(Ctx *c, int n_slots) {
  CK(cudaEventRecord(c->ev0, c->stream));
  launch_permute(c->n, c->d_cls0, c->d_pcls, c->d_scratch, c->d_blocks, c->d_blkoff, c->d_blkcnt, n_slots, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  launch_bucket(c->n, c->n_pad, c->P.n_cls, c->d_pcls, (size_t)c->n, c->d_pos, c->d_null, c->ts,
                n_slots, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stats.permute_kernel_ms += ms;
  c->stats.permute_launches++;
  return FSCL_OK;
}

extern "C" int fscl_block_permute(void *ctx, const fscl_permute_opts_t *o, int t, int32_t *cls_out,
                                  int32_t *blocks_out, int max_blocks) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !o || !cls_out || t < 0) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  if (cudaSetDevice(c->device) != cudaSuccess) { fail(FSCL_ERR_CUDA, "cudaSetDevice"); return -1; }
  if (ensure_slots(c, 1)) return -1;
  DrawSource src;
  src.init(o->rng_kind, o->seed);
  BlockGen gen{c->n, c->h_pos.data(), c->chr_end_index.data(), o->permute_nbp, o->scan_width_mb * 1e6};
  std::vector<int4> blocks;
  int d7 = 0;
  if (src.sequential()) {
    if (o->rng_kind != FSCL_RNG_LIBC) (void)src.block_draw();   // usleep burn, scan-chromosome.c:440
    for (int q = 0; q <= t; q++) { blocks.clear(); gen.generate(src, blocks, &d7); }
  } else {
    src.set_perm((uint32_t)t);
    gen.generate(src, blocks, &d7);
  }
  std::vector<int> off = {0, (int)blocks.size()};
  if (upload_blocks(c, blocks, off)) return -1;
  if (apply_blocks(c, 1)) return -1;
  if (cudaMemcpy(cls_out, c->d_pcls, sizeof(int) * c->n, cudaMemcpyDeviceToHost) != cudaSuccess) {
    fail(FSCL_ERR_CUDA, "copy back"); return -1;
  }
  for (size_t b = 0; b < blocks.size() && blocks_out && (int)b < max_blocks; b++) {
    blocks_out[3 * b] = blocks[b].x; blocks_out[3 * b + 1] = blocks[b].y; blocks_out[3 * b + 2] = blocks[b].z;
  }
  return (int)blocks.size();
}

// ---- device-resident, pipelined rounds (counter-based stream) -------------------------------------
//
// A round = R consecutive permutations.  Everything a round needs from the previous one (the active
// list and its length, the counters) stays on the device; the host only enqueues.  Two streams:
//
//   prep   : [wait: scan of round r-2 released this slot set]  gen_blocks(r) -> permute(r) -> bucket(r)
//   stream : [wait: prep(r)]  taskgen(r) -> scan_cells(r) -> [all-gather over the communicator] ->
//            replay(r) -> compact(r) -> count to pinned memory
//
// The permuted genomes of round r do not depend on the outcome of round r-1 (only the active list
// does), so prep(r) overlaps scan(r-1): two slot sets alternate.  The host sizes the launches of round
// r from the active count after round r-2 (read from pinned memory, long complete by then); the
// kernels take the true count from device memory and surplus CTAs leave at once.  The only host waits
// are for events that completed a round earlier and for the flags of gen_blocks(r) (about a millisecond,
// while scan(r-1) keeps the GPU busy), so between two scan kernels the stream holds only
// replay + compact + taskgen.  Every rank of a sharded run computes the same R from the same stale
// count, so the round boundaries agree without communication.
static int ensure_pts(Ctx *c, int n_pts) {
  if (n_pts > c->pts_cap) {
    for (int *&p : {std::ref(c->d_active[0]), std::ref(c->d_active[1]), std::ref(c->d_pt_chr), std::ref(c->d_pt_sweep),
                    std::ref(c->d_pp), std::ref(c->d_pn), std::ref(c->d_pf)}) { if (p) dev_free(p, c->stream); p = nullptr; }
    if (c->d_pt_clr) dev_free(c->d_pt_clr, c->stream);
    c->d_pt_clr = nullptr; c->pts_cap = 0;
    const size_t b = sizeof(int) * (size_t)n_pts;
    CK(dev_malloc(&c->d_active[0], b)); CK(dev_malloc(&c->d_active[1], b));
    CK(dev_malloc(&c->d_pt_chr, b)); CK(dev_malloc(&c->d_pt_sweep, b));
    CK(dev_malloc(&c->d_pp, b)); CK(dev_malloc(&c->d_pn, b)); CK(dev_malloc(&c->d_pf, b));
    CK(dev_malloc(&c->d_pt_clr, sizeof(double) * (size_t)n_pts));
    c->pts_cap = n_pts;
  }
  return FSCL_OK;
}

// stored permuted CLRs on the device: [n_pts][clr_cap] floats
static int ensure_pclr(Ctx *c, int n_pts, int clr_cap) {
  const size_t need = (size_t)std::max(n_pts, 1) * (size_t)std::max(clr_cap, 1);
  if (need > c->pclr_floats) {
    if (c->d_pclr) dev_free(c->d_pclr, c->stream);
    c->d_pclr = nullptr; c->pclr_floats = 0;
    CK(dev_malloc(&c->d_pclr, sizeof(float) * need));
    c->pclr_floats = need;
  }
  return FSCL_OK;
}

static int ensure_small(Ctx *c, int n) {
  if (n <= c->small_cap) return FSCL_OK;
  if (c->d_small) dev_free(c->d_small, c->stream);
  c->d_small = nullptr; c->small_cap = 0;
  CK(dev_malloc(&c->d_small, sizeof(int) * (5 * (size_t)n + 8)));
  c->small_cap = n;
  return FSCL_OK;
}

// permutations per rank and round at most: a round of few active cells is sized for 16384 searches, but 64
// permuted genomes per launch already fill the GPU, cost a quarter of the memory of 256 and waste fewer
// evaluations on cells that finish inside the round
constexpr int kMaxRoundPerRank = 64;

static int engine_resources(Ctx *c) {
  if (c->prep) return FSCL_OK;
  // highest priority: its few, short kernels must slip in between the CTAs of the running scan instead of
  // queueing behind that whole grid (equal priorities serialise prep(r) after scan(r-1): measured)
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CK(cudaStreamCreateWithPriority(&c->prep, cudaStreamNonBlocking, prio_hi));
  for (int k = 0; k < 2; k++) {
    CK(cudaEventCreateWithFlags(&c->ev_gen[k], cudaEventDisableTiming | cudaEventBlockingSync));
    CK(cudaEventCreateWithFlags(&c->ev_prep[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_scan[k], cudaEventDisableTiming));
  }
  for (int k = 0; k < 4; k++) {
    CK(cudaEventCreateWithFlags(&c->ev_round[k], cudaEventDisableTiming));
    CK(cudaEventCreate(&c->ev_k0[k]));
    CK(cudaEventCreate(&c->ev_k1[k]));
  }
  CK(cudaHostAlloc((void **)&c->h_pin, sizeof(int) * (4 + 2 * 3 * kPinSlots), cudaHostAllocDefault));
  CK(dev_malloc(&c->d_eng, sizeof(int) * (4 + 2 * 3 * kPinSlots)));
  CK(dev_malloc(&c->d_totals, 2 * sizeof(unsigned long long)));
  return FSCL_OK;
}

// Permuted genomes per rank and round that fit the memory budget of the two slot sets: 12 GB by default,
// FSCL_B200_SLOT_GB=<n> to change it (c5, 20 M sites = 400 MB per genome: 15 per set by default; a long test on such an
// input wants 48, so that the scan of a late round outlasts the swap phase of the next one -- one CTA per permutation,
// 19 ms at 20 M sites).  Every rank of a sharded run must use the same value: the round size follows from it.
static int slot_limit(const Ctx *c) {
  const size_t slot_bytes = (size_t)c->n * 20;        // tiles (key, pos, runs) + classes + scratch
  return (int)std::max<size_t>(1, std::min<size_t>(kPinSlots, c->slot_memory / 2 / slot_bytes));
}

// Block-list capacity per permutation on the device.  A permutation of c4 (3 M sites, -b 0.1 blocks per Mb) has about
// 3 000 blocks; a list that does not fit is flagged by gen_blocks_kernel and the round's lists are drawn on the host
// instead (same draws), so the capacity is a speed matter only.  (Round 2 first sized it n/4 + 1024 per permutation
// and grew the buffer whenever a round had more permutations than any before: 126 re-allocations of up to 770 MB in
// a c4 -p 10000 run, 5 s of its 21 s.)
static int block_list_stride(const Ctx *c) {
  if (c->block_cap > 0) return c->block_cap;                 // FSCL_B200_BLOCK_CAP (tests of the overflow path)
  return std::min(c->n / 4 + 1024, std::max(65536, c->n / 32));
}

static int ensure_block_lists(Ctx *c, int set, int slots) {
  SlotRef Q = slot_set(c, set);
  const size_t need = (size_t)slots * (size_t)block_list_stride(c);
  if (need <= Q.blocks_cap) return FSCL_OK;
  CK(cudaStreamSynchronize(c->prep));
  if (Q.blocks) dev_free(Q.blocks, c->prep);
  Q.blocks = nullptr; Q.blocks_cap = 0;
  CK(dev_malloc(&Q.blocks, need * sizeof(int4)));
  Q.blocks_cap = need;
  return FSCL_OK;
}

static int device_begin(Ctx *c) {
  PermState &S = c->perm;
  const int n = S.n_pts;
  int rc = ensure_pts(c, std::max(n, 1)); if (rc) return rc;
  rc = engine_resources(c); if (rc) return rc;
  S.clr_cap = std::min(std::max(S.o.n_permute + 1, 1), FSCL_NULL_SAVE);
  S.keep_clr = S.permute_clr != nullptr || S.o.save_clr != 0;
  if (S.keep_clr) { rc = ensure_pclr(c, n, S.clr_cap); if (rc) return rc; }
  std::vector<int> chr(n), sweep(n), pp(n), pn(n), pf(n), act;
  std::vector<double> clr(n);
  for (int i = 0; i < n; i++) {
    const fscl_pt_t &p = S.pts[i];
    chr[i] = p.chr; sweep[i] = p.sweep_pos; clr[i] = p.clr; pp[i] = p.permute_p; pn[i] = p.permute_n;
    pf[i] = p.permute_finished;
    if (!p.permute_finished) act.push_back(i);
  }
  const size_t b = sizeof(int) * (size_t)n;
  if (n > 0) {
    CK(cudaMemcpy(c->d_pt_chr, chr.data(), b, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_pt_sweep, sweep.data(), b, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_pp, pp.data(), b, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_pn, pn.data(), b, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_pf, pf.data(), b, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_pt_clr, clr.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    if (!act.empty()) CK(cudaMemcpy(c->d_active[0], act.data(), sizeof(int) * act.size(), cudaMemcpyHostToDevice));
  }
  S.cur = 0;
  S.n_active = (int)act.size();
  S.round = 0; S.collected = 0;
  const int counts[3] = {S.n_active, S.n_active, 0};
  CK(cudaMemcpy(c->d_eng, counts, sizeof counts, cudaMemcpyHostToDevice));
  CK(cudaMemset(c->d_totals, 0, 2 * sizeof(unsigned long long)));
  c->set_used[0] = c->set_used[1] = false;
  // reserve the two permutation slot sets once: rounds grow as cells are pruned, re-allocating on the way is slow
  int want = std::min(kMaxRoundPerRank, slot_limit(c));
  want = std::min(want, (S.o.n_permute + S.stride) / S.stride);
  if (S.o.round_perms > 0) want = std::min(std::min(slot_limit(c), kPinSlots), (S.o.round_perms + S.stride - 1) / S.stride);
  want = std::max(want, 1);
  rc = ensure_slots(c, want, 0); if (rc) return rc;
  rc = ensure_block_lists(c, 0, want); if (rc) return rc;
  if ((long)S.o.n_permute + 1 > (long)want * S.stride) {
    rc = ensure_slots(c, want, 1); if (rc) return rc;
    rc = ensure_block_lists(c, 1, want); if (rc) return rc;
  }
  return FSCL_OK;
}

// bookkeeping for the rounds below `upto`, which the host now knows to be complete
static int engine_collect(Ctx *c, long upto) {
  PermState &S = c->perm;
  for (; S.collected < upto; S.collected++) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev_k0[S.collected & 3], c->ev_k1[S.collected & 3]));
    c->stats.score_kernel_ms += ms;
    if (c->round_log) {
      // FSCL_B200_ROUND_LOG=1: per round, the scan kernel's time and the idle time of the scan stream before it
      float gap = 0;
      if (S.collected > 0 && cudaEventElapsedTime(&gap, c->ev_k1[(S.collected - 1) & 3], c->ev_k0[S.collected & 3]) != cudaSuccess) { cudaGetLastError(); gap = -1; }
      fprintf(stderr, "fscl round %ld R %d active<= %d scan_ms %.3f gap_ms %.3f\n", S.collected, c->log_R[S.collected & 3],
              c->log_na[S.collected & 3], ms, gap);
    }
  }
  return FSCL_OK;
}

static int device_round_score(Ctx *c, int *R_out, int *n_active_out) {
  PermState &S = c->perm;
  const fscl_permute_opts_t &o = S.o;
  *R_out = 0;
  // the active count after round r-2 (the initial one for the first two rounds)
  const long r = S.round;
  if (r >= 2) {
    CK(cudaEventSynchronize(c->ev_round[(r - 2) & 3]));
    S.n_active = c->h_pin[(r - 2) & 3];
    int rc = engine_collect(c, r - 1); if (rc) return rc;
  }
  if (n_active_out) *n_active_out = S.n_active;
  if (S.n_active == 0 || S.t_next > o.n_permute) return FSCL_OK;
  const size_t na = (size_t)S.n_active;
  const int lim = slot_limit(c);
  // enough permutations that EVERY rank launches about 16384 searches: one per rank while most cells are active (the
  // genomes of the next round are built under the scan of this one, so a short round costs nothing, and a cell that
  // finishes inside a round has been searched to its end for nothing: 6.4 % of the searches at 8 GPUs with rounds of 4
  // per rank), up to 64 once few cells are left
  int R = o.round_perms > 0 ? o.round_perms
                            : (int)std::min<size_t>((size_t)kMaxRoundPerRank * S.stride,
                                                    std::max<size_t>(1, (16384 + na - 1) / na) * (size_t)S.stride);
  R = (R + S.stride - 1) / S.stride * S.stride;
  R = std::min(R, lim * S.stride);
  R = std::min(R, o.n_permute - S.t_next + 1);
  std::vector<int> tl;
  for (int q = 0; q < R; q++)
    if (((S.t_next + q) % S.stride) == o.first_perm) tl.push_back(S.t_next + q);
  const int ns = (int)tl.size(), ns_max = (R + S.stride - 1) / S.stride;
  const int set = (int)(r & 1);
  int rc = ensure_slots(c, ns_max, set); if (rc) return rc;
  SlotRef Q = slot_set(c, set);
  const size_t chunk = (size_t)ns_max * na;
  {
    const size_t need = chunk * (size_t)c->world;
    if (need > c->matrix_cap) {
      CK(cudaStreamSynchronize(c->stream));
      if (c->d_matrix) dev_free(c->d_matrix, c->stream);
      c->d_matrix = nullptr; c->matrix_cap = 0;
      CK(dev_malloc(&c->d_matrix, need * sizeof(double)));
      c->matrix_cap = need;
    }
    if (R > c->scored_cap) {
      CK(cudaStreamSynchronize(c->stream));
      if (c->d_scored) dev_free(c->d_scored, c->stream);
      c->d_scored = nullptr; c->scored_cap = 0;
      const int cap = std::max(R, 256);
      CK(dev_malloc(&c->d_scored, sizeof(int) * (size_t)cap));
      CK(cudaMemset(c->d_scored, 0, sizeof(int) * (size_t)cap));
      c->scored_cap = cap;
    }
    if (chunk > c->tasks_cap) {
      CK(cudaStreamSynchronize(c->stream));
      rc = ensure_tasks(c, chunk); if (rc) return rc;
    }
  }
  int *d_cnt = c->d_eng + S.cur;                            // active count going into this round
  int *d_ntasks = c->d_eng + 2;
  int *d_tl = c->d_eng + 4 + set * 3 * kPinSlots, *d_flags = d_tl + kPinSlots, *d_d7 = d_flags + kPinSlots;
  int *h_flags = c->h_pin + 4 + set * 3 * kPinSlots, *h_d7 = h_flags + kPinSlots, *h_cnt = h_d7 + kPinSlots;
  if (ns > 0) {
    // ---- prep stream: the permuted genomes of this round ----
    if (c->set_used[set]) CK(cudaStreamWaitEvent(c->prep, c->ev_scan[set], 0));      // the scan of round r-2 read this set
    const int stride = block_list_stride(c);                  // block-list capacity per permutation
    rc = ensure_block_lists(c, set, ns_max); if (rc) return rc;      // sized at begin: no allocation between rounds
    std::vector<int> starts(ns);
    for (int q = 0; q < ns; q++) starts[q] = q * stride;
    CK(cudaMemcpyAsync(d_tl, tl.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, c->prep));
    CK(cudaMemcpyAsync(Q.blkoff, starts.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, c->prep));
    launch_gen_blocks(c->n, c->d_pos, c->d_chr_end, o.seed, d_tl, o.permute_nbp, o.scan_width_mb * 1e6, Q.blocks,
                      stride, Q.blkcnt, d_flags, d_d7, ns, c->prep, d_cnt);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * ns, cudaMemcpyDeviceToHost, c->prep));
    CK(cudaMemcpyAsync(h_d7, d_d7, sizeof(int) * ns, cudaMemcpyDeviceToHost, c->prep));
    CK(cudaMemcpyAsync(h_cnt, Q.blkcnt, sizeof(int) * ns, cudaMemcpyDeviceToHost, c->prep));
    CK(cudaEventRecord(c->ev_gen[set], c->prep));
    CK(cudaEventSynchronize(c->ev_gen[set]));                 // the draws only; the previous round's scan is still running
    bool redo = false;
    for (int q = 0; q < ns; q++) {
      if (h_flags[q] & 6) redo = true;                        // ambiguous truncation or overflow: host draws
      c->stats.d7_hits += h_d7[q];
      c->stats.n_blocks += h_cnt[q];
    }
    if (redo) {
      BlockGen gen{c->n, c->h_pos.data(), c->chr_end_index.data(), o.permute_nbp, o.scan_width_mb * 1e6};
      std::vector<int4> blocks; std::vector<int> off(1, 0), cnt(ns);
      int dd = 0;
      for (int q = 0; q < ns; q++) { S.src.set_perm((uint32_t)tl[q]); gen.generate(S.src, blocks, &dd); off.push_back((int)blocks.size()); }
      if (blocks.size() > Q.blocks_cap) {
        CK(cudaStreamSynchronize(c->prep));
        if (Q.blocks) dev_free(Q.blocks, c->prep);
        Q.blocks = nullptr; Q.blocks_cap = 0;
        CK(dev_malloc(&Q.blocks, (blocks.size() + 1024) * sizeof(int4)));
        Q.blocks_cap = blocks.size() + 1024;
      }
      for (int q = 0; q < ns; q++) cnt[q] = off[q + 1] - off[q];
      CK(cudaMemcpyAsync(Q.blocks, blocks.data(), blocks.size() * sizeof(int4), cudaMemcpyHostToDevice, c->prep));
      CK(cudaMemcpyAsync(Q.blkoff, off.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, c->prep));
      CK(cudaMemcpyAsync(Q.blkcnt, cnt.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, c->prep));
      CK(cudaStreamSynchronize(c->prep));                     // the vectors are locals
    }
    launch_permute(c->n, c->d_cls0, Q.pcls, Q.scratch, Q.blocks, Q.blkoff, Q.blkcnt, ns, c->prep, d_cnt);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
    launch_bucket(c->n, c->n_pad, c->P.n_cls, Q.pcls, (size_t)c->n, c->d_pos, c->d_null, Q.ts, ns, c->prep, d_cnt);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_prep[set], c->prep));
    c->stats.permute_launches++;
    // ---- scan stream ----
    CK(cudaStreamWaitEvent(c->stream, c->ev_prep[set], 0));
    launch_taskgen((int)na, ns, c->d_active[S.cur], d_cnt, c->d_pt_chr, c->d_pt_sweep, o.large_grid_sp, c->d_tasks, d_ntasks,
                   c->d_totals, c->stream);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_k0[r & 3], c->stream));
    launch_scan(c->P, c->d_tasks, nullptr, Q.ts, (int)((size_t)ns * na), scan_flags(c), c->d_terms,
                c->d_matrix + (size_t)c->rank * chunk, c->stream, d_ntasks, c->d_queue, c->n_sm);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_k1[r & 3], c->stream));
    CK(cudaEventRecord(c->ev_scan[set], c->stream));
    c->set_used[set] = true;
    c->stats.score_launches++;
  } else {
    CK(cudaEventRecord(c->ev_k0[r & 3], c->stream));
    CK(cudaEventRecord(c->ev_k1[r & 3], c->stream));
  }
  S.R = R; S.na_round = (int)na; S.ns_round = ns; S.chunk = chunk;
  c->log_R[r & 3] = R; c->log_na[r & 3] = (int)na;
  *R_out = R;
  return FSCL_OK;
}

static int device_round_commit(Ctx *c) {
  PermState &S = c->perm;
  const fscl_permute_opts_t &o = S.o;
  const long r = S.round;
  if (c->world > 1) {
    int rc = nccl_all_gather(c, c->d_matrix, S.chunk * sizeof(double)); if (rc) return rc;
  }
  int *d_cnt = c->d_eng + S.cur, *d_cnt_next = c->d_eng + (S.cur ^ 1);
  launch_replay(S.na_round, d_cnt, S.R, S.t_next, S.stride, S.chunk, c->d_active[S.cur], c->d_matrix, c->d_pt_clr, c->d_pp,
                c->d_pn, c->d_pf, S.keep_clr ? c->d_pclr : nullptr, S.clr_cap, o.seed, o.prune, c->d_scored, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  launch_compact(S.na_round, d_cnt, c->d_active[S.cur], c->d_pf, c->d_active[S.cur ^ 1], d_cnt_next, c->d_scored, S.R,
                 c->d_totals + 1, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c->h_pin + (r & 3), d_cnt_next, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaEventRecord(c->ev_round[r & 3], c->stream));
  S.cur ^= 1;
  c->stats.n_rounds++;
  S.t_next += S.R;
  S.R = 0;
  S.round = r + 1;
  return FSCL_OK;
}

// wait for everything enqueued, bring the host-side view (active count, timings, totals) up to date
static int device_drain(Ctx *c) {
  PermState &S = c->perm;
  CK(cudaStreamSynchronize(c->prep));
  CK(cudaStreamSynchronize(c->stream));
  { int rc = engine_collect(c, S.round); if (rc) return rc; }
  if (S.round > 0) S.n_active = c->h_pin[(S.round - 1) & 3];
  S.round = 0; S.collected = 0;                    // the ring starts over: the next round waits for nothing
  const int counts[2] = {S.n_active, S.n_active};
  CK(cudaMemcpy(c->d_eng, counts, sizeof counts, cudaMemcpyHostToDevice));
  unsigned long long tot[2] = {0, 0};
  CK(cudaMemcpy(tot, c->d_totals, sizeof tot, cudaMemcpyDeviceToHost));
  CK(cudaMemset(c->d_totals, 0, sizeof tot));
  c->stats.n_tasks += (int64_t)tot[0];
  c->stats.n_grid_evals += (int64_t)tot[0];
  c->stats.n_scored += (int64_t)tot[1];
  return FSCL_OK;
}

// stored CLRs: device table <-> host, one kernel + one copy.  pn = permute_n per cell.
// dense != NULL: rows of FSCL_NULL_SAVE floats; else packed (cell i at off[i]).
static int clr_offsets(const std::vector<int> &pn, int clr_cap, std::vector<long long> &off) {
  off.resize(pn.size() + 1);
  off[0] = 0;
  for (size_t i = 0; i < pn.size(); i++) off[i + 1] = off[i] + std::min(std::max(pn[i], 0), clr_cap);
  return FSCL_OK;
}

static int ensure_pack(Ctx *c, int n_pts, size_t total) {
  if (n_pts + 1 > c->off_cap) {
    if (c->d_off) dev_free(c->d_off, c->stream);
    c->d_off = nullptr; c->off_cap = 0;
    CK(dev_malloc(&c->d_off, sizeof(long long) * (size_t)(n_pts + 1)));
    c->off_cap = n_pts + 1;
  }
  if (total > c->pack_cap) {
    if (c->d_pack) dev_free(c->d_pack, c->stream);
    c->d_pack = nullptr; c->pack_cap = 0;
    CK(dev_malloc(&c->d_pack, sizeof(float) * std::max<size_t>(total, 1)));
    c->pack_cap = std::max<size_t>(total, 1);
  }
  return FSCL_OK;
}

static int clr_to_host(Ctx *c, const std::vector<int> &pn, float *dense, float *packed, const std::vector<long long> &off) {
  PermState &S = c->perm;
  const int n = (int)pn.size();
  const size_t total = (size_t)off[n];
  if (n == 0 || total == 0) return FSCL_OK;
  int rc = ensure_pack(c, n, total); if (rc) return rc;
  CK(cudaMemcpyAsync(c->d_off, off.data(), sizeof(long long) * (size_t)(n + 1), cudaMemcpyHostToDevice, c->stream));
  launch_pack_clr(n, c->d_pn, c->d_off, S.clr_cap, c->d_pclr, c->d_pack, 0, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  std::vector<float> tmp;
  float *dst = packed;
  if (!dst) { tmp.resize(total); dst = tmp.data(); }
  CK(cudaMemcpyAsync(dst, c->d_pack, sizeof(float) * total, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (dense)
    for (int i = 0; i < n; i++)
      if (off[i + 1] > off[i]) memcpy(dense + (size_t)i * FSCL_NULL_SAVE, dst + off[i], sizeof(float) * (size_t)(off[i + 1] - off[i]));
  return FSCL_OK;
}

static int clr_to_device(Ctx *c, const std::vector<int> &pn, const float *dense) {
  PermState &S = c->perm;
  const int n = (int)pn.size();
  std::vector<long long> off;
  clr_offsets(pn, S.clr_cap, off);
  const size_t total = (size_t)off[n];
  if (n == 0 || total == 0) return FSCL_OK;
  int rc = ensure_pack(c, n, total); if (rc) return rc;
  std::vector<float> tmp(total);
  for (int i = 0; i < n; i++)
    if (off[i + 1] > off[i]) memcpy(tmp.data() + off[i], dense + (size_t)i * FSCL_NULL_SAVE, sizeof(float) * (size_t)(off[i + 1] - off[i]));
  CK(cudaMemcpyAsync(c->d_off, off.data(), sizeof(long long) * (size_t)(n + 1), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_pack, tmp.data(), sizeof(float) * total, cudaMemcpyHostToDevice, c->stream));
  launch_pack_clr(n, c->d_pn, c->d_off, S.clr_cap, c->d_pclr, c->d_pack, 1, c->stream);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return FSCL_OK;
}

static int device_end(Ctx *c) {
  PermState &S = c->perm;
  const int n = S.n_pts;
  int rc = device_drain(c); if (rc) return rc;
  if (n == 0) return FSCL_OK;
  std::vector<int> pp(n), pn(n), pf(n);
  const size_t b = sizeof(int) * (size_t)n;
  CK(cudaMemcpy(pp.data(), c->d_pp, b, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pn.data(), c->d_pn, b, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pf.data(), c->d_pf, b, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) { S.pts[i].permute_p = pp[i]; S.pts[i].permute_n = pn[i]; S.pts[i].permute_finished = pf[i]; }
  if (S.permute_clr) {
    // only the entries written are meaningful (first permute_n, at most clr_cap)
    std::vector<long long> off;
    clr_offsets(pn, S.clr_cap, off);
    rc = clr_to_host(c, pn, S.permute_clr, nullptr, off); if (rc) return rc;
  }
  return FSCL_OK;
}

// Permutation test as rounds (scan_permute_thread scan-chromosome.c:412-546).
// A round covers permutations [t_next, t_next+R).  Each rank scores the
// permutations t with t % stride == first into rows of a device matrix
// clr[R][n_active] (other rows stay 0); after the caller has summed the matrix
// over ranks (NCCL allreduce; nothing to do on one GPU) commit() replays the
// reference's sequential counting / pruning rule over all R rows, so counts do
// not depend on the number of GPUs.
static int perm_begin(Ctx *c, fscl_pt_t *pts, int n_pts, const fscl_permute_opts_t *o, float *permute_clr) {
  if (!c || !o || (n_pts > 0 && !pts) || n_pts < 0 || o->large_grid_sp <= 0 || o->permute_nbp <= 0)
    return fail(FSCL_ERR_ARG, "bad argument");
  PermState &S = c->perm;
  S = PermState();
  S.o = *o;
  S.stride = o->perm_stride > 0 ? o->perm_stride : 1;
  if (o->first_perm < 0 || o->first_perm >= S.stride) return fail(FSCL_ERR_ARG, "first_perm out of range");
  S.src.init(o->rng_kind, o->seed);
  if (S.src.sequential() && S.stride != 1) return fail(FSCL_ERR_ARG, "sequential RNG streams cannot be sharded");
  if (o->rng_kind == FSCL_RNG_GLIBC) (void)S.src.block_draw();     // usleep burn, scan-chromosome.c:440
  if (o->rng_kind == FSCL_RNG_LIBC) (void)rand();
  S.pts = pts; S.n_pts = n_pts; S.permute_clr = permute_clr;
  S.active.resize(n_pts);
  for (int i = 0; i < n_pts; i++) S.active[i] = i;
  S.t_next = 0;
  S.open = true;
  S.rounds_at_begin = c->stats.n_rounds;
  // counter-based stream: the loop body lives on the GPU (FSCL_B200_HOST_ROUNDS=1 keeps the host path)
  const char *hr = getenv("FSCL_B200_HOST_ROUNDS");
  S.on_device = !S.src.sequential() && !(hr && atoi(hr) != 0);
  if (c->comm) {
    // sharded over the communicator: rank g scores the permutations t with t % world == g
    if (S.src.sequential()) return fail(FSCL_ERR_ARG, "sequential RNG streams cannot be sharded");
    S.stride = c->world; S.o.first_perm = c->rank; S.o.perm_stride = c->world;
    if (!S.on_device) return fail(FSCL_ERR_ARG, "FSCL_B200_HOST_ROUNDS cannot be combined with a communicator");
  } else if (S.stride > 1) {
    S.on_device = false;      // the caller sums an external matrix over its own ranks: synchronous host rounds
  }
  if (S.on_device) {
    CK(cudaSetDevice(c->device));
    return device_begin(c);
  }
  return FSCL_OK;
}

// returns FSCL_OK and *R_out == 0 when the test is finished
static int perm_round_score(Ctx *c, double *d_matrix, int64_t cap, int *R_out, int *n_active_out) {
  PermState &S = c->perm;
  if (!S.open) return fail(FSCL_ERR_ARG, "no permutation test in progress");
  CK(cudaSetDevice(c->device));
  if (S.on_device && d_matrix) {
    // an external matrix means the caller runs the collective itself: that needs the synchronous host rounds
    if (S.round != 0 || c->stats.n_rounds != S.rounds_at_begin || c->comm)
      return fail(FSCL_ERR_ARG, "an external CLR matrix must be used from the first round on (and without a communicator)");
    S.on_device = false;
  }
  if (S.on_device) return device_round_score(c, R_out, n_active_out);
  const fscl_permute_opts_t &o = S.o;
  const int G = o.large_grid_sp;
  fscl_pt_t *pts = S.pts;
  // compact (scan-chromosome.c:446-457)
  size_t k = 0;
  for (size_t a = 0; a < S.active.size(); a++)
    if (!pts[S.active[a]].permute_finished) S.active[k++] = S.active[a];
  S.active.resize(k);
  *R_out = 0;
  if (n_active_out) *n_active_out = (int)S.active.size();
  if (S.active.empty() || S.t_next > o.n_permute) return FSCL_OK;        // D5: 0..n_permute inclusive
  const size_t na = S.active.size();
  int R;
  if (S.src.sequential()) {
    // block draws of permutation t+1 follow the pruning draws of t in the one
    // stream; a pruning draw needs permute_p >= 20, so 20 - max_p permutations
    // can be generated ahead without reordering the stream
    int maxp = 0;
    for (int s : S.active) maxp = std::max(maxp, pts[s].permute_p);
    R = o.prune ? std::max(1, FSCL_PRUNE_MIN_P - maxp) : 64;
  } else {
    R = o.round_perms > 0 ? o.round_perms
                          : (int)std::min<size_t>(256, std::max<size_t>(S.stride, (16384 + na - 1) / na));
    R = (R + S.stride - 1) / S.stride * S.stride;
  }
  const size_t slot_bytes = (size_t)c->n * 8;
  const int mem_cap = (int)std::max<size_t>(1, ((size_t)8 << 30) / slot_bytes);
  R = std::min(R, mem_cap * S.stride);
  R = std::min(R, o.n_permute - S.t_next + 1);
  if (d_matrix && (int64_t)R * (int64_t)na > cap) R = (int)std::max<int64_t>(1, cap / (int64_t)na);
  if (d_matrix && (int64_t)na > cap) return fail(FSCL_ERR_ARG, "clr matrix too small for one permutation");
  S.my_rows.clear();
  S.blocks.clear(); S.off.assign(1, 0);
  int d7 = 0;
  BlockGen gen{c->n, c->h_pos.data(), c->chr_end_index.data(), o.permute_nbp, o.scan_width_mb * 1e6};
  for (int r = 0; r < R; r++) {
    int t = S.t_next + r;
    bool mine = (t % S.stride) == o.first_perm;
    if (S.src.sequential()) gen.generate(S.src, S.blocks, &d7);
    else if (mine) { S.src.set_perm((uint32_t)t); gen.generate(S.src, S.blocks, &d7); }
    if (mine) { S.my_rows.push_back(r); S.off.push_back((int)S.blocks.size()); }
  }
  c->stats.d7_hits += d7;
  const int ns = (int)S.my_rows.size();
  // matrix
  double *dm = d_matrix;
  if (!dm) {
    size_t need = (size_t)R * na;
    if (need > c->matrix_cap) {
      if (c->d_matrix) dev_free(c->d_matrix, c->stream);
      c->d_matrix = nullptr; c->matrix_cap = 0;
      CK(dev_malloc(&c->d_matrix, need * sizeof(double)));
      c->matrix_cap = need;
    }
    dm = c->d_matrix;
  }
  S.d_matrix = dm;
  if (ns < R) CK(cudaMemsetAsync(dm, 0, (size_t)R * na * sizeof(double), c->stream));
  if (ns > 0) {
    int rc = ensure_slots(c, ns); if (rc) return rc;
    rc = upload_blocks(c, S.blocks, S.off); if (rc) return rc;
    rc = apply_blocks(c, ns); if (rc) return rc;
    c->stats.n_blocks += (int64_t)S.blocks.size();
    c->h_tasks.resize((size_t)ns * na);
    for (int q = 0; q < ns; q++)
      for (size_t a = 0; a < na; a++) {
        const fscl_pt_t &p = pts[S.active[a]];
        int start = p.sweep_pos - (p.sweep_pos % G);                    // D4
        c->h_tasks[(size_t)q * na + a] = Task{q, p.chr, start, start + G, (int)((size_t)S.my_rows[q] * na + a)};
      }
    rc = run_tasks(c, true, 0, false, dm); if (rc) return rc;
    c->stats.n_grid_evals += (int64_t)ns * (int64_t)na;
  } else {
    CK(cudaStreamSynchronize(c->stream));
  }
  S.R = R;
  *R_out = R;
  return FSCL_OK;
}

static int perm_round_commit(Ctx *c, const double *d_matrix) {
  PermState &S = c->perm;
  if (!S.open || S.R <= 0) return fail(FSCL_ERR_ARG, "no scored round to commit");
  CK(cudaSetDevice(c->device));
  if (S.on_device) return device_round_commit(c);
  const fscl_permute_opts_t &o = S.o;
  const size_t na = S.active.size();
  const double *dm = d_matrix ? d_matrix : S.d_matrix;
  S.h_matrix.resize((size_t)S.R * na);
  CK(cudaMemcpyAsync(S.h_matrix.data(), dm, S.h_matrix.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->stats.n_scored += replay_round(S.pts, S.active.data(), na, S.h_matrix.data(), S.R, S.t_next, S.src, o.prune,
                                    S.permute_clr);
  c->stats.n_rounds++;
  S.t_next += S.R;
  S.R = 0;
  return FSCL_OK;
}

extern "C" int fscl_perm_begin(void *ctx, fscl_pt_t *pts, int n_pts, const fscl_permute_opts_t *o,
                               float *permute_clr) {
  return perm_begin((Ctx *)ctx, pts, n_pts, o, permute_clr);
}
extern "C" int fscl_perm_round_score(void *ctx, double *d_matrix, int64_t cap, int *R_out, int *n_active_out) {
  if (!ctx || !R_out) return fail(FSCL_ERR_ARG, "bad argument");
  return perm_round_score((Ctx *)ctx, d_matrix, cap, R_out, n_active_out);
}
extern "C" int fscl_perm_round_commit(void *ctx, const double *d_matrix) {
  if (!ctx) return fail(FSCL_ERR_ARG, "bad argument");
  return perm_round_commit((Ctx *)ctx, d_matrix);
}
extern "C" int fscl_perm_end(void *ctx, fscl_stats_t *stats_out) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(FSCL_ERR_ARG, "bad argument");
  if (c->perm.open && c->perm.on_device) {
    CK(cudaSetDevice(c->device));
    int rc = device_end(c);
    if (rc) return rc;
  }
  c->perm.open = false;
  if (stats_out) *stats_out = c->stats;
  return FSCL_OK;
}

// Bring the caller's pts / permute_clr up to date between rounds without ending the test
// (what the reference's SIGINT handler reads, scan-chromosome.c:557-569; also the checkpoint).
extern "C" int fscl_perm_snapshot(void *ctx, int32_t *t_next_out) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !c->perm.open) return fail(FSCL_ERR_ARG, "no permutation test in progress");
  if (c->perm.on_device) {
    CK(cudaSetDevice(c->device));
    int rc = device_end(c);
    if (rc) return rc;
  }
  if (t_next_out) *t_next_out = c->perm.t_next;
  return FSCL_OK;
}

// Continue a test at permutation t_next (right after fscl_perm_begin): the counters and finished
// flags were taken from pts by fscl_perm_begin; the stored permuted CLRs are re-loaded here.
// Counter-based stream only -- a sequential rand() stream cannot be re-entered in the middle.
extern "C" int fscl_perm_seek(void *ctx, int32_t t_next) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !c->perm.open) return fail(FSCL_ERR_ARG, "no permutation test in progress");
  PermState &S = c->perm;
  if (S.src.sequential()) return fail(FSCL_ERR_ARG, "only the counter-based stream can resume");
  if (t_next < 0 || S.t_next != 0) return fail(FSCL_ERR_ARG, "seek is only valid right after fscl_perm_begin");
  S.t_next = t_next;
  if (S.on_device && S.permute_clr) {
    CK(cudaSetDevice(c->device));
    std::vector<int> pn(S.n_pts);
    for (int i = 0; i < S.n_pts; i++) pn[i] = S.pts[i].permute_n;
    return clr_to_device(c, pn, S.permute_clr);
  }
  return FSCL_OK;
}

extern "C" int fscl_scan_permute(void *ctx, fscl_pt_t *pts, int n_pts, const fscl_permute_opts_t *o,
                                 float *permute_clr, fscl_stats_t *stats_out) {
  Ctx *c = (Ctx *)ctx;
  if (o && o->perm_stride > 1 && !(c && c->comm))
    return fail(FSCL_ERR_ARG, "sharded runs need a communicator (fscl_comm_init_*) or the round API with an external matrix");
  int rc = perm_begin(c, pts, n_pts, o, permute_clr);
  if (rc) return rc;
  for (;;) {
    int R = 0;
    rc = perm_round_score(c, nullptr, 0, &R, nullptr);
    if (rc) return rc;
    if (R == 0) break;
    rc = perm_round_commit(c, nullptr);
    if (rc) return rc;
  }
  return fscl_perm_end(c, stats_out);
}

// --------------------------------------------------------------------------- multi-GPU (NCCL)
// The reference's fscl-master.c:1-18 is an unfinished master / slave split; here the permutation
// replicates shard over the GPUs of one box: rank g scores the permutations
// t = g (mod world) of every round into its own rows of the round's CLR matrix, ONE ncclAllGather per
// round (on the library's stream, NVLink / NVSwitch) makes the matrix complete on every rank, and
// every rank replays the same sequential counting / pruning rule -- counts are identical on all ranks
// and independent of the number of GPUs.  libnccl.so.2 is resolved at run time (dlopen), so the
// library loads and runs on one GPU without NCCL; in a process that already holds NCCL (torch) the
// loader hands back that copy.
extern "C" int fscl_comm_unique_id(char *id128) {
  if (!id128) return fail(FSCL_ERR_ARG, "bad argument");
  int rc = nccl_load(); if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return FSCL_OK;
}

extern "C" int fscl_comm_init_rank(void *ctx, int world, int rank, const char *id128) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(FSCL_ERR_ARG, "bad argument");
  if (c->comm) return fail(FSCL_ERR_ARG, "context already has a communicator");
  int rc = nccl_load(); if (rc) return rc;
  CK(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NK(g_nccl.CommInitRank(&c->comm, world, id, rank));
  c->world = world; c->rank = rank;
  return FSCL_OK;
}

// one process, n contexts on n different devices
extern "C" int fscl_comm_init_all(void **ctxs, int n) {
  if (!ctxs || n < 1 || n > kMaxDevices) return fail(FSCL_ERR_ARG, "bad argument");
  int rc = nccl_load(); if (rc) return rc;
  std::vector<int> devs(n);
  std::vector<ncclComm_t> comms(n);
  for (int i = 0; i < n; i++) {
    Ctx *c = (Ctx *)ctxs[i];
    if (!c || c->comm) return fail(FSCL_ERR_ARG, "bad context in the list");
    devs[i] = c->device;
    for (int j = 0; j < i; j++) if (devs[j] == devs[i]) return fail(FSCL_ERR_ARG, "two contexts on one device");
  }
  NK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; i++) { Ctx *c = (Ctx *)ctxs[i]; c->comm = comms[i]; c->world = n; c->rank = i; }
  return FSCL_OK;
}

extern "C" int fscl_comm_destroy(void *ctx) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(FSCL_ERR_ARG, "bad argument");
  if (c->comm) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy(c->comm);
  }
  c->comm = nullptr; c->world = 1; c->rank = 0;
  return FSCL_OK;
}

// hand the communicator of one context to another context on the same device (a fresh problem on the
// same rank does not need a fresh NCCL bootstrap); src is left without one
extern "C" int fscl_comm_move(void *dst_ctx, void *src_ctx) {
  Ctx *d = (Ctx *)dst_ctx, *s = (Ctx *)src_ctx;
  if (!d || !s || d == s) return fail(FSCL_ERR_ARG, "bad argument");
  if (d->comm) return fail(FSCL_ERR_ARG, "destination already has a communicator");
  if (d->device != s->device) return fail(FSCL_ERR_ARG, "contexts are on different devices");
  CK(cudaSetDevice(s->device));
  CK(cudaStreamSynchronize(s->stream));
  d->comm = s->comm; d->world = s->world; d->rank = s->rank;
  s->comm = nullptr; s->world = 1; s->rank = 0;
  return FSCL_OK;
}

extern "C" int fscl_comm_info(void *ctx, int *world, int *rank) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(FSCL_ERR_ARG, "bad argument");
  if (world) *world = c->world;
  if (rank) *rank = c->rank;
  return FSCL_OK;
}

// ---- stored permuted CLRs and the projection, device-resident ----------------------------------
// After fscl_perm_end / fscl_perm_snapshot of a test begun with permute_clr != NULL or opts.save_clr:
// offsets[n_pts + 1] (cell i holds offsets[i+1] - offsets[i] = min(permute_n, FSCL_NULL_SAVE, n_permute + 1)
// values); packed == NULL only fills the offsets (size the buffer with offsets[n_pts] floats first).
extern "C" int fscl_perm_fetch_clr(void *ctx, int64_t *offsets, float *packed, int64_t cap) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !offsets) return fail(FSCL_ERR_ARG, "bad argument");
  PermState &S = c->perm;
  if (!S.on_device || !S.keep_clr || !c->d_pclr) return fail(FSCL_ERR_ARG, "no permuted CLRs were kept on the device for this test");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const int n = S.n_pts;
  std::vector<int> pn(n);
  if (n > 0) CK(cudaMemcpy(pn.data(), c->d_pn, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
  std::vector<long long> off;
  clr_offsets(pn, S.clr_cap, off);
  for (int i = 0; i <= n; i++) offsets[i] = off[i];
  if (!packed) return FSCL_OK;
  if (cap < off[n]) return fail(FSCL_ERR_ARG, "packed buffer too small");
  return clr_to_host(c, pn, nullptr, packed, off);
}

// p-value projection on the device from the CLRs kept there (definition: DESIGN.md 3.6; the reference
// describes it at README.md:11 and never implemented it)
extern "C" int fscl_perm_project(void *ctx, double df, double *lambda_out, double *p_out) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !p_out || !(df > 0)) return fail(FSCL_ERR_ARG, "bad argument");
  PermState &S = c->perm;
  if (!S.on_device || !S.keep_clr || !c->d_pclr) return fail(FSCL_ERR_ARG, "no permuted CLRs were kept on the device for this test");
  CK(cudaSetDevice(c->device));
  const int n = S.n_pts;
  if (n == 0) return FSCL_OK;
  double *d_out = nullptr;
  CK(dev_malloc(&d_out, sizeof(double) * 2 * (size_t)n));
  launch_project(n, c->d_pn, c->d_pt_clr, c->d_pclr, S.clr_cap, df, d_out, d_out + n, c->stream);
  c->stats.kernel_launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && lambda_out) e = cudaMemcpyAsync(lambda_out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p_out, d_out + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(d_out, c->stream);
  if (e != cudaSuccess) return fail(FSCL_ERR_CUDA, std::string("projection: ") + cudaGetErrorString(e));
  return FSCL_OK;
}

// Block list of permutation t drawn ON THE DEVICE (gen_blocks_kernel); test hook.
// flags: 1 D7, 2 ambiguous truncation, 4 overflow.
extern "C" int fscl_device_gen_blocks(void *ctx, const fscl_permute_opts_t *o, int t, int32_t *blocks_out,
                                      int max_blocks, int32_t *flags_out) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !o || t < 0 || o->rng_kind != FSCL_RNG_PHILOX) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  if (cudaSetDevice(c->device) != cudaSuccess) { fail(FSCL_ERR_CUDA, "cudaSetDevice"); return -1; }
  if (ensure_slots(c, 1) || ensure_small(c, 1)) return -1;
  const int stride = c->n / 4 + 1024;
  if ((size_t)stride > c->blocks_cap) {
    if (c->d_blocks) dev_free(c->d_blocks, c->stream);
    c->d_blocks = nullptr; c->blocks_cap = 0;
    if (dev_malloc(&c->d_blocks, (size_t)stride * sizeof(int4)) != cudaSuccess) { fail(FSCL_ERR_CUDA, "cudaMalloc"); return -1; }
    c->blocks_cap = stride;
  }
  const int sc = c->small_cap;
  int *d_tl = c->d_small, *d_flags = c->d_small + 2 * sc, *d_d7 = c->d_small + 3 * sc;
  auto bad = [&](cudaError_t e, const char *what) {
    if (e == cudaSuccess) return false;
    fail(FSCL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return true;
  };
  if (bad(cudaMemcpy(d_tl, &t, sizeof(int), cudaMemcpyHostToDevice), "gen_blocks: upload")) return -1;
  launch_gen_blocks(c->n, c->d_pos, c->d_chr_end, o->seed, d_tl, o->permute_nbp, o->scan_width_mb * 1e6, c->d_blocks,
                    stride, c->d_blkcnt, d_flags, d_d7, 1, c->stream);
  c->stats.kernel_launches++;
  int cnt = 0, flags = 0;
  if (bad(cudaGetLastError(), "gen_blocks_kernel launch")) return -1;
  if (bad(cudaStreamSynchronize(c->stream), "gen_blocks_kernel")) return -1;
  if (bad(cudaMemcpy(&cnt, c->d_blkcnt, sizeof(int), cudaMemcpyDeviceToHost), "gen_blocks: count")) return -1;
  if (bad(cudaMemcpy(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost), "gen_blocks: flags")) return -1;
  std::vector<int4> b(std::max(0, std::min(cnt, stride)));
  if (!b.empty() && bad(cudaMemcpy(b.data(), c->d_blocks, sizeof(int4) * b.size(), cudaMemcpyDeviceToHost), "gen_blocks: list")) return -1;
  for (size_t q = 0; q < b.size() && blocks_out && (int)q < max_blocks; q++) {
    blocks_out[3 * q] = b[q].x; blocks_out[3 * q + 1] = b[q].y; blocks_out[3 * q + 2] = b[q].z;
  }
  if (flags_out) *flags_out = flags;
  return cnt;
}

// --------------------------------------------------------------------------- host-only entry points
// These run without a device (tests exercise them on CPU-only machines); they are
// the same code the GPU path uses for its host side.
// The range-maximum table of the candidate bounds, as fscl_create builds it (host only; test hook):
// coef = [n_cls][n_knots][4], null_tab = [n_cls]; out = [levels][n_knots+2][n_cls], *levels_out = levels.
extern "C" int fscl_host_piece_maxima(int n_cls, int n_knots, double step, const double *coef, const double *null_tab,
                                      double *out, int64_t cap, int32_t *levels_out) {
  if (!coef || !null_tab || !out || !levels_out || n_cls <= 0 || n_knots <= 0) return fail(FSCL_ERR_ARG, "bad argument");
  const int K = n_knots;
  std::vector<double> coefC((size_t)(K + 2) * n_cls * 4), ptab;
  for (int cl = 0; cl < n_cls; cl++)
    for (int k = 0; k < K + 2; k++) {
      const double *q = coef + ((size_t)cl * K + std::min(k, K - 1)) * 4;
      double *d = coefC.data() + ((size_t)cl * (K + 2) + k) * 4;
      d[0] = q[0]; d[1] = q[1]; d[2] = q[2]; d[3] = q[3] - null_tab[cl];
    }
  build_piece_maxima(coefC.data(), n_cls, K, step, ptab);
  if ((int64_t)ptab.size() > cap) return fail(FSCL_ERR_ARG, "output too small");
  memcpy(out, ptab.data(), sizeof(double) * ptab.size());
  *levels_out = (int32_t)(ptab.size() / ((size_t)(K + 2) * n_cls));
  return FSCL_OK;
}

extern "C" int fscl_host_alpha_tables(const double *log_table, double *la, int32_t *dmax, int32_t *n) {
  if (!la || !dmax || !n) return fail(FSCL_ERR_ARG, "bad argument");
  std::vector<double> T(65536);
  if (log_table) memcpy(T.data(), log_table, sizeof(double) * 65536);
  else { T[0] = 0.0; for (int i = 1; i <= 0xFFFF; i++) T[i] = std::log((double)i); }
  AlphaTables A;
  build_alpha_tables(T.data(), A);
  for (int r = 0; r < kRows; r++) {
    n[r] = A.n[r];
    for (int j = 0; j < kMaxFine; j++) { la[r * kMaxFine + j] = A.la[r][j]; dmax[r * kMaxFine + j] = A.dmax[r][j]; }
  }
  return FSCL_OK;
}

extern "C" int fscl_host_enum_cells(int n_chr, const int32_t *chr_startpos, const int32_t *chr_bplen, int G,
                                    int32_t *cell_chr, int32_t *cell_start, int32_t *cell_end, int cap) {
  if (n_chr <= 0 || !chr_startpos || !chr_bplen || G <= 0) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  std::vector<Task> cells;
  int n = enum_cells_raw(n_chr, chr_startpos, chr_bplen, G, &cells);
  for (int i = 0; i < n && i < cap; i++) {
    if (cell_chr) cell_chr[i] = cells[i].chr;
    if (cell_start) cell_start[i] = cells[i].start_pos;
    if (cell_end) cell_end[i] = cells[i].end_pos;
  }
  return n;
}

// draw phase of snp_block_permute for permutation t: (i, j, len) triples
extern "C" int fscl_host_gen_blocks(int n_snps, const int32_t *chr, const int32_t *pos, const fscl_permute_opts_t *o,
                                    int t, int32_t *blocks_out, int max_blocks, int32_t *d7_out) {
  if (n_snps <= 0 || !chr || !pos || !o || t < 0 || o->permute_nbp <= 0) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  std::vector<int> cend(n_snps);
  for (int i = n_snps - 1; i >= 0; i--) cend[i] = (i == n_snps - 1 || chr[i + 1] != chr[i]) ? i + 1 : cend[i + 1];
  DrawSource src;
  src.init(o->rng_kind, o->seed);
  BlockGen gen{n_snps, pos, cend.data(), o->permute_nbp, o->scan_width_mb * 1e6};
  std::vector<int4> blocks;
  int d7 = 0;
  if (src.sequential()) {
    if (o->rng_kind != FSCL_RNG_LIBC) (void)src.block_draw();   // usleep burn, scan-chromosome.c:440
    for (int q = 0; q <= t; q++) { blocks.clear(); gen.generate(src, blocks, &d7); }
  } else {
    src.set_perm((uint32_t)t);
    gen.generate(src, blocks, &d7);
  }
  for (size_t b = 0; b < blocks.size() && blocks_out && (int)b < max_blocks; b++) {
    blocks_out[3 * b] = blocks[b].x; blocks_out[3 * b + 1] = blocks[b].y; blocks_out[3 * b + 2] = blocks[b].z;
  }
  if (d7_out) *d7_out = d7;
  return (int)blocks.size();
}

// counting + pruning of one round of R permutations starting at t_first
// (Philox streams only: the pruning uniform is addressed by (t, scan point))
extern "C" int fscl_host_replay_round(fscl_pt_t *pts, const int32_t *active, int n_active, const double *matrix,
                                      int R, int t_first, const fscl_permute_opts_t *o, float *permute_clr) {
  if (!pts || !active || n_active < 0 || !matrix || R < 0 || !o) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  if (o->rng_kind != FSCL_RNG_PHILOX) { fail(FSCL_ERR_ARG, "replay needs the counter-based stream"); return -1; }
  DrawSource src;
  src.init(o->rng_kind, o->seed);
  return replay_round(pts, active, (size_t)n_active, matrix, R, t_first, src, o->prune, permute_clr);
}

// The same for a round whose rows arrive as the library's all-gather delivers them: rank g's rows [q][n_active] at
// gathered + g * chunk (host statement of replay_kernel's addressing; the 2-rank gloo test drives it)
extern "C" int fscl_host_replay_gathered(fscl_pt_t *pts, const int32_t *active, int n_active, const double *gathered,
                                         int world, int64_t chunk, int R, int t_first, const fscl_permute_opts_t *o,
                                         float *permute_clr) {
  if (!pts || !active || n_active < 0 || !gathered || R < 0 || !o || world < 1 || chunk < 0) { fail(FSCL_ERR_ARG, "bad argument"); return -1; }
  if (o->rng_kind != FSCL_RNG_PHILOX) { fail(FSCL_ERR_ARG, "replay needs the counter-based stream"); return -1; }
  std::vector<double> dense((size_t)R * (size_t)n_active);
  for (int r = 0; r < R; r++) {
    int g, q;
    round_row_owner(t_first, world, r, g, q);
    if ((int64_t)(q + 1) * n_active > chunk) { fail(FSCL_ERR_ARG, "chunk too small for the round"); return -1; }
    memcpy(dense.data() + (size_t)r * n_active, gathered + (size_t)g * (size_t)chunk + (size_t)q * n_active, sizeof(double) * (size_t)n_active);
  }
  DrawSource src;
  src.init(o->rng_kind, o->seed);
  return replay_round(pts, active, (size_t)n_active, dense.data(), R, t_first, src, o->prune, permute_clr);
}

// 31-bit draws of the two generators (known-answer tests)
extern "C" int fscl_host_rng_draws(int rng_kind, uint64_t seed, uint32_t t, uint32_t stream, int n, int32_t *out) {
  if (!out || n < 0) return fail(FSCL_ERR_ARG, "bad argument");
  if (rng_kind == FSCL_RNG_GLIBC) {
    GlibcRand g; g.seed((unsigned)seed);
    for (int i = 0; i < n; i++) out[i] = g.next();
  } else if (rng_kind == FSCL_RNG_PHILOX) {
    for (int i = 0; i < n; i++) out[i] = philox_draw(seed, t, stream, (uint32_t)i);
  } else return fail(FSCL_ERR_ARG, "unknown generator");
  return FSCL_OK;
}

// --------------------------------------------------------------------------- p-value projection
// README.md:11 of the reference: "P-values are projected by estimating the non-centrality parameter of
// the chisq distribution of twice the composite log-likelihood ratio and then estimating the integral
// from the actual data CLR to infinity".  The reference never implemented it, so
// this is a DEFINITION, not a port: df is fixed by the caller (default 2: sweep position and strength),
// lambda = max(mean of the stored permuted CLRs - df, 0) (method of moments), p = P[ncx2(df, lambda) >= CLR].
#include "fscl_ncx2.cuh"
using fscl::ncx2_sf;

extern "C" int fscl_ncx2_sf(int n, const double *x, double df, const double *lambda, double *out) {
  if (n < 0 || !x || !lambda || !out || !(df > 0)) return fail(FSCL_ERR_ARG, "bad argument");
  for (int i = 0; i < n; i++) out[i] = ncx2_sf(x[i], df, lambda[i]);
  return FSCL_OK;
}

extern "C" int fscl_project_pvalues(int n_pts, const fscl_pt_t *pts, const float *permute_clr, double df,
                                    double *lambda_out, double *p_out) {
  if (n_pts < 0 || (n_pts > 0 && (!pts || !permute_clr || !p_out)) || !(df > 0)) return fail(FSCL_ERR_ARG, "bad argument");
  for (int i = 0; i < n_pts; i++) {
    const int m = std::min(pts[i].permute_n, FSCL_NULL_SAVE);
    double mean = 0.0;
    for (int q = 0; q < m; q++) mean += (double)permute_clr[(size_t)i * FSCL_NULL_SAVE + q];
    mean = m > 0 ? mean / m : 0.0;
    const double lambda = std::max(mean - df, 0.0);
    if (lambda_out) lambda_out[i] = lambda;
    p_out[i] = m > 0 ? ncx2_sf(pts[i].clr, df, lambda) : 1.0;
  }
  return FSCL_OK;
}

// --------------------------------------------------------------------------- bench / stats
// debug builds (-DFSCL_TIMING, tools/probe_phases.py): slots 2..15 hold phase latencies summed over CTAs
extern "C" int fscl_debug_counters(void *ctx, unsigned long long *out16) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !out16) return fail(FSCL_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out16, c->d_terms, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return FSCL_OK;
}

// -DFSCL_HIST builds: SNP-terms of the COUNT launches by [pass: coarse, fine][path: direct, covering, crossing,
// behind-a-bound][candidate 0..15] (tools/probe_terms.py)
extern "C" int fscl_debug_histogram(void *ctx, unsigned long long *out128) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !out128) return fail(FSCL_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out128, c->d_terms + 16, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return FSCL_OK;
}

extern "C" int fscl_set_exact_candidates(void *ctx, int on) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(FSCL_ERR_ARG, "null context");
  c->no_prune = on != 0;
  return FSCL_OK;
}

extern "C" int fscl_set_endpoint_sharing(void *ctx, int on) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(FSCL_ERR_ARG, "null context");
  c->no_chain = on == 0;
  return FSCL_OK;
}

extern "C" int fscl_get_stats(void *ctx, fscl_stats_t *s) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !s) return fail(FSCL_ERR_ARG, "bad argument");
  cudaSetDevice(c->device);
  unsigned long long redo = 0, near[3] = {0, 0, 0};
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(&redo, c->d_terms + 1, sizeof redo, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(near, c->d_terms + kCtrNearAlpha, sizeof near, cudaMemcpyDeviceToHost));
  c->stats.bound_failures = (int32_t)std::min<unsigned long long>(redo, 0x7FFFFFFFull);
  c->stats.near_ties_alpha = (int64_t)near[0];
  c->stats.near_ties_position = (int64_t)near[1];
  c->stats.fine_bound_failures = (int64_t)near[2];
  *s = c->stats;
  return FSCL_OK;
}
extern "C" int fscl_reset_stats(void *ctx) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(FSCL_ERR_ARG, "bad argument");
  cudaSetDevice(c->device);
  cudaMemsetAsync(c->d_terms + 1, 0, sizeof(unsigned long long), c->stream);
  cudaMemsetAsync(c->d_terms + kCtrNearAlpha, 0, 3 * sizeof(unsigned long long), c->stream);
  memset(&c->stats, 0, sizeof c->stats);
  return FSCL_OK;
}

// Whole permuted scans with everything resident in HBM: n_perm_slots Philox
// permutations (t = 0..n-1, seed 1) of the observed classes are built on the
// device and every cell of every permuted genome is searched (no pruning, so
// scans are equal-work).  n_perm_slots == 0 scans the observed genome once.
extern "C" int fscl_bench_scan(void *ctx, int G, int n_perm_slots, int count_terms, int64_t *terms_out,
                               fscl_stats_t *stats_out) {
  Ctx *c = (Ctx *)ctx;
  if (!c || G <= 0 || n_perm_slots < 0) return fail(FSCL_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  std::vector<Task> cells;
  enum_cells(c, G, &cells);
  int ns = 1;
  if (n_perm_slots > 0) {
    ns = n_perm_slots;
    int rc = ensure_slots(c, ns); if (rc) return rc;
    DrawSource src; src.init(FSCL_RNG_PHILOX, 1);
    BlockGen gen{c->n, c->h_pos.data(), c->chr_end_index.data(), 0.1, 1e6};
    std::vector<int4> blocks; std::vector<int> off(1, 0);
    int d7 = 0;
    for (int q = 0; q < ns; q++) { src.set_perm((uint32_t)q); gen.generate(src, blocks, &d7); off.push_back((int)blocks.size()); }
    rc = upload_blocks(c, blocks, off); if (rc) return rc;
    rc = apply_blocks(c, ns); if (rc) return rc;
    c->stats.n_blocks += (int64_t)blocks.size();
  }
  c->h_tasks.resize((size_t)ns * cells.size());
  const bool cell_major = getenv("FSCL_B200_CELL_MAJOR") != nullptr;      // experiment: task order (cell, slot) instead of (slot, cell)
  for (int q = 0; q < ns; q++)
    for (size_t a = 0; a < cells.size(); a++) {
      Task t = cells[a]; t.slot = q;
      if (n_perm_slots > 0) { t.start_pos = t.start_pos - (t.start_pos % G); t.end_pos = t.start_pos + G; }
      c->h_tasks[cell_major ? a * (size_t)ns + q : (size_t)q * cells.size() + a] = t;
    }
  if (count_terms) CK(cudaMemsetAsync(c->d_terms, 0, sizeof(unsigned long long), c->stream));
  int rc = run_tasks(c, n_perm_slots > 0, count_terms ? kFlagCount : 0, false);
  if (rc) return rc;
  if (count_terms && terms_out) {
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, c->d_terms, sizeof v, cudaMemcpyDeviceToHost));
    *terms_out = (int64_t)v;
  }
  c->stats.n_scored += ns;
  if (stats_out) *stats_out = c->stats;
  return FSCL_OK;
}

extern "C" int fscl_fp64_peak(int device, double *tflops) {
  if (!tflops) return fail(FSCL_ERR_ARG, "bad argument");
  if (fscl_device_count() <= 0) return fail(FSCL_ERR_NODEVICE, "no CUDA device");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  int grid = prop.multiProcessorCount * 8, iters = 1 << 16;
  double *d = nullptr;
  CK(dev_malloc(&d, sizeof(double) * grid * 256));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0, 0));
    launch_fma_peak(d, iters, grid, 0);
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 8.0 * iters * (double)grid * 256.0 / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
  }
  dev_free(d, 0); cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return FSCL_OK;
}
