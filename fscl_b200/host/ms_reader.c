/* ms_reader.c -- Hudson `ms` text -> scan_t, behind the reference's own entry points
 *
 *   void    ms_openfile(char *ms_fname);                         fscl.h:119  (ms-input.c:13)
 *   scan_t *ms_background(char*, int, int, int, int);             fscl.h:120  (ms-input.c:25)
 *   scan_t *ms_next_block(int, int, int, int);                    fscl.h:122  (ms-input.c:93)
 *
 * so that the reference's unmodified main() (fscl.c:281-310, the `-m` branch) links
 * against it in place of ms-input.o + the bison/flex parser (ms-parser.y,
 * ms-scanner.lex), which cannot be generated in this image and are broken at the
 * surveyed commit (box at top, item 4: msparser_block() is called without
 * its scanner handle, chromosome limits are never filled in, n_depths is never
 * incremented in ms_next_block, folded sites get obs_freq = 0).  This is a hand-written
 * line parser; it implements the INTENDED semantics, which nothing in the reference pins:
 *
 *  - a block starts at a line beginning with "//"; "segsites: N"; optional "prob: x";
 *    "positions: f1 ... fN" (fractions of the segment); then one [01]+ row per
 *    haplotype; a blank line (or end of file) ends the block (ms-scanner.lex:11-34,
 *    ms-parser.y:96-246).  Newick tree lines between "//" and "segsites:" are skipped.
 *  - derived count of a site = number of '1' among haplotypes [first, first+size)
 *    (--ms-sample-first, --ms-sample-size; size 0 = all remaining); sites with count 0 or
 *    size are dropped (ms-input.c:69-72, :127-131).
 *  - position = (int)(fraction * segment_length)  (ms-input.c:64, :123).
 *  - --ms-folded: the site is folded and carries the MINOR count (fix of ms-input.c:74-77).
 *  - every block is one chromosome "ms<k>" of ONE scan object: ms_background() returns all
 *    blocks (background spectrum from all of them, the reference's first pass) and the
 *    first ms_next_block() after ms_openfile() returns the same object again, the second
 *    returns NULL -- so main()'s loop scans all blocks in ONE launch and writes ONE output
 *    file (the reference re-opens its output with "w" for every block,
 *    scan-chromosome.c:675, i.e. only the last block would survive).
 *  - depth index = distinct block sample sizes in first-seen order, identical in both
 *    passes (the tables built after the first pass are indexed by it).
 *
 * The same function is exported as fscl_ms_read() for the Python host layer.
 * Errors follow the reference's convention: message on stderr, exit(1).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fscl_dropin.h"

static char *g_path = NULL;     /* file named by the last ms_openfile() */
static int g_delivered = 0;     /* ms_next_block() calls since then */

static void fatal(const char *what, const char *unit, long n) {
  fprintf(stderr, "\nfscl_b200 ms reader: %s (%s %ld)\n", what, unit, n);
  exit(1);
}

static void *xrealloc(void *p, size_t n) {
  void *q = realloc(p, n ? n : 1);
  if (!q) { fprintf(stderr, "\nout of memory\n"); abort(); }
  return q;
}

static int is_blank(const char *s) {
  while (*s) { if (!isspace((unsigned char)*s)) return 0; s++; }
  return 1;
}

static int is_haplotype(const char *s, size_t *len_out) {
  size_t n = 0;
  while (s[n] == '0' || s[n] == '1') n++;
  size_t m = n;
  while (s[m]) { if (!isspace((unsigned char)s[m])) return 0; m++; }
  *len_out = n;
  return n > 0;
}

typedef struct {
  int segsites;
  double *positions;    /* [segsites] */
  int n_hap;
  char **rows;          /* [n_hap][segsites] */
  int have_positions;
} block_t;

static void block_reset(block_t *b) {
  for (int i = 0; i < b->n_hap; i++) free(b->rows[i]);
  free(b->rows); free(b->positions);
  memset(b, 0, sizeof *b);
  b->segsites = -1;
}

/* sites of one finished block appended to the scan object as chromosome "ms<k>" */
static void flush_block(scan_t *so, block_t *b, long block_index, int seglen, int folded, int first, int size,
                        long line_no) {
  if (b->segsites < 0) fatal("no segsites line", "block", block_index);
  if (b->segsites == 0) return;
  if (!b->have_positions) fatal("no positions line", "block", block_index);
  if (b->n_hap == 0) fatal("no haplotype rows", "block", block_index);
  const int n = size > 0 ? size : b->n_hap - first;
  if (first < 0 || n <= 0 || first + n > b->n_hap)
    fatal("sample range exceeds the haplotypes of the block", "ending at line", line_no);
  /* depth index, first-seen order (ms-input.c:49-56) */
  int m = 0;
  while (m < so->n_depths && so->sample_depths[m] != n) m++;
  if (m == so->n_depths) {
    so->sample_depths = (int *)xrealloc(so->sample_depths, sizeof(int) * (size_t)(so->n_depths + 1));
    so->sample_depths[so->n_depths++] = n;
  }
  const int chr = so->n_chromosomes;
  const int start_index = so->n_snps;
  int cap = so->n_snps;
  for (int i = 0; i < b->segsites; i++) {
    int d = 0;
    for (int j = first; j < first + n; j++) d += b->rows[j][i] == '1';
    if (d == 0 || d == n) continue;
    if (so->n_snps == cap) {
      cap = cap ? cap * 2 : 4096;
      so->snps = (snp_t *)xrealloc(so->snps, sizeof(snp_t) * (size_t)cap);
    }
    snp_t *s = so->snps + so->n_snps++;
    memset(s, 0, sizeof *s);
    s->chr = chr;
    s->pos = (int)(b->positions[i] * seglen);
    s->depth_p = m;
    if (folded) { s->folded = 1; s->obs_freq = d > n - d ? n - d : d; }
    else { s->folded = 0; s->obs_freq = d; }
  }
  const int cnt = so->n_snps - start_index;
  if (cnt == 0) return;
  /* ms prints positions in ascending order; a stable insertion pass repairs the rare inversion
     left by equal fractions (the scan needs non-decreasing positions per chromosome) */
  for (int i = start_index + 1; i < so->n_snps; i++) {
    snp_t v = so->snps[i];
    int j = i;
    while (j > start_index && so->snps[j - 1].pos > v.pos) { so->snps[j] = so->snps[j - 1]; j--; }
    so->snps[j] = v;
  }
  so->chr_limits = (chr_limits_t *)xrealloc(so->chr_limits, sizeof(chr_limits_t) * (size_t)(chr + 1));
  chr_limits_t *L = so->chr_limits + chr;
  char name[32];
  snprintf(name, sizeof name, "ms%ld", block_index);
  L->chr = chr;
  L->name = strdup(name);
  L->start_index = start_index;
  L->n_snps = cnt;
  L->start_pos = so->snps[start_index].pos;          /* snp-input.c:134-135: first / last site position */
  L->bp_length = so->snps[so->n_snps - 1].pos;
  so->n_chromosomes = chr + 1;
}

scan_t *fscl_ms_read(const char *path, int seglen, int folded, int first, int size) {
  if (seglen <= 0) fatal("--ms-segment-length must be positive", "given", (long)seglen);
  FILE *f = fopen(path, "r");
  if (!f) { fprintf(stderr, "\nCan't open ms input file \"%s\" (%s)\n", path, strerror(errno)); exit(1); }
  scan_t *so = (scan_t *)calloc(1, sizeof(scan_t));
  if (!so) abort();
  block_t b;
  memset(&b, 0, sizeof b);
  b.segsites = -1;
  int in_block = 0;
  long block_index = -1, line_no = 0;
  char *line = NULL;
  size_t cap = 0;
  ssize_t len;
  while ((len = getline(&line, &cap, f)) >= 0) {
    line_no++;
    char *s = line;
    while (*s == ' ' || *s == '\t') s++;
    if (s[0] == '/' && s[1] == '/') {
      if (in_block) flush_block(so, &b, block_index, seglen, folded, first, size, line_no);
      block_reset(&b);
      in_block = 1;
      block_index++;
      continue;
    }
    if (!in_block) continue;                       /* command line, seeds, anything before the first block */
    size_t hl;
    if (strncmp(s, "segsites:", 9) == 0) {
      b.segsites = atoi(s + 9);
      if (b.segsites < 0) fatal("negative segsites", "line", line_no);
    } else if (strncmp(s, "positions:", 10) == 0) {
      if (b.segsites < 0) fatal("positions before segsites", "line", line_no);
      b.positions = (double *)xrealloc(b.positions, sizeof(double) * (size_t)(b.segsites > 0 ? b.segsites : 1));
      char *p = s + 10, *e;
      int k = 0;
      for (;;) {
        double v = strtod(p, &e);
        if (e == p) break;
        if (k < b.segsites) b.positions[k] = v;
        k++;
        p = e;
      }
      if (k != b.segsites) fatal("positions do not match segsites", "line", line_no);
      b.have_positions = 1;
    } else if (strncmp(s, "prob:", 5) == 0 || s[0] == '[' || s[0] == '(') {
      /* probability line / Newick trees: not needed by the scan */
    } else if (is_haplotype(s, &hl)) {
      if ((int)hl != b.segsites) fatal("haplotype row does not match segsites", "line", line_no);
      b.rows = (char **)xrealloc(b.rows, sizeof(char *) * (size_t)(b.n_hap + 1));
      b.rows[b.n_hap] = (char *)xrealloc(NULL, hl + 1);
      memcpy(b.rows[b.n_hap], s, hl);
      b.rows[b.n_hap][hl] = 0;
      b.n_hap++;
    } else if (is_blank(s)) {
      if (b.n_hap > 0 || b.segsites == 0) {
        flush_block(so, &b, block_index, seglen, folded, first, size, line_no);
        block_reset(&b);
        in_block = 0;
      }
    } else {
      fatal("unexpected text in an ms block", "line", line_no);
    }
  }
  if (in_block) flush_block(so, &b, block_index, seglen, folded, first, size, line_no);
  block_reset(&b);
  free(line);
  fclose(f);
  if (so->n_snps == 0) { fprintf(stderr, "\nNo usable snps found in ms file \"%s\"\n", path); exit(1); }
  return so;
}

void ms_openfile(char *ms_fname) {
  free(g_path);
  g_path = strdup(ms_fname);
  g_delivered = 0;
  FILE *f = fopen(ms_fname, "r");          /* fail here, as the reference does (ms-input.c:18-21) */
  if (!f) { fprintf(stderr, "\nCan't open ms input file \"%s\" (%s)\n", ms_fname, strerror(errno)); exit(1); }
  fclose(f);
}

scan_t *ms_background(char *ms_fname, int ms_segment_length, int ms_folded, int ms_sample_first,
                      int ms_sample_size) {
  ms_openfile(ms_fname);
  return fscl_ms_read(ms_fname, ms_segment_length, ms_folded, ms_sample_first, ms_sample_size);
}

scan_t *ms_next_block(int ms_segment_length, int ms_folded, int ms_sample_first, int ms_sample_size) {
  if (!g_path) { fprintf(stderr, "\nfscl_b200 ms reader: ms_next_block before ms_openfile\n"); exit(1); }
  if (g_delivered++ > 0) return NULL;
  return fscl_ms_read(g_path, ms_segment_length, ms_folded, ms_sample_first, ms_sample_size);
}

/* release a scan object made here (what scan_free does, fscl.c:260-270; that one is static) */
void fscl_ms_free(scan_t *so) {
  if (!so) return;
  free(so->scan_pts); free(so->sample_depths); free(so->snps);
  for (int i = 0; i < so->n_chromosomes; i++) free(so->chr_limits[i].name);
  free(so->chr_limits);
  free(so);
}
