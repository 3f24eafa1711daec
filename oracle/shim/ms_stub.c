/* Link stub (test infrastructure).  The reference's ms path needs a
 * bison/flex-generated parser that cannot be produced in this image and is
 * broken at the surveyed commit (box at top, item 4).  The three
 * entry points declared at /root/reference/fscl.h:118-123 are stubbed so the
 * SNP-file path links; calling them is a fatal error. */
#include <stdio.h>
#include <stdlib.h>
struct scan_t_tag;
static void *die(const char *what) {
  fprintf(stderr, "oracle/_ref: %s is not available (ms parser not built)\n", what);
  exit(1);
  return 0;
}
void ms_openfile(char *f) { (void)f; die("ms_openfile"); }
void *ms_background(char *f, int a, int b, int c, int d) { (void)f;(void)a;(void)b;(void)c;(void)d; return die("ms_background"); }
void *ms_next_block(int a, int b, int c, int d) { (void)a;(void)b;(void)c;(void)d; return die("ms_next_block"); }
