/* Test infrastructure.  Textually includes the UNMODIFIED reference translation
 * unit (found on the include path, never copied) and adds non-static wrappers
 * so tests can call its file-local functions directly:
 *   snp_block_permute   /root/reference/scan-chromosome.c:336
 *   init_scan_result    /root/reference/scan-chromosome.c:58
 *   search_maxpos       /root/reference/scan-chromosome.c:126
 */
#include "scan-chromosome.c"

void fx_snp_block_permute(snp_t *p_snps, snp_t *snps, int n_snps, double permute_nbp,
                          double scan_width_mb) {
  snp_block_permute(p_snps, snps, n_snps, permute_nbp, scan_width_mb);
}
void fx_init_scan_result(scan_pt_t *pt, int chr, snp_t *snps, chr_limits_t *limits, int eval_range,
                         int pos) {
  init_scan_result(pt, chr, snps, limits, eval_range, pos);
}
void fx_search_maxpos(scan_pt_t *out, int chr, int start_pos, int end_pos, snp_t *snps,
                      chr_limits_t *limits, int eval_range, int bp_resl, sm_ptable_t *sm_p) {
  *out = search_maxpos(chr, start_pos, end_pos, snps, limits, eval_range, bp_resl, sm_p);
}
int fx_sizeof_snp(void) { return (int)sizeof(snp_t); }
int fx_sizeof_scan_pt(void) { return (int)sizeof(scan_pt_t); }
