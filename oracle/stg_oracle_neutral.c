/* oracle/stg_oracle_neutral.c — TEST INFRASTRUCTURE ONLY (see stg_oracle.h).
 *
 * CPU restatement of the single-exon predictions of the neutral bundles (de novo part):
 * build_graphs, rlink.cpp:15500-15634, without guides / rawreads / long reads.  A neutral bundle (strand lane 1: groups
 * that touch no stranded group) whose coverage is not mostly multi-mapped is cut, bundlenode by bundlenode, at the trim
 * points find_all_trims (rlink.cpp:2118) finds; every piece longer than mintranscriptlen becomes a prediction with the
 * mean coverage get_cov (rlink.cpp:1830) gives on the all-strands lane.
 * parity: pinned — tests/test_oracle_vs_reference.py compares with the prediction list dumped INSIDE the reference's
 * build_graphs (hook before rlink.cpp:15637, arrays a19_*).
 */
#include "stg_oracle.h"
#include <stdlib.h>

#define CHI_WIN 100      /* rlink.h:17 */
#define ERROR_PERC 0.1   /* rlink.h:14 */

int orc_neutral_preds(const float* cov, int64_t blen, int64_t refstart, const orc_groups_out* G, int mintranscriptlen,
                      float mcov, int cap, uint32_t* o_start, uint32_t* o_end, double* o_cov, int32_t* o_tlen,
                      int32_t* o_geneno) {
  int np = 0, geneno = 0;
  const int b0 = G->bun_off[1], b1 = G->bun_off[2], n0 = G->bn_off[1];
#define EMIT(a, z, len) do { \
    if (np < cap) { \
      o_start[np] = (a); o_end[np] = (z); o_tlen[np] = (len); o_geneno[np] = geneno - 1; \
      o_cov[np] = orc_get_cov(cov, blen, 1, (uint32_t)((int64_t)(a) - refstart), (uint32_t)((int64_t)(z) - refstart)) / (len); \
    } \
    np++; t++; } while (0)
  for (int b = b0; b < b1; b++) {
    const float bc = G->bun_cov[b], bm = G->bun_multi[b];
    if (!(bc && (bm / bc) <= mcov * (1 - ERROR_PERC))) continue;
    int t = 1;
    for (int k = G->bun_start[b]; k != -1; k = G->bn_next[n0 + k]) {
      const uint32_t ns = G->bn_start[n0 + k], ne = G->bn_end[n0 + k];
      if (t == 1) geneno++;
      const int tcap = orc_trims_capacity(ns, ne);
      uint32_t* tp = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(tcap + 1));
      float* ta = (float*)malloc(sizeof(float) * (size_t)(tcap + 1));
      uint8_t* ts = (uint8_t*)malloc((size_t)tcap + 1);
      const int nt = orc_find_all_trims(cov, blen, refstart, 0, ns, ne, tp, ta, ts);
      uint32_t predstart = ns;
      const uint32_t predend = ne;
      for (int i = 0; i < nt; i++) {
        if (!tp[i]) continue;
        if (ts[i]) {
          const int len = (int)(tp[i] - CHI_WIN - predstart + 1);
          if (len > mintranscriptlen) EMIT(predstart, tp[i] - CHI_WIN, len);
          predstart = tp[i];
        } else {
          const int len = (int)(tp[i] - predstart + 1);
          if (len > mintranscriptlen) EMIT(predstart, tp[i], len);
          predstart = tp[i] + CHI_WIN;
        }
      }
      const int len = (int)(predend - predstart + 1);
      if (len > mintranscriptlen) EMIT(predstart, predend, len);
      free(tp); free(ta); free(ts);
    }
  }
#undef EMIT
  return np;
}
