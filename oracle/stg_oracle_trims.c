/* oracle/stg_oracle_trims.c — TEST INFRASTRUCTURE ONLY (see stg_oracle.h).
 *
 * CPU restatement of find_all_trims (rlink.cpp:2118-2419; part of SURVEY.md §8 row a10), short-read path without point
 * features (tstartend empty): for every base of a region compare the mean signed coverage (get_cov_sign, :1842) on
 * its left and on its right — whole-region halves for regions under 301 bp, windows of CHI_WIN+CHI_THR = 150 bp
 * otherwise, in four phases with their own normalisation and threshold — keep the sharpest source / sink drop within
 * 50 bp, then re-validate every kept point against up to 300 bp flanks (rejected points stay in the list with pos = 0,
 * as in the reference).
 * parity: pinned — tests/test_oracle_vs_reference.py compares with the reference's own find_all_trims called from the
 * dump hook on every bundlenode (and on its halves) of the fixtures.
 */
#include "stg_oracle.h"
#include <math.h>
#include <stdlib.h>

#define DROP 0.5
#define ERROR_PERC 0.1
#define CHI_WIN 100
#define CHI_THR 50
#define LONGINTRONANCHOR 25u

typedef struct { uint32_t* pos; float* ab; uint8_t* st; int n, cap; } tp_t;

static void tp_add(tp_t* T, uint32_t pos, double ab, int start) {
  if (T->n < T->cap) { T->pos[T->n] = pos; T->ab[T->n] = (float)ab; T->st[T->n] = (uint8_t)start; }
  T->n++;
}

/* one candidate position: the shared tail of the four loops (rlink.cpp:2157-2187 and its three siblings) */
static void consider(tp_t* T, uint32_t i, double covleft, double covright, double localdrop, double* lastdrop, double norm,
                     int sink_plus_one) {
  const int last = T->n - 1;
  if (covleft < covright) {
    double thisdrop = covleft / covright;
    if (thisdrop < localdrop) {
      if (!T->n || (!T->st[last] && i + 1 - (uint32_t)(int)T->pos[last] > CHI_THR)) {
        tp_add(T, i + 1, (covright - covleft) / norm, 1);
        *lastdrop = thisdrop;
      } else if (thisdrop < *lastdrop) {
        *lastdrop = thisdrop;
        T->pos[last] = i + 1; T->ab[last] = (float)((covright - covleft) / norm); T->st[last] = 1;
      }
    }
  } else if (covright != covleft) {
    double thisdrop = sink_plus_one ? (covright + 1) / (covleft + 1) : covright / covleft;
    if (thisdrop < localdrop) {
      if (!T->n || (T->st[last] && i - (uint32_t)(int)T->pos[last] > CHI_THR)) {
        tp_add(T, i, (covleft - covright) / norm, 0);
        *lastdrop = thisdrop;
      } else if (thisdrop < *lastdrop) {
        *lastdrop = thisdrop;
        T->pos[last] = i; T->ab[last] = (float)((covleft - covright) / norm); T->st[last] = 0;
      }
    }
  }
}

int orc_trims_capacity(uint32_t start, uint32_t end) { return (int)((end - start + 1) / CHI_THR) + 2; }

/* cov: the bundle's bpcov [3*len]; sno in {0, 2}; start/end genomic; returns the number of trim points written
 * (<= orc_trims_capacity) */
int orc_find_all_trims(const float* cov, int64_t blen, int64_t refstart, int sno, uint32_t start, uint32_t end,
                       uint32_t* o_pos, float* o_abund, uint8_t* o_start) {
  tp_t T = {o_pos, o_abund, o_start, 0, orc_trims_capacity(start, end)};
  int len = (int)(end - start + 1);
  if (len < CHI_THR) return 0;
#define GCS(a, b) orc_get_cov_sign(cov, blen, sno, (uint32_t)((int64_t)(a) - refstart), (uint32_t)((int64_t)(b) - refstart))
  double localdrop = ERROR_PERC / DROP;
  if (len < 2 * (CHI_WIN + CHI_THR) + 1) {
    if (len < CHI_WIN) localdrop = ERROR_PERC / (10 * DROP);
    double lastdrop = localdrop;
    for (uint32_t i = start + LONGINTRONANCHOR; i < end - LONGINTRONANCHOR; i++) {
      double covleft = GCS(start, i - 1) / (i - start);
      double covright = GCS(i, end) / (end - i + 1);
      consider(&T, i, covleft, covright, localdrop, &lastdrop, DROP, 1);
    }
    return T.n < T.cap ? T.n : T.cap;
  }
  const int winlen = CHI_WIN + CHI_THR;
  double lastdrop = localdrop;
  for (uint32_t i = start + CHI_THR - 1; i < start + winlen - 1; i++) {
    double covleft = GCS(start, i) / (i - start);
    double covright = GCS(i + 1, i + winlen) / winlen;
    consider(&T, i, covleft, covright, localdrop, &lastdrop, DROP, 0);
  }
  localdrop = ERROR_PERC / DROP;
  if (!T.n) lastdrop = localdrop;
  len += (int)(start - (uint32_t)winlen);
  for (uint32_t i = start + winlen - 1; i < (uint32_t)len; i++) {
    double covleft = GCS(i - winlen + 1, i);
    double covright = GCS(i + 1, i + winlen);
    consider(&T, i, covleft, covright, localdrop, &lastdrop, DROP * winlen, 0);
  }
  localdrop = ERROR_PERC;
  for (uint32_t i = (uint32_t)len; i < end - CHI_THR - 1; i++) {
    double covleft = GCS(i - winlen + 1, i) / winlen;
    double covright = GCS(i + 1, end) / (end - i);
    consider(&T, i, covleft, covright, localdrop, &lastdrop, DROP, 0);
  }
  if (T.n > T.cap) T.n = T.cap;
  /* global re-validation: rlink.cpp:2351-2404 */
  localdrop = ERROR_PERC / DROP;
  uint32_t laststart = start;
  for (int i = 0; i < T.n; i++) {
    uint32_t midpos = T.pos[i];
    if (T.st[i]) midpos--;
    if (midpos - laststart > 3 * CHI_WIN) laststart = midpos - 3 * CHI_WIN;
    double covleft = GCS(laststart, midpos) / (midpos - laststart + 1);
    uint32_t endpos = end;
    if (i < T.n - 1) endpos = T.pos[i + 1];
    if (endpos - midpos > 3 * CHI_WIN) endpos = midpos + 3 * CHI_WIN;
    double covright = GCS(midpos + 1, endpos) / (endpos - midpos);
    if ((T.st[i] && covleft < localdrop * covright) || (covright < localdrop * covleft)) {
      laststart = T.pos[i] + 1;
    } else {
      T.pos[i] = 0;
      int k = i - 1;
      while (k >= 0) {
        while (k >= 0 && !T.pos[k]) k--;
        if (k >= 0) {
          midpos = T.pos[k];
          if (T.st[k]) midpos--;
          uint32_t leftstart = start;
          int j = k - 1;
          while (j >= 0) { if (T.pos[j]) { leftstart = T.pos[j]; break; } j--; }
          if (midpos - leftstart > 3 * CHI_WIN) leftstart = midpos - 3 * CHI_WIN;
          covleft = GCS(leftstart, midpos) / (midpos - leftstart + 1);
          if (endpos - midpos > 3 * CHI_WIN) endpos = midpos + 3 * CHI_WIN;
          covright = GCS(midpos + 1, endpos) / (endpos - midpos);
          if ((T.st[k] && covleft < localdrop * covright) || (covright < localdrop * covleft)) break;
          else T.pos[k] = 0;
          k--;
        }
      }
    }
  }
#undef GCS
  return T.n;
}
