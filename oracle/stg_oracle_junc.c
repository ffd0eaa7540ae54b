/* oracle/stg_oracle_junc.c — TEST INFRASTRUCTURE ONLY (see stg_oracle.h).
 *
 * CPU restatement of the junction pass that opens build_graphs(), short-read de novo path
 * (no --viral, no --mix / -L, no --rseq genomic sequence, no --ptf point features, no guides):
 *   ejunction = junction re-sorted by (end, start)        rlink.cpp:14387-14389, juncCmpEnd :1236-1244,
 *                                                         with the reference's own quicksort (gclib/GVec.hh:896-920:
 *                                                         Hoare partition, middle pivot, NOT stable)
 *   per-start / per-end support sums, opposite-strand      rlink.cpp:14391-14531
 *   conflict on equal coordinates, `higherr` detection
 *   higherr pass: bad junctions (all supporting reads      rlink.cpp:14538-14789 (non-viral branch :14591-14787)
 *   mismatchy) marked mm=-1 / demoted to a better          ok_to_demote :362-370, junc_ratio_threshold :354-359
 *   neighbour (index stored negated in nreads or
 *   nreads_good)
 *   good_junc on every stranded junction (on a copy)       rlink.cpp:13967-14070
 * parity: pinned — tests/test_oracle_vs_reference.py compares with dumps taken INSIDE the reference's build_graphs
 * (hook inserted before rlink.cpp:14843 by oracle/Makefile) on generated batches and on the golden fixtures.
 */
#include "stg_oracle.h"
#include "stg_oracle_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DROP 0.5          /* rlink.h:13 */
#define ERROR_PERC 0.1    /* rlink.h:14 */
#define CHI_WIN 100       /* rlink.h:17 */
#define LONGINTRON 100000u /* rlink.h:31 */

static int junc_cmp_end(const junc_t* a, const junc_t* b) { /* rlink.cpp:1236-1244 */
  if (a->end < b->end) return -1;
  if (a->end > b->end) return 1;
  if (a->start < b->start) return -1;
  if (a->start > b->start) return 1;
  return 0;
}

/* gclib/GVec.hh:896-920 on an index array */
void orc_qsort_ref(int32_t* list, const junc_t* J, int L, int R, int (*cmp)(const junc_t*, const junc_t*)) {
  int I, Jx;
  do {
    I = L; Jx = R;
    const junc_t* P = &J[list[L + ((R - L) >> 1)]];
    do {
      while (cmp(&J[list[I]], P) < 0) I++;
      while (cmp(&J[list[Jx]], P) > 0) Jx--;
      if (I <= Jx) {
        int32_t t = list[I]; list[I] = list[Jx]; list[Jx] = t;
        I++; Jx--;
      }
    } while (I <= Jx);
    if (L < Jx) orc_qsort_ref(list, J, L, Jx, cmp);
    L = I;
  } while (I < R);
}

static double junc_ratio_threshold(double na) { /* rlink.cpp:354-359 */
  if (na <= 25.0) return 1.0;
  else if (na <= 100.0) return 0.50;
  else if (na <= 200.0) return 0.25;
  else return 0.10;
}

static int ok_to_demote(const junc_t* cur, const junc_t* better) { /* rlink.cpp:362-370 */
  const double nb = (better->nreads > 0.0 ? better->nreads : 0.0);
  if (nb <= 0.0) return 0;
  const double na = (cur->nreads > 0.0 ? cur->nreads : 0.0);
  const double thr = junc_ratio_threshold(na);
  return (na / nb) <= thr;
}

/* rlink.cpp:13967-14070 for !eonly, !longreads; returns the verdict, *strand_out = jd.strand afterwards */
int orc_good_junc(const junc_t* jd_in, int64_t refstart, const float* cov, int64_t len, int junctionthr,
                  int8_t* strand_out) {
  junc_t jd = *jd_in;
  int res = 1;
  if (jd.guide_match) goto done;
  if (jd.nreads_good < junctionthr) { jd.strand = 0; res = 0; goto done; }
  {
    int mismatch = 0;
    if (jd.nm && round(jd.nm) == round(jd.nreads)) {
      mismatch = 1;
      if (jd.end - jd.start > LONGINTRON && jd.nreads < CHI_WIN * ERROR_PERC) { jd.strand = 0; res = 0; goto done; }
    }
    float mult = 1 / ERROR_PERC;
    int bw = 5;
    int j = (int)((int64_t)jd.start - refstart);
    double lleftcov = 0, lrightcov = 0;
    if (j + bw + 1 < len && j >= bw) {
      lleftcov = orc_get_cov(cov, len, 1, (uint32_t)(j - bw + 1), (uint32_t)j);
      lrightcov = orc_get_cov(cov, len, 1, (uint32_t)(j + 1), (uint32_t)(j + bw));
    }
    if (lleftcov > 1 / ERROR_PERC && jd.leftsupport * mult < ERROR_PERC * lleftcov &&
        (mismatch || lrightcov > lleftcov * (1 - ERROR_PERC))) { jd.strand = 0; res = 0; goto done; }
    j = (int)((int64_t)jd.end - refstart - 1);
    double rleftcov = 0, rrightcov = 0;
    if (j + bw + 1 < len && j >= bw) {
      rleftcov = orc_get_cov(cov, len, 1, (uint32_t)(j - bw + 1), (uint32_t)j);
      rrightcov = orc_get_cov(cov, len, 1, (uint32_t)(j + 1), (uint32_t)(j + bw));
    }
    if (rrightcov > 1 / ERROR_PERC && jd.rightsupport * mult < rrightcov * ERROR_PERC &&
        (mismatch || rleftcov > rrightcov * (1 - ERROR_PERC))) { jd.strand = 0; res = 0; goto done; }
    if (jd.nreads * 10 < ERROR_PERC * jd.leftsupport || jd.nreads * 10 < ERROR_PERC * jd.rightsupport) {
      jd.strand = 0; res = 0; goto done;
    }
  }
done:
  *strand_out = jd.strand;
  return res;
}

/* the pass itself: returns the junction rows (malloc'ed, n of them, with `extra` spare slots) and the ejunction order */
void orc_junction_state(const stg_packed_batch* B, const orc_params* P, int64_t b, const float* cov,
                        const double* good_in, const double* left_in, const double* right_in, int extra,
                        junc_t** J_out, int32_t** ej_out, int* n_out) {
  const uint32_t j0 = B->b_junc_off[b];
  const int n = (int)(B->b_junc_off[b + 1] - j0);
  *J_out = NULL; *ej_out = NULL; *n_out = n;
  if (n == 0) return;
  const int64_t refstart = B->b_start[b];
  const int64_t len = orc_bundle_len(B, b);
  const uint32_t junctionsupport = P->junctionsupport, sserror = 25; /* stringtie.cpp:141 */
  const int junctionthr = P->junctionthr;
  junc_t* J = (junc_t*)malloc(sizeof(junc_t) * (size_t)(n + extra));
  int32_t* ej = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + extra));
  for (int i = 0; i < n; i++) {
    J[i].start = B->j_start[j0 + i]; J[i].end = B->j_end[j0 + i];
    J[i].strand = B->j_strand[j0 + i]; J[i].guide_match = B->j_guide_match[j0 + i];
    J[i].consleft = -1; J[i].consright = -1;
    J[i].nreads = B->j_nreads[j0 + i]; J[i].nm = B->j_nm[j0 + i]; J[i].mm = B->j_mm[j0 + i];
    J[i].nreads_good = good_in[i]; J[i].leftsupport = left_in[i]; J[i].rightsupport = right_in[i];
    ej[i] = i;
  }
  orc_qsort_ref(ej, J, 0, n - 1, junc_cmp_end); /* GList(junction) copy, then setSorted(juncCmpEnd): rlink.cpp:14387-14389 */
#define EJ(k) J[ej[k]]

  /* ---- per-start / per-end support: rlink.cpp:14391-14531 (gseq == NULL, checkfeat == false, !mixedMode) ---- */
  uint32_t start = 0, end = 0;
  double leftsupport[2] = {0, 0}, rightsupport[2] = {0, 0};
  int higherr = 0;
  int8_t leftcons = -1, rightcons = -1;
  for (int i = 0; i < n; i++) {
    if (!higherr && J[i].strand && J[i].nm == J[i].nreads && !J[i].guide_match) higherr = 1;
    if (J[i].start != start) {
      int j = i - 1;
      while (j >= 0 && J[j].start == start) {
        J[j].consleft = leftcons;
        if (!J[j].guide_match && !leftcons && J[j].nreads_good < DROP / ERROR_PERC) J[j].strand = 0;
        if (J[j].strand) J[j].leftsupport = leftsupport[(1 + J[j].strand) / 2];
        j--;
      }
      j = i + 1;
      if (J[i].strand)
        while (j < n && J[j].start == J[i].start && J[j].end == J[i].end) {
          if (J[j].strand && J[i].strand != J[j].strand) {
            if (J[i].nreads > J[j].nreads && J[i].nreads_good > J[j].nreads_good) J[j].strand = 0;
            if (J[i].nreads < J[j].nreads && J[i].nreads_good < J[j].nreads_good) J[i].strand = 0;
          }
          j++;
        }
      leftsupport[0] = 0; leftsupport[1] = 0;
      if (J[i].strand) leftsupport[(1 + J[i].strand) / 2] = J[i].leftsupport;
      start = J[i].start;
    } else if (J[i].strand) leftsupport[(1 + J[i].strand) / 2] += J[i].leftsupport;

    if (EJ(i).end != end) {
      int j = i - 1;
      while (j >= 0 && EJ(j).end == end) {
        EJ(j).consright = rightcons;
        if (!EJ(j).guide_match && !rightcons && EJ(j).nreads_good < DROP / ERROR_PERC) EJ(j).strand = 0;
        if (EJ(j).strand) EJ(j).rightsupport = rightsupport[(1 + EJ(j).strand) / 2];
        j--;
      }
      rightsupport[0] = 0; rightsupport[1] = 0;
      if (EJ(i).strand) rightsupport[(1 + EJ(i).strand) / 2] = EJ(i).rightsupport;
      end = EJ(i).end;
    } else if (EJ(i).strand) rightsupport[(1 + EJ(i).strand) / 2] += EJ(i).rightsupport;
  }

  /* ---- higherr: rlink.cpp:14538-14787, non-viral branch ---- */
  if (higherr) {
    const uint32_t juncsupport = junctionsupport;
    const float tolerance = 1 - ERROR_PERC;
    for (int i = 1; i < n; i++) {
      if (J[i].strand) {
        if (J[i].nm && !J[i].guide_match && J[i].nm >= J[i].nreads) {
          if (J[i].nreads_good >= 0 && J[i].mm >= 0) {
            if (J[i].nreads_good < 1.25 * junctionthr) J[i].mm = -1;
            else {
              int point = (int)((int64_t)J[i].start - refstart);
              double leftcov = orc_get_cov(cov, len, 1, (uint32_t)point, (uint32_t)point);
              point = (int)((int64_t)J[i].start + 1 - refstart);
              double rightcov = orc_get_cov(cov, len, 1, (uint32_t)point, (uint32_t)point);
              if (rightcov > tolerance * leftcov) J[i].mm = -1;
            }
          }
          int j = i - 1;
          float support = 0;
          int searchjunc = 1, reliable = 0;
          while (j > 0 && J[i].start - J[j].start < juncsupport) {
            if (J[j].strand == J[i].strand) {
              if (J[j].start == J[i].start) {
                if (J[j].nreads < 0) {
                  int point = (int)(-J[j].nreads);
                  if (ok_to_demote(&J[i], &J[point])) { J[i].nreads = J[j].nreads; searchjunc = 0; }
                }
                break;
              } else if (J[j].guide_match || J[j].nm < J[j].nreads) {
                if (J[j].leftsupport > J[i].leftsupport * tolerance) {
                  if (ok_to_demote(&J[i], &J[j])) { reliable = 1; J[i].nreads = -j; support = (float)J[j].leftsupport; break; }
                }
              } else if (J[j].leftsupport > support && J[i].start - J[j].start < sserror &&
                         J[j].leftsupport * tolerance > J[i].leftsupport) {
                if (ok_to_demote(&J[i], &J[j])) { J[i].nreads = -j; support = (float)J[j].leftsupport; }
              }
            }
            j--;
          }
          if (searchjunc) {
            j = i + 1;
            int dist = (int)juncsupport;
            if (J[i].nreads < 0) dist = (int)(J[i].start - J[abs((int)J[i].nreads)].start);
            while (j < n && J[j].start - J[i].start < juncsupport) {
              if (J[j].strand == J[i].strand && J[j].start != J[i].start) {
                int d = (int)(J[j].start - J[i].start);
                if (J[j].guide_match || J[j].nm < J[j].nreads) {
                  if ((d < dist || (d == dist && J[j].leftsupport > support)) && J[j].leftsupport > J[i].leftsupport * tolerance) {
                    if (ok_to_demote(&J[i], &J[j])) { J[i].nreads = -j; support = (float)J[j].leftsupport; break; }
                  }
                } else if (!reliable && J[j].leftsupport > support && (uint32_t)d < sserror &&
                           J[j].leftsupport * tolerance > J[i].leftsupport) {
                  if (ok_to_demote(&J[i], &J[j])) { J[i].nreads = -j; support = (float)J[j].leftsupport; }
                }
              }
              j++;
            }
          }
        }
      }
      if (EJ(i).strand) {
        if (EJ(i).nm && !EJ(i).guide_match && EJ(i).nm >= EJ(i).nreads) {
          if (EJ(i).nreads_good >= 0 && EJ(i).mm >= 0) {
            if (EJ(i).nreads_good < 1.25 * junctionthr) EJ(i).mm = -1;
            else {
              int point = (int)((int64_t)EJ(i).end - 1 - refstart);
              double leftcov = orc_get_cov(cov, len, 1, (uint32_t)point, (uint32_t)point);
              point = (int)((int64_t)EJ(i).end - refstart);
              double rightcov = orc_get_cov(cov, len, 1, (uint32_t)point, (uint32_t)point);
              if (leftcov > tolerance * rightcov) EJ(i).mm = -1;
            }
          }
          int j = i - 1;
          float support = 0;
          int searchjunc = 1, reliable = 0;
          while (j > 0 && EJ(i).end - EJ(j).end < juncsupport) {
            if (EJ(j).strand == EJ(i).strand) {
              if (EJ(j).end == EJ(i).end) {
                if (EJ(j).nreads_good < 0) {
                  int point = (int)(-EJ(j).nreads_good);
                  if (ok_to_demote(&EJ(i), &EJ(point))) { EJ(i).nreads_good = EJ(j).nreads_good; searchjunc = 0; }
                }
                break;
              } else if (EJ(j).guide_match || EJ(j).nm < EJ(j).nreads) {
                if (EJ(j).rightsupport > EJ(i).rightsupport * tolerance) {
                  if (ok_to_demote(&EJ(i), &EJ(j))) { reliable = 1; EJ(i).nreads_good = -j; support = (float)EJ(j).rightsupport; break; }
                }
              } else if (EJ(j).rightsupport > support && EJ(i).end - EJ(j).end < sserror &&
                         EJ(j).rightsupport * tolerance > EJ(i).rightsupport) {
                if (ok_to_demote(&EJ(i), &EJ(j))) { EJ(i).nreads_good = -j; support = (float)EJ(j).rightsupport; }
              }
            }
            j--;
          }
          if (searchjunc) {
            j = i + 1;
            int dist = (int)juncsupport;
            if (EJ(i).nreads_good < 0) dist = (int)(EJ(i).end - EJ(abs((int)EJ(i).nreads_good)).end);
            while (j < n && EJ(j).end - EJ(i).end < juncsupport) {
              if (EJ(j).strand == EJ(i).strand && EJ(j).end != EJ(i).end) {
                int d = (int)(EJ(j).end - EJ(i).end);
                if (EJ(j).guide_match || EJ(j).nm < EJ(j).nreads) {
                  if ((d < dist || (d == dist && EJ(j).rightsupport > support)) && EJ(j).rightsupport > EJ(i).rightsupport * tolerance) {
                    if (ok_to_demote(&EJ(i), &EJ(j))) { EJ(i).nreads_good = -j; support = (float)EJ(j).rightsupport; break; }
                  }
                } else if ((!reliable && EJ(j).rightsupport > support && EJ(j).end - EJ(i).end < sserror &&
                            EJ(j).rightsupport * tolerance > EJ(i).rightsupport) ||
                           ((int)(EJ(j).end - EJ(i).end) < dist && (EJ(j).guide_match || EJ(j).nm < EJ(j).nreads))) {
                  if (ok_to_demote(&EJ(i), &EJ(j))) { EJ(i).nreads_good = -j; support = (float)EJ(j).rightsupport; }
                }
              }
              j++;
            }
          }
        }
      }
    }
  }

#undef EJ
  *J_out = J; *ej_out = ej;
}

void orc_junction_pass(const stg_packed_batch* B, const orc_params* P, int64_t b, const float* cov,
                       const double* good_in, const double* left_in, const double* right_in, int8_t* o_strand,
                       double* o_nreads, double* o_nreads_good, double* o_left, double* o_right, double* o_mm,
                       int32_t* o_ej, int8_t* o_good, int8_t* o_good_strand) {
  junc_t* J; int32_t* ej; int n;
  orc_junction_state(B, P, b, cov, good_in, left_in, right_in, 0, &J, &ej, &n);
  if (n == 0) return;
  const int64_t refstart = B->b_start[b];
  const int64_t len = orc_bundle_len(B, b);
  const int junctionthr = P->junctionthr;
  for (int i = 0; i < n; i++) {
    o_strand[i] = J[i].strand; o_nreads[i] = J[i].nreads; o_nreads_good[i] = J[i].nreads_good;
    o_left[i] = J[i].leftsupport; o_right[i] = J[i].rightsupport; o_mm[i] = J[i].mm;
    o_ej[i] = ej[i];
    o_good[i] = -1; o_good_strand[i] = J[i].strand;
    if (J[i].strand) o_good[i] = (int8_t)orc_good_junc(&J[i], refstart, cov, len, junctionthr, &o_good_strand[i]);
  }
  free(J); free(ej);
}
