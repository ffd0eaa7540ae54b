/* oracle/stg_oracle.c — TEST INFRASTRUCTURE ONLY (see stg_oracle.h).
 *
 * CPU restatement, in reference order and reference precision, of
 *   cov_edge_add            rlink.cpp:133-142
 *   add_read_to_cov         rlink.cpp:144-219
 *   count_good_junctions    rlink.cpp:16875-17063 (short-read path: resize, per-read loop, blocked clamped scan)
 *   get_cov / get_cov_sign  rlink.cpp:1830-1853
 * Float32 accumulation order is the reference's: reads in readlist order, per read its mate links in
 * pair_idx order (only links with pair_idx <= n emit), merged segments left to right, then the unpaired
 * remainder.  Built with -ffp-contract=off (the reference is a baseline x86-64 build without FMA).
 * parity: pinned — see tests/test_oracle_vs_reference.py and tests/golden/.
 */
#include "stg_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LONGINTRON 100000u      /* rlink.h:31 */
#define LONGINTRONANCHOR 25u    /* rlink.h:32 */
static const double EPSILON = 0.000001; /* rlink.h:25 */

int64_t orc_bundle_len(const stg_packed_batch* B, int64_t b) {
  return (int64_t)B->b_end[b] - (int64_t)B->b_start[b] + 3;
}

/* one difference-array deposit on lane `lane` (and on lane 1 when stranded): rlink.cpp:133-142 with the
 * caller's argument shift (start-refstart, end-refstart+1) folded in */
static inline void deposit(float* cov, int64_t len, int lane, int64_t start_rel, int64_t end_rel, float v,
                           int64_t* nivl) {
  int64_t a = start_rel + 1, z = end_rel + 2;
  if (nivl) (*nivl)++;
  if (!cov) return;
  if (a < 0 || z >= len) return; /* malformed input (read outside bundle): the reference would write out of bounds */
  cov[(int64_t)lane * len + a] += v;
  cov[(int64_t)lane * len + z] -= v;
  if (lane != 1) {
    cov[len + a] += v;
    cov[len + z] -= v;
  }
}

/* coverage deposits of read n of bundle [r0, r1): rlink.cpp:144-219 */
static void read_to_cov(const stg_packed_batch* B, uint32_t r0, uint32_t n, int64_t refstart, float* cov,
                        int64_t len, int64_t* nivl) {
  const uint32_t* ss = B->seg_start;
  const uint32_t* se = B->seg_end;
  uint32_t rs0 = B->r_seg_off[n], rs1 = B->r_seg_off[n + 1];
  int sno = (int)B->r_strand[n] + 1;
  float single_count = B->r_count[n];
  for (uint32_t k = B->r_pair_off[n]; k < B->r_pair_off[n + 1]; k++) {
    float pc = B->pair_cnt[k];
    single_count -= pc;
    int32_t pl = B->pair_idx[k];
    if (pl < 0) continue;              /* the reference does not guard this (rd[-1]); we define it as "no mate" */
    uint32_t np = r0 + (uint32_t)pl;
    if (np > n) continue;              /* the later mate deposits the fragment when its turn comes */
    uint32_t ps0 = B->r_seg_off[np], ps1 = B->r_seg_off[np + 1];
    int snop = (int)B->r_strand[np] + 1;
    int lane = sno;
    if (sno != snop) {
      if (sno == 1) lane = snop;
      else if (snop != 1) lane = 1;
    }
    uint32_t p = ps0, r = rs0;
    while (p < ps1 || r < rs1) {
      int64_t start, end;
      if (r < rs1) {
        if (p == ps1 || ss[p] > se[r]) { start = ss[r]; end = se[r]; r++; }
        else if (se[p] < ss[r]) { start = ss[p]; end = se[p]; p++; }
        else { /* the two mates overlap here: one merged interval */
          start = ss[p]; end = se[p];
          if ((int64_t)ss[r] < start) start = ss[r];
          p++;
          int grew = 1;
          while (grew) {
            grew = 0;
            if (r < rs1 && (int64_t)ss[r] <= end) { if ((int64_t)se[r] > end) end = se[r]; r++; grew = 1; }
            if (p < ps1 && (int64_t)ss[p] <= end) { if ((int64_t)se[p] > end) end = se[p]; p++; grew = 1; }
          }
        }
      } else { start = ss[p]; end = se[p]; p++; }
      deposit(cov, len, lane, start - refstart, end - refstart, pc, nivl);
    }
  }
  if ((double)single_count > EPSILON) {
    for (uint32_t s = rs0; s < rs1; s++)
      deposit(cov, len, sno, (int64_t)ss[s] - refstart, (int64_t)se[s] - refstart, single_count, nivl);
  }
}

/* junction support of read n: rlink.cpp:16890-17016 (short-read path) */
/* good/left/right point at the bundle's first junction row (outputs are bundle-local, inputs global at j0) */
static void read_junction_support(const stg_packed_batch* B, const orc_params* P, uint32_t j0, uint32_t n,
                                  double* good, double* left, double* right, int64_t* nj) {
  uint32_t s0 = B->r_seg_off[n], s1 = B->r_seg_off[n + 1];
  uint32_t nex = s1 - s0;
  if (nex < 2) return;
  float rdcount = B->r_count[n];
  if ((B->r_flags[n] & STG_RF_UNITIG) && rdcount > 1) rdcount = 1;
  const int32_t* rj = B->rjunc_idx + (s0 - n);
  uint32_t stackbuf[64];
  uint32_t* rsup = nex <= 64 ? stackbuf : (uint32_t*)malloc(sizeof(uint32_t) * nex);
  /* rsup[i] = longest exon among segs[i..nex-1]; running left maximum kept in lmax */
  uint32_t m = 0;
  for (uint32_t i = nex; i-- > 0;) {
    uint32_t l = B->seg_end[s0 + i] - B->seg_start[s0 + i] + 1;
    if (l > m) m = l;
    rsup[i] = m;
  }
  uint32_t lmax = 0;
  for (uint32_t i = 1; i < nex; i++) {
    uint32_t l = B->seg_end[s0 + i - 1] - B->seg_start[s0 + i - 1] + 1;
    if (l > lmax) lmax = l;
    int32_t jl = rj[i - 1];
    if (nj) (*nj)++;
    if (jl < 0 || !good) continue;
    uint32_t j = j0 + (uint32_t)jl;
    uint32_t anchor = P->junctionsupport;
    uint32_t jlen = B->j_end[j] - B->j_start[j] + 1;
    if (jlen > LONGINTRON ||
        (!P->longreads && B->j_nreads[j] - B->j_nm[j] < (double)P->junctionthr && !B->j_mm[j])) {
      if (anchor < LONGINTRONANCHOR) anchor = LONGINTRONANCHOR;
    }
    if (lmax >= anchor) {
      left[jl] += rdcount;
      if (rsup[i] >= anchor) { right[jl] += rdcount; good[jl] += rdcount; }
    } else if (rsup[i] >= anchor) {
      right[jl] += rdcount;
    }
  }
  if (rsup != stackbuf) free(rsup);
}

/* in-place: clamped running coverage, then prefix sums restarted every STG_BSIZE entries (rlink.cpp:17027-17063) */
static void blocked_scan(float* lane, int64_t len) {
  float prev_val = 0, prev_sum = 0;
  for (int64_t i = 0; i < len; i++) {
    if (i % STG_BSIZE == 0) prev_sum = 0;
    float v = lane[i];
    v += prev_val;
    if (v < 0) v = 0;
    prev_val = v;
    v += prev_sum;
    prev_sum = v;
    lane[i] = v;
  }
}

void orc_bundle(const stg_packed_batch* B, const orc_params* P, int64_t b, float* cov, double* good, double* left,
                double* right) {
  int64_t len = orc_bundle_len(B, b);
  int64_t refstart = B->b_start[b];
  uint32_t r0 = B->b_read_off[b], r1 = B->b_read_off[b + 1];
  if (cov) memset(cov, 0, sizeof(float) * 3 * (size_t)len);
  for (uint32_t n = r0; n < r1; n++) {
    if (cov && !(B->r_flags[n] & STG_RF_UNITIG)) read_to_cov(B, r0, n, refstart, cov, len, NULL);
    if (good) read_junction_support(B, P, B->b_junc_off[b], n, good, left, right, NULL);
  }
  if (cov)
    for (int s = 0; s < 3; s++) blocked_scan(cov + (int64_t)s * len, len);
}

void orc_batch_run(const stg_packed_batch* B, const orc_params* P, const int64_t* cov_off, float* cov, double* good,
                   double* left, double* right, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t b = 0; b < B->n_bundles; b++) {
    uint32_t j0 = B->b_junc_off[b];
    orc_bundle(B, P, b, cov ? cov + cov_off[b] : NULL, good ? good + j0 : NULL, left ? left + j0 : NULL,
               right ? right + j0 : NULL);
  }
}

double orc_get_cov(const float* cov, int64_t len, int s, uint32_t start, uint32_t end) {
  const float* c = cov + (int64_t)s * len;
  int m = (int)((end + 1) / STG_BSIZE);
  int k = (int)(start / STG_BSIZE);
  double sum = 0;
  for (int i = k + 1; i <= m; i++) sum += c[(int64_t)i * STG_BSIZE - 1];
  float edge = c[end + 1] - c[start];
  sum += edge;
  if (sum < 0) sum = 0;
  return sum;
}

double orc_get_cov_sign(const float* cov, int64_t len, int s, uint32_t start, uint32_t end) {
  const float* all = cov + len;
  const float* opp = cov + (int64_t)(2 - s) * len;
  int m = (int)((end + 1) / STG_BSIZE);
  int k = (int)(start / STG_BSIZE);
  double sum = 0;
  for (int i = k + 1; i <= m; i++) {
    float t = all[(int64_t)i * STG_BSIZE - 1] - opp[(int64_t)i * STG_BSIZE - 1];
    sum += t;
  }
  float edge = all[end + 1] - all[start];
  edge = edge - opp[end + 1];
  edge = edge + opp[start];
  sum += edge;
  if (sum < 0) sum = 0;
  return sum;
}

void orc_batch_stats(const stg_packed_batch* B, int64_t* L, int64_t* S, int64_t* J) {
  int64_t l = 0, s = 0, j = 0;
  orc_params P = {10, 1, 0};
  for (int64_t b = 0; b < B->n_bundles; b++) {
    l += orc_bundle_len(B, b);
    uint32_t r0 = B->b_read_off[b], r1 = B->b_read_off[b + 1];
    for (uint32_t n = r0; n < r1; n++) {
      if (!(B->r_flags[n] & STG_RF_UNITIG)) read_to_cov(B, r0, n, B->b_start[b], NULL, orc_bundle_len(B, b), &s);
      read_junction_support(B, &P, B->b_junc_off[b], n, NULL, NULL, NULL, &j);
    }
  }
  if (L) *L = l;
  if (S) *S = s;
  if (J) *J = j;
}
