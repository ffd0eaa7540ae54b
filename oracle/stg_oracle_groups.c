/* oracle/stg_oracle_groups.c — TEST INFRASTRUCTURE ONLY (see stg_oracle.h).
 *
 * CPU restatement of the read loop of build_graphs() (SURVEY.md §8 rows a8 + a9, first half), short-read de novo
 * path (no --viral, --mix, -L, guides):
 *   per-read junction repair, un-pairing, un-stranding        rlink.cpp:14845-15092
 *   add_read_to_group / merge_read_to_group / merge_fwd_groups rlink.cpp:1633-1727, 1281-1631, 1246-1278
 *   fragment counting                                          rlink.cpp:15105-15116
 *   re-sort of the junction table after appended rows          rlink.cpp:15122-15125
 *   merging of neighbouring groups (bundledist)                rlink.cpp:15134-15164
 *   colour pass over the three lanes (set_strandcol, neg_prop) rlink.cpp:15196-15299, 1028-1056, 984-1026
 *   bundles and bundlenodes (create_bundle, add_group_to_bundle) rlink.cpp:15340-15427, 1058-1095
 * The three are ONE sequential loop in the reference (repair read n -> group read n -> count it), and they share
 * state both ways (merge_read_to_group itself un-pairs reads, rlink.cpp:1383-1389), so they are restated together.
 * Pointers become indices: group g = slot in the group arrays (== CGroup::grid at creation), junction = row.
 * parity: pinned — tests/test_oracle_vs_reference.py compares with dumps taken INSIDE the reference's build_graphs
 * (hook inserted before the colour pass, rlink.cpp:15196, by oracle/Makefile).
 */
#include "stg_oracle.h"
#include "stg_oracle_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LONGINTRONANCHOR 25u /* rlink.h:32 */
static const double EPSILON = 0.000001; /* rlink.h:25 */

typedef struct {
  /* reads of the bundle (mutable copies) */
  int nr;
  uint32_t r0;                  /* first read of the bundle in the batch */
  const stg_packed_batch* B;
  int8_t* strand;               /* [nr] */
  uint32_t *seg_s, *seg_e;      /* batch-wide copies are too big: bundle slices, index = B->r_seg_off[n] - seg0 + i */
  uint32_t seg0, pair0, intr0;
  int32_t* pair;                /* bundle slice of pair_idx */
  int32_t* rj;                  /* bundle slice of rjunc_idx -> row in J */
  /* junctions */
  junc_t* J; int nJ, capJ, njunc0;
  int32_t* ej; int nE;
  /* groups */
  int ng, capg;
  uint32_t *g_start, *g_end; int32_t *g_color, *g_next; float *g_cov, *g_multi; int8_t* g_alive;
  int32_t* merge;               /* [ng] */
  int32_t* eqcol; int ncol, capcol;
  /* readgroup */
  int32_t** rg; int* rgn; int* rgcap;
  int32_t curr[3], startg[3];
  /* parameters */
  uint32_t junctionsupport, bundledist;
  int junctionthr;
  int64_t refstart, len;
  const float* cov;
} st_t;

#define NSEG(S, n) ((int)((S)->B->r_seg_off[(S)->r0 + (n) + 1] - (S)->B->r_seg_off[(S)->r0 + (n)]))
#define SEGI(S, n, i) ((S)->B->r_seg_off[(S)->r0 + (n)] - (S)->seg0 + (uint32_t)(i))
#define NJUNC(S, n) (NSEG(S, n) - 1)
#define RJI(S, n, i) ((S)->B->r_seg_off[(S)->r0 + (n)] - ((S)->r0 + (n)) - (S)->intr0 + (uint32_t)(i))
#define NPAIR(S, n) ((int)((S)->B->r_pair_off[(S)->r0 + (n) + 1] - (S)->B->r_pair_off[(S)->r0 + (n)]))
#define PAIRI(S, n, p) ((S)->B->r_pair_off[(S)->r0 + (n)] - (S)->pair0 + (uint32_t)(p))
#define RD_START(S, n) ((S)->B->r_start[(S)->r0 + (n)])
#define RD_END(S, n) ((S)->B->r_end[(S)->r0 + (n)])
#define RD_NH(S, n) ((S)->B->r_nh[(S)->r0 + (n)])
#define RD_LEN(S, n) ((S)->B->r_len[(S)->r0 + (n)])
#define RD_COUNT(S, n) ((S)->B->r_count[(S)->r0 + (n)])
#define RD_UNITIG(S, n) (((S)->B->r_flags[(S)->r0 + (n)] & STG_RF_UNITIG) != 0)
#define SLEN(S, n, i) ((S)->seg_e[SEGI(S, n, i)] - (S)->seg_s[SEGI(S, n, i)] + 1u)
/* strand of junction i of read n (CReadAln::juncs[i]->strand) */
#define JSTRAND(S, n, i) ((S)->J[(S)->rj[RJI(S, n, i)]].strand)

/* good_junc, mutating like the reference (rlink.cpp:13967): on failure the junction loses its strand */
static int good_junc_mut(st_t* S, int j) {
  int8_t s;
  int g = orc_good_junc(&S->J[j], S->refstart, S->cov, S->len, S->junctionthr, &s);
  S->J[j].strand = s;
  return g;
}

static int new_color(st_t* S, int* color) { /* eqcol.Add(color); color++ */
  if (S->ncol == S->capcol) { S->capcol = S->capcol * 2 + 64; S->eqcol = (int32_t*)realloc(S->eqcol, sizeof(int32_t) * (size_t)S->capcol); }
  S->eqcol[S->ncol++] = *color;
  return (*color)++;
}

static int new_group(st_t* S, uint32_t start, uint32_t end, int color, float cov, float multi) {
  if (S->ng == S->capg) {
    S->capg = S->capg * 2 + 64;
    S->g_start = (uint32_t*)realloc(S->g_start, 4 * (size_t)S->capg); S->g_end = (uint32_t*)realloc(S->g_end, 4 * (size_t)S->capg);
    S->g_color = (int32_t*)realloc(S->g_color, 4 * (size_t)S->capg); S->g_next = (int32_t*)realloc(S->g_next, 4 * (size_t)S->capg);
    S->g_cov = (float*)realloc(S->g_cov, 4 * (size_t)S->capg); S->g_multi = (float*)realloc(S->g_multi, 4 * (size_t)S->capg);
    S->g_alive = (int8_t*)realloc(S->g_alive, (size_t)S->capg); S->merge = (int32_t*)realloc(S->merge, 4 * (size_t)S->capg);
  }
  int g = S->ng++;
  S->g_start[g] = start; S->g_end[g] = end; S->g_color[g] = color; S->g_next[g] = -1;
  S->g_cov[g] = cov; S->g_multi[g] = multi; S->g_alive[g] = 1; S->merge[g] = g;
  return g;
}

static void rg_add(st_t* S, int n, int g) {
  if (S->rgn[n] == S->rgcap[n]) { S->rgcap[n] = S->rgcap[n] * 2 + 4; S->rg[n] = (int32_t*)realloc(S->rg[n], 4 * (size_t)S->rgcap[n]); }
  S->rg[n][S->rgn[n]++] = g;
}

/* rlink.cpp:1246-1278 */
static void merge_fwd_groups(st_t* S, int g1, int g2) {
  S->g_end[g1] = S->g_end[g2];
  while (S->eqcol[S->g_color[g1]] != S->g_color[g1]) S->g_color[g1] = S->eqcol[S->g_color[g1]];
  while (S->eqcol[S->g_color[g2]] != S->g_color[g2]) S->g_color[g2] = S->eqcol[S->g_color[g2]];
  if (S->g_color[g1] < S->g_color[g2]) S->eqcol[S->g_color[g2]] = S->g_color[g1];
  else if (S->g_color[g1] > S->g_color[g2]) { S->eqcol[S->g_color[g1]] = S->g_color[g2]; S->g_color[g1] = S->g_color[g2]; }
  S->g_cov[g1] += S->g_cov[g2];
  S->g_next[g1] = S->g_next[g2];
  S->merge[g2] = g1;            /* group ids are slots here: CGroup::grid of a live group is its slot (or its merge root) */
  S->g_multi[g1] += S->g_multi[g2];
  S->g_alive[g2] = 0;
}

static void unpair(st_t* S, int n, int p, int np) { /* readlist[n]->pair_idx[p]=-1 and the mirror entry of np */
  S->pair[PAIRI(S, n, p)] = -1;
  for (int j = 0; j < NPAIR(S, np); j++)
    if (S->pair[PAIRI(S, np, j)] == n) { S->pair[PAIRI(S, np, j)] = -1; break; }
}

/* CGroup::grid can be rewritten to its merge root (rlink.cpp:1358-1362); slots keep a separate grid field */
typedef struct { int32_t* grid; } grid_t;

/* rlink.cpp:1281-1631 */
static int merge_read_to_group(st_t* S, int32_t* grid, int n, int np, int p, float readcov, int sno, int readcol, int color,
                               int* usedcol) {
  uint32_t localdist = S->bundledist + LONGINTRONANCHOR; /* !longreads && !mixedMode */
  if (np > -1 && np < n && RD_END(S, np) < RD_START(S, n) && RD_END(S, np) + localdist > RD_START(S, n) &&
      (SLEN(S, np, NSEG(S, np) - 1) >= S->junctionsupport || (NJUNC(S, np) && JSTRAND(S, np, NJUNC(S, np) - 1))) &&
      (SLEN(S, n, 0) >= S->junctionsupport || (NJUNC(S, n) && JSTRAND(S, n, 0)))) {
    return color;
  }
  int first = 1;
  int currgroup = S->curr[sno];
  if (currgroup != -1) {
    int lastgroup = -1;
    while (currgroup != -1 && RD_START(S, n) > S->g_end[currgroup]) { lastgroup = currgroup; currgroup = S->g_next[currgroup]; }
    if (currgroup == -1 || S->seg_e[SEGI(S, n, 0)] < S->g_start[currgroup]) currgroup = lastgroup;
    int thisgroup = currgroup;
    int ncoord = NSEG(S, n);
    int lastpushedgroup = -1;
    int i = 0;
    while (i < ncoord) {
      int keep = 1;
      if (SLEN(S, n, i) < S->junctionsupport) { /* short reads: the longread clause never applies */
        if (i < NJUNC(S, n)) {
          if (!JSTRAND(S, n, i)) {
            if (!i || !JSTRAND(S, n, i - 1)) {
              keep = 0;
              if (!i && lastgroup != -1) currgroup = lastgroup;
            }
          }
        } else if (NJUNC(S, n)) {
          if (!JSTRAND(S, n, i - 1)) keep = 0;
        }
      }
      if (keep) {
        const uint32_t ss = S->seg_s[SEGI(S, n, i)], se = S->seg_e[SEGI(S, n, i)];
        while (thisgroup != -1 && ss > S->g_end[thisgroup]) { lastgroup = thisgroup; thisgroup = S->g_next[thisgroup]; }
        if (thisgroup != -1 && se + localdist >= S->g_start[thisgroup]) {
          if (!i) {
            int gr = grid[thisgroup];
            while (S->merge[gr] != gr) gr = S->merge[gr];
            grid[thisgroup] = gr;
            for (int g = 0; g < S->rgn[n]; g++) {
              gr = S->rg[n][g];
              while (S->merge[gr] != gr) gr = S->merge[gr];
              if (gr == grid[thisgroup]) { first = 0; break; }
            }
            if (np > -1 && RD_NH(S, np) && np < n) {
              int thiscol = S->g_color[thisgroup];
              while (S->eqcol[thiscol] != thiscol) thiscol = S->eqcol[thiscol];
              S->g_color[thisgroup] = thiscol;
              if (thiscol != readcol) {
                unpair(S, n, p, np);
                readcol = usedcol[sno];
                if (readcol < 0) { usedcol[sno] = color; readcol = new_color(S, &color); }
              }
            }
          }
          if (RD_NH(S, n) > 1) S->g_multi[thisgroup] += readcov * (float)(se - ss + 1u);
          if (ss < S->g_start[thisgroup]) S->g_start[thisgroup] = ss;
          int nextgroup = S->g_next[thisgroup];
          while (nextgroup != -1 && se >= S->g_start[nextgroup]) {
            merge_fwd_groups(S, thisgroup, nextgroup);
            nextgroup = S->g_next[thisgroup];
          }
          if (se > S->g_end[thisgroup]) S->g_end[thisgroup] = se;
          while (S->eqcol[S->g_color[thisgroup]] != S->g_color[thisgroup]) S->g_color[thisgroup] = S->eqcol[S->g_color[thisgroup]];
          if (readcol != S->g_color[thisgroup]) {
            if (i && !JSTRAND(S, n, i - 1)) readcol = S->g_color[thisgroup];
            else {
              if (readcol < S->g_color[thisgroup]) { S->eqcol[S->g_color[thisgroup]] = readcol; S->g_color[thisgroup] = readcol; }
              else { S->eqcol[readcol] = S->g_color[thisgroup]; readcol = S->g_color[thisgroup]; }
            }
          }
          if (grid[thisgroup] != lastpushedgroup) {
            if (first) rg_add(S, n, grid[thisgroup]);
            lastpushedgroup = grid[thisgroup];
          }
          S->g_cov[thisgroup] += (float)(se - ss + 1u) * readcov;
        } else {
          if (!i && np > -1 && RD_NH(S, np) && np < n) {
            unpair(S, n, p, np);
            readcol = usedcol[sno];
            if (readcol < 0) { usedcol[sno] = color; readcol = new_color(S, &color); }
          } else if (i && !JSTRAND(S, n, i - 1)) {
            usedcol[sno] = color; readcol = new_color(S, &color);
          }
          float multi = 0;
          if (RD_NH(S, n) > 1) multi = readcov * (float)(se - ss + 1u);
          int ngroup = new_group(S, ss, se, readcol, (float)(se - ss + 1u) * readcov, multi);
          grid[ngroup] = ngroup;
          if (lastgroup != -1) S->g_next[lastgroup] = ngroup;
          S->g_next[ngroup] = thisgroup;
          rg_add(S, n, ngroup);
          lastpushedgroup = ngroup;
          thisgroup = ngroup;
        }
        if (i == ncoord - 1) {
          if (np != -1 && np > n && RD_END(S, n) < RD_START(S, np) && RD_END(S, n) + localdist > RD_START(S, np) &&
              (SLEN(S, np, 0) >= S->junctionsupport || (NJUNC(S, np) && JSTRAND(S, np, 0)))) {
            n = np; np = -1;
            ncoord = NSEG(S, n);
            first = 1;
            if (RD_START(S, n) > S->g_end[thisgroup]) {
              int nextgroup = S->g_next[thisgroup];
              while (nextgroup != -1 && RD_START(S, n) >= S->g_start[nextgroup]) {
                merge_fwd_groups(S, thisgroup, nextgroup);
                nextgroup = S->g_next[thisgroup];
              }
              if (RD_START(S, n) > S->g_end[thisgroup]) S->g_end[thisgroup] = RD_START(S, n);
            }
            lastpushedgroup = -1;
            i = -1;
          }
        }
      }
      i++;
    }
  } else {
    if (np > -1 && RD_NH(S, np) && np < n) {
      unpair(S, n, p, np);
      readcol = usedcol[sno];
      if (readcol < 0) { usedcol[sno] = color; readcol = new_color(S, &color); }
    }
    int ncoord = NSEG(S, n);
    int lastgroup = -1;
    float multi = 0;
    if (RD_NH(S, n) > 1) multi = readcov;
    int i = 0;
    while (i < ncoord) {
      int keep = 1;
      if (SLEN(S, n, i) < S->junctionsupport) {
        if (i < NJUNC(S, n)) {
          if (!JSTRAND(S, n, i)) { if (!i || !JSTRAND(S, n, i - 1)) keep = 0; }
        } else if (NJUNC(S, n)) {
          if (!JSTRAND(S, n, i - 1)) keep = 0;
        }
      }
      if (!i || keep) {
        if (i && !JSTRAND(S, n, i - 1)) { usedcol[sno] = color; readcol = new_color(S, &color); }
        const uint32_t ss = S->seg_s[SEGI(S, n, i)], se = S->seg_e[SEGI(S, n, i)];
        const float flen = (float)(se - ss + 1u);
        int ngroup = new_group(S, ss, se, readcol, flen * readcov, flen * multi);
        grid[ngroup] = ngroup;
        if (lastgroup != -1) S->g_next[lastgroup] = ngroup;
        else currgroup = ngroup;
        lastgroup = ngroup;
        if (keep) {
          rg_add(S, n, ngroup);
          if (i == ncoord - 1) {
            if (np != -1 && np > n && RD_END(S, n) < RD_START(S, np) && RD_END(S, n) + localdist > RD_START(S, np) &&
                (SLEN(S, np, 0) >= S->junctionsupport || (NJUNC(S, np) && JSTRAND(S, np, 0)))) {
              n = np; np = -1;
              ncoord = NSEG(S, n);
              if (S->seg_e[SEGI(S, n, 0)] > S->g_end[lastgroup]) {
                S->g_end[lastgroup] = S->seg_e[SEGI(S, n, 0)];
                if (first) rg_add(S, n, ngroup);
                S->g_cov[lastgroup] += (float)SLEN(S, n, 0) * readcov;
              }
              i = 0;
            }
          }
        }
      }
      i++;
    }
  }
  S->curr[sno] = currgroup;
  if (S->startg[sno] == -1) S->startg[sno] = currgroup;
  return color;
}

/* rlink.cpp:1633-1727 */
static int add_read_to_group(st_t* S, int32_t* grid, int n, int color) {
  int usedcol[3] = {-1, -1, -1};
  float single_count = RD_COUNT(S, n);
  for (int p = 0; p < NPAIR(S, n); p++) {
    int sno = (int)S->strand[n] + 1;
    int np = S->pair[PAIRI(S, n, p)];
    if (np > -1 && RD_NH(S, np)) {
      int snop = (int)S->strand[np] + 1;
      if (sno != snop) {
        if (sno == 1) sno = snop;
        else if (snop != 1) { unpair(S, n, p, np); np = -1; }
      }
      if (np > -1) {
        single_count -= S->B->pair_cnt[S->pair0 + PAIRI(S, n, p)];
        int readcol = usedcol[sno];
        if (np < n) {
          int grouppair = S->rg[np][0];
          while (S->merge[grouppair] != grouppair) grouppair = S->merge[grouppair];
          S->rg[np][0] = grouppair;
          readcol = S->g_color[grouppair];
          while (S->eqcol[readcol] != readcol) readcol = S->eqcol[readcol];
          S->g_color[grouppair] = readcol;
        } else {
          if (usedcol[sno] < 0) { usedcol[sno] = color; readcol = new_color(S, &color); }
        }
        double readcov = 0;
        if (!RD_UNITIG(S, n)) readcov = S->B->pair_cnt[S->pair0 + PAIRI(S, n, p)];
        color = merge_read_to_group(S, grid, n, np, p, (float)readcov, sno, readcol, color, usedcol);
      }
    }
  }
  if (single_count > EPSILON) {
    int sn = (int)S->strand[n] + 1;
    int readcol = usedcol[sn];
    if (readcol < 0) { usedcol[sn] = color; readcol = new_color(S, &color); }
    double readcov = 0;
    if (!RD_UNITIG(S, n)) readcov = single_count;
    color = merge_read_to_group(S, grid, n, -1, -1, (float)readcov, sn, readcol, color, usedcol);
  }
  return color;
}

static int junc_lt(const junc_t* a, const junc_t* b) { /* CJunction::operator<, bundle.h:155-162 */
  if (a->start < b->start) return 1;
  if (a->start > b->start) return 0;
  if (a->end < b->end) return 1;
  if (a->end > b->end) return 0;
  return a->strand > b->strand;
}
static int junc_default_cmp(const junc_t* a, const junc_t* b) { /* GList::DefaultCompareProc, gclib/GList.hh:85-90 */
  if (junc_lt(b, a)) return 1;
  else if (junc_lt(a, b)) return -1;
  return 0;
}


/* ---- second half of a9: colours across the three lanes, then bundles ---------------------------------------- */
static int get_min_start(const st_t* S, const int* cur) { /* rlink.cpp:984-1026 */
  if (cur[0] != -1) {
    if (cur[1] != -1) {
      if (cur[2] != -1) {
        int two = S->g_start[cur[0]] < S->g_start[cur[1]] ? 0 : 1;
        return S->g_start[cur[two]] < S->g_start[cur[2]] ? two : 2;
      }
      return S->g_start[cur[0]] < S->g_start[cur[1]] ? 0 : 1;
    }
    if (cur[2] != -1) return S->g_start[cur[0]] < S->g_start[cur[2]] ? 0 : 2;
    return 0;
  }
  if (cur[1] != -1) {
    if (cur[2] != -1) return S->g_start[cur[1]] < S->g_start[cur[2]] ? 1 : 2;
    return 1;
  }
  return 2;
}

/* rlink.cpp:1028-1056: U = the unstranded group, G = the stranded one, grcol = G's colour */
static void set_strandcol(int32_t* color, int U, int G, int grcol, int32_t* eqcol, int32_t* equalcolor) {
  int zerocol = eqcol[color[U]];
  if (zerocol > -1) {
    while (eqcol[zerocol] != -1 && eqcol[zerocol] != zerocol) zerocol = eqcol[zerocol];
    int tmpcol = zerocol;
    while (equalcolor[zerocol] != zerocol) zerocol = equalcolor[zerocol];
    eqcol[color[U]] = zerocol;
    eqcol[tmpcol] = zerocol;
    if (zerocol < grcol) { equalcolor[grcol] = zerocol; color[G] = zerocol; }
    else if (grcol < zerocol) { equalcolor[zerocol] = grcol; eqcol[color[U]] = grcol; }
  } else eqcol[color[U]] = grcol;
}

typedef struct { int n, cap; int32_t *len, *start, *last; float *cov, *multi; } bun_t;
typedef struct { int n, cap; uint32_t *s, *e; float* cov; int32_t* next; } bn_t;

static int bn_add(bn_t* N, uint32_t s, uint32_t e, float cov) {
  if (N->n == N->cap) {
    N->cap = N->cap * 2 + 16;
    N->s = (uint32_t*)realloc(N->s, 4 * (size_t)N->cap); N->e = (uint32_t*)realloc(N->e, 4 * (size_t)N->cap);
    N->cov = (float*)realloc(N->cov, 4 * (size_t)N->cap); N->next = (int32_t*)realloc(N->next, 4 * (size_t)N->cap);
  }
  N->s[N->n] = s; N->e[N->n] = e; N->cov[N->n] = cov; N->next[N->n] = -1;
  return N->n++;
}

static int create_bundle(const st_t* S, bun_t* Bn, int g, bn_t* N) { /* rlink.cpp:1083-1095 */
  int bid = bn_add(N, S->g_start[g], S->g_end[g], S->g_cov[g]);
  if (Bn->n == Bn->cap) {
    Bn->cap = Bn->cap * 2 + 16;
    Bn->len = (int32_t*)realloc(Bn->len, 4 * (size_t)Bn->cap); Bn->start = (int32_t*)realloc(Bn->start, 4 * (size_t)Bn->cap);
    Bn->last = (int32_t*)realloc(Bn->last, 4 * (size_t)Bn->cap); Bn->cov = (float*)realloc(Bn->cov, 4 * (size_t)Bn->cap);
    Bn->multi = (float*)realloc(Bn->multi, 4 * (size_t)Bn->cap);
  }
  int bno = Bn->n++;
  Bn->len[bno] = (int32_t)(S->g_end[g] - S->g_start[g] + 1); Bn->cov[bno] = S->g_cov[g]; Bn->multi[bno] = S->g_multi[g];
  Bn->start[bno] = bid; Bn->last[bno] = bid;
  return bno;
}

static void add_group_to_bundle(const st_t* S, int g, bun_t* Bn, int bno, bn_t* N, uint32_t localdist) { /* rlink.cpp:1058-1081 */
  int last = Bn->last[bno];
  if (S->g_start[g] > N->e[last] + localdist) {
    int bid = bn_add(N, S->g_start[g], S->g_end[g], S->g_cov[g]);
    N->next[last] = bid;
    Bn->last[bno] = bid;
    Bn->len[bno] += (int32_t)(S->g_end[g] - S->g_start[g] + 1);
    Bn->cov[bno] += S->g_cov[g];
    Bn->multi[bno] += S->g_multi[g];
  } else {
    if (N->e[last] < S->g_end[g]) { Bn->len[bno] += (int32_t)(S->g_end[g] - N->e[last]); N->e[last] = S->g_end[g]; }
    Bn->cov[bno] += S->g_cov[g];
    Bn->multi[bno] += S->g_multi[g];
    N->cov[last] += S->g_cov[g];
  }
}

static void second_half(st_t* S, orc_groups_out* o) {
  const int ng = S->ng, nc = S->ncol;
  const uint32_t bd = S->bundledist;
  int32_t* color = (int32_t*)malloc(4 * ((size_t)ng + 1)); memcpy(color, S->g_color, 4 * (size_t)ng);
  int32_t* eq = (int32_t*)malloc(4 * ((size_t)nc + 1)); memcpy(eq, S->eqcol, 4 * (size_t)nc);
  int32_t* eqneg = (int32_t*)malloc(4 * ((size_t)nc + 1)); int32_t* eqpos = (int32_t*)malloc(4 * ((size_t)nc + 1));
  for (int c = 0; c < nc; c++) { eqneg[c] = -1; eqpos[c] = -1; }
  float* negp = (float*)calloc((size_t)ng + 1, sizeof(float));
#define GLEN(g) (S->g_end[g] - S->g_start[g] + 1u)
#define OVL(U, G, acc) do { \
    uint32_t maxstart = S->g_start[U] > S->g_start[G] ? S->g_start[U] : S->g_start[G]; \
    uint32_t minend = S->g_end[U] < S->g_end[G] ? S->g_end[U] : S->g_end[G]; \
    if (minend < maxstart) minend = maxstart; \
    acc += S->g_cov[G] * (float)(minend - maxstart + 1u) / (float)GLEN(G); } while (0)
  int cur[3], prev[3] = {-1, -1, -1};
  for (int i = 0; i < 3; i++) cur[i] = S->startg[i];
  while (cur[0] != -1 || cur[1] != -1 || cur[2] != -1) {
    const int nextgr = get_min_start(S, cur);
    const int g = cur[nextgr];
    int grcol = color[g];
    while (eq[grcol] != grcol) grcol = eq[grcol];
    color[g] = grcol;
    if (nextgr == 1) {
      if (prev[0] != -1 && S->g_start[g] <= S->g_end[prev[0]] + bd) {
        set_strandcol(color, g, prev[0], color[prev[0]], eqneg, eq);
        OVL(g, prev[0], negp[g]);
      }
      while (cur[0] != -1 && S->g_start[g] <= S->g_end[cur[0]] + bd && S->g_start[cur[0]] <= S->g_end[g] + bd) {
        int c0 = color[cur[0]];
        while (eq[c0] != c0) c0 = eq[c0];
        color[cur[0]] = c0;
        negp[cur[0]] = 1;
        set_strandcol(color, g, cur[0], color[cur[0]], eqneg, eq);
        OVL(g, cur[0], negp[g]);
        prev[0] = cur[0];
        cur[0] = S->g_next[cur[0]];
      }
      float pos_prop = 0;
      if (prev[2] != -1 && S->g_start[g] <= S->g_end[prev[2]] + bd) {
        set_strandcol(color, g, prev[2], color[prev[2]], eqpos, eq);
        if (negp[g]) OVL(g, prev[2], pos_prop);
      }
      while (cur[2] != -1 && S->g_start[g] <= S->g_end[cur[2]] + bd && S->g_start[cur[2]] <= S->g_end[g] + bd) {
        int c2 = color[cur[2]];
        while (eq[c2] != c2) c2 = eq[c2];
        color[cur[2]] = c2;
        set_strandcol(color, g, cur[2], color[cur[2]], eqpos, eq);
        if (negp[g]) OVL(g, cur[2], pos_prop);
        prev[2] = cur[2];
        cur[2] = S->g_next[cur[2]];
      }
      if (pos_prop) negp[g] /= (negp[g] + pos_prop);
      else if (negp[g]) negp[g] = 1;
    } else if (nextgr == 0) negp[g] = 1;
    prev[nextgr] = g;
    cur[nextgr] = S->g_next[g];
  }

  bun_t Bn[3]; bn_t N[3];
  memset(Bn, 0, sizeof(Bn)); memset(N, 0, sizeof(N));
  int32_t* g2b = (int32_t*)malloc(4 * (3 * (size_t)ng + 1));
  for (int k = 0; k < 3 * ng; k++) g2b[k] = -1;
  int32_t* bundlecol = (int32_t*)malloc(4 * ((size_t)nc + 1));
  for (int c = 0; c < nc; c++) bundlecol[c] = -1;
  for (int i = 0; i < 3; i++) cur[i] = S->startg[i];
  while (cur[0] != -1 || cur[1] != -1 || cur[2] != -1) {
    const int nextgr = get_min_start(S, cur);
    const int g = cur[nextgr];
    int grcol = color[g];
    while (eq[grcol] != grcol) grcol = eq[grcol];
    color[g] = grcol;
    if (nextgr == 0 || nextgr == 2 || (nextgr == 1 && eqneg[grcol] == -1 && eqpos[grcol] == -1)) {
      int bno = bundlecol[grcol];
      if (bno > -1) add_group_to_bundle(S, g, &Bn[nextgr], bno, &N[nextgr], bd);
      else { bno = create_bundle(S, &Bn[nextgr], g, &N[nextgr]); bundlecol[grcol] = bno; }
      g2b[nextgr * ng + g] = Bn[nextgr].last[bno];
    } else {
      if (eqneg[grcol] != -1) {
        int negcol = eqneg[grcol];
        while (eq[negcol] != negcol) negcol = eq[negcol];
        int bno = bundlecol[negcol];
        if (bno > -1) add_group_to_bundle(S, g, &Bn[0], bno, &N[0], bd);
        else { bno = create_bundle(S, &Bn[0], g, &N[0]); bundlecol[negcol] = bno; }
        g2b[0 * ng + g] = Bn[0].last[bno];
      }
      if (eqpos[grcol] != -1) {
        int poscol = eqpos[grcol];
        while (eq[poscol] != poscol) poscol = eq[poscol];
        int bno = bundlecol[poscol];
        if (bno > -1) add_group_to_bundle(S, g, &Bn[2], bno, &N[2], bd);
        else { bno = create_bundle(S, &Bn[2], g, &N[2]); bundlecol[poscol] = bno; }
        g2b[2 * ng + g] = Bn[2].last[bno];
      }
    }
    cur[nextgr] = S->g_next[g];
  }
#undef OVL
#undef GLEN
  o->g_color2 = color; o->g_neg_prop = negp; o->eqcol2 = eq; o->eqnegcol = eqneg; o->eqposcol = eqpos; o->group2bundle = g2b;
  int nbun = Bn[0].n + Bn[1].n + Bn[2].n, nbn = N[0].n + N[1].n + N[2].n;
  o->bun_len = (int32_t*)malloc(4 * ((size_t)nbun + 1)); o->bun_start = (int32_t*)malloc(4 * ((size_t)nbun + 1));
  o->bun_last = (int32_t*)malloc(4 * ((size_t)nbun + 1)); o->bun_cov = (float*)malloc(4 * ((size_t)nbun + 1));
  o->bun_multi = (float*)malloc(4 * ((size_t)nbun + 1));
  o->bn_start = (uint32_t*)malloc(4 * ((size_t)nbn + 1)); o->bn_end = (uint32_t*)malloc(4 * ((size_t)nbn + 1));
  o->bn_cov = (float*)malloc(4 * ((size_t)nbn + 1)); o->bn_next = (int32_t*)malloc(4 * ((size_t)nbn + 1));
  int ob = 0, on = 0;
  for (int s = 0; s < 3; s++) {
    o->bun_off[s] = ob; o->bn_off[s] = on;
    for (int k = 0; k < Bn[s].n; k++, ob++) {
      o->bun_len[ob] = Bn[s].len[k]; o->bun_start[ob] = Bn[s].start[k]; o->bun_last[ob] = Bn[s].last[k];
      o->bun_cov[ob] = Bn[s].cov[k]; o->bun_multi[ob] = Bn[s].multi[k];
    }
    for (int k = 0; k < N[s].n; k++, on++) {
      o->bn_start[on] = N[s].s[k]; o->bn_end[on] = N[s].e[k]; o->bn_cov[on] = N[s].cov[k]; o->bn_next[on] = N[s].next[k];
    }
    free(Bn[s].len); free(Bn[s].start); free(Bn[s].last); free(Bn[s].cov); free(Bn[s].multi);
    free(N[s].s); free(N[s].e); free(N[s].cov); free(N[s].next);
  }
  o->bun_off[3] = ob; o->bn_off[3] = on;
  free(bundlecol);
}

void orc_groups_free(orc_groups_out* o) {
  free(o->r_strand); free(o->pair_idx); free(o->seg_start); free(o->seg_end); free(o->rjunc);
  free(o->j_start); free(o->j_end); free(o->j_strand); free(o->j_nreads); free(o->j_nreads_good); free(o->j_left);
  free(o->j_right); free(o->j_nm); free(o->j_mm);
  free(o->g_start); free(o->g_end); free(o->g_color); free(o->g_next); free(o->g_cov_sum); free(o->g_multi);
  free(o->g_alive); free(o->merge); free(o->eqcol); free(o->rg_off); free(o->rg);
  free(o->g_color2); free(o->g_neg_prop); free(o->eqcol2); free(o->eqnegcol); free(o->eqposcol);
  free(o->bun_len); free(o->bun_start); free(o->bun_last); free(o->bun_cov); free(o->bun_multi);
  free(o->bn_start); free(o->bn_end); free(o->bn_cov); free(o->bn_next); free(o->group2bundle);
  memset(o, 0, sizeof(*o));
}

void orc_groups_bundle(const stg_packed_batch* B, const orc_params* P, int64_t b, const float* cov,
                       const double* good_in, const double* left_in, const double* right_in, uint32_t bundledist,
                       orc_groups_out* o) {
  memset(o, 0, sizeof(*o));
  st_t S;
  memset(&S, 0, sizeof(S));
  S.B = B; S.r0 = B->b_read_off[b]; S.nr = (int)(B->b_read_off[b + 1] - S.r0);
  S.junctionsupport = P->junctionsupport; S.junctionthr = P->junctionthr; S.bundledist = bundledist;
  S.refstart = B->b_start[b]; S.len = orc_bundle_len(B, b); S.cov = cov;
  const int nr = S.nr;
  S.seg0 = B->r_seg_off[S.r0]; S.pair0 = B->r_pair_off[S.r0]; S.intr0 = S.seg0 - S.r0;
  const uint32_t nseg = B->r_seg_off[S.r0 + nr] - S.seg0, npair = B->r_pair_off[S.r0 + nr] - S.pair0;
  const uint32_t nintr = nseg - (uint32_t)nr;
  /* junction pass first; spare rows for the junctions the repair may append (at most one per intron) */
  orc_junction_state(B, P, b, cov, good_in, left_in, right_in, (int)nintr + 1, &S.J, &S.ej, &S.nJ);
  if (!S.J) { S.J = (junc_t*)calloc(nintr + 1, sizeof(junc_t)); S.ej = (int32_t*)calloc(nintr + 1, sizeof(int32_t)); }
  S.nE = S.nJ; S.njunc0 = S.nJ;
  S.strand = (int8_t*)malloc((size_t)nr + 1); memcpy(S.strand, B->r_strand + S.r0, (size_t)nr);
  S.seg_s = (uint32_t*)malloc(4 * ((size_t)nseg + 1)); memcpy(S.seg_s, B->seg_start + S.seg0, 4 * (size_t)nseg);
  S.seg_e = (uint32_t*)malloc(4 * ((size_t)nseg + 1)); memcpy(S.seg_e, B->seg_end + S.seg0, 4 * (size_t)nseg);
  S.pair = (int32_t*)malloc(4 * ((size_t)npair + 1)); memcpy(S.pair, B->pair_idx + S.pair0, 4 * (size_t)npair);
  S.rj = (int32_t*)malloc(4 * ((size_t)nintr + 1)); memcpy(S.rj, B->rjunc_idx + S.intr0, 4 * (size_t)nintr);
  for (uint32_t k = 0; k < nintr; k++) if (S.rj[k] < 0) S.rj[k] = 0;   /* as ref_hot's build_bundle: no row -> the sentinel */
  S.rg = (int32_t**)calloc((size_t)nr + 1, sizeof(int32_t*)); S.rgn = (int*)calloc((size_t)nr + 1, sizeof(int));
  S.rgcap = (int*)calloc((size_t)nr + 1, sizeof(int));
  S.curr[0] = S.curr[1] = S.curr[2] = -1; S.startg[0] = S.startg[1] = S.startg[2] = -1;
  int32_t* grid = NULL; int gridcap = 0;
  int color = 0;
  double num_fragments = 0, frag_len = 0;
#define EJ(k) S.J[S.ej[k]]

  for (int n = 0; n < nr; n++) {
    int keep = 1;
    int i = 0;
    const int nj = NJUNC(&S, n);
    while (i < nj) {
      const int jd = S.rj[RJI(&S, n, i)];
      junc_t* J = S.J;
      const int changeright = J[jd].nreads_good < 0;
      const int changeleft = J[jd].nreads < 0;
      if (J[jd].strand) {
        if (!good_junc_mut(&S, jd)) J[jd].mm = -1;
        if (changeleft || changeright) J[jd].strand = 0;
        if (J[jd].mm < 0) J[jd].strand = 0;
      }
      if (!J[jd].strand) {
        if (changeleft || changeright) {
          uint32_t newstart = S.seg_e[SEGI(&S, n, i)];
          uint32_t newend = S.seg_s[SEGI(&S, n, i + 1)];
          int jk = -1, ek = -1;
          if (changeleft) {
            jk = abs((int)J[jd].nreads);
            if (J[jk].nreads < 0) newstart = S.seg_s[SEGI(&S, n, i)] - 1;
            else { newstart = J[jk].start; if (J[jk].strand) jk = -1; }
          }
          if (changeright) {
            ek = abs((int)J[jd].nreads_good);
            if (EJ(ek).nreads_good < 0) newend = S.seg_e[SEGI(&S, n, i + 1)] + 1;
            else { newend = EJ(ek).end; if (EJ(ek).strand) ek = -1; }
          }
          if (newstart >= S.seg_s[SEGI(&S, n, i)] && newend <= S.seg_e[SEGI(&S, n, i + 1)] && newstart <= newend) {
            int searchjunc = 1;
            S.seg_e[SEGI(&S, n, i)] = newstart;
            S.seg_s[SEGI(&S, n, i + 1)] = newend;
            if (J[jd].mm >= 0) {
              const int8_t rs = S.strand[n];
              if (jk > 0) {
                int k = jk;
                while (k > 0 && J[k].start == newstart) {
                  if (rs == J[k].strand && J[k].end == newend) { S.rj[RJI(&S, n, i)] = k; searchjunc = 0; break; }
                  k--;
                }
                if (searchjunc) {
                  k = jk + 1;
                  while (k < S.nJ && J[k].start == newstart) {
                    if (rs == J[k].strand && J[k].end == newend) { S.rj[RJI(&S, n, i)] = k; searchjunc = 0; break; }
                    k++;
                  }
                }
                if (searchjunc) {
                  for (k = S.njunc0; k < S.nJ; k++)
                    if (rs == J[k].strand && J[k].start == newstart && J[k].end == newend) { S.rj[RJI(&S, n, i)] = k; searchjunc = 0; break; }
                  if (searchjunc) {
                    junc_t nj2; memset(&nj2, 0, sizeof(nj2));
                    nj2.start = newstart; nj2.end = newend; nj2.strand = rs; nj2.consleft = -1; nj2.consright = -1;
                    J[S.nJ] = nj2; S.ej[S.nE++] = S.nJ; S.rj[RJI(&S, n, i)] = S.nJ; S.nJ++;
                  }
                }
              } else if (ek > 0) {
                int k = ek;
                while (k > 0 && EJ(k).end == newend) {
                  if (rs == EJ(k).strand && EJ(k).start == newstart) { S.rj[RJI(&S, n, i)] = S.ej[k]; searchjunc = 0; break; }
                  k--;
                }
                if (searchjunc) {
                  k = ek + 1;
                  while (k < S.nE && EJ(k).end == newend) {
                    if (rs == EJ(k).strand && EJ(k).start == newstart) { S.rj[RJI(&S, n, i)] = S.ej[k]; searchjunc = 0; break; }
                    k++;
                  }
                }
                if (searchjunc) {
                  for (k = S.njunc0; k < S.nE; k++)
                    if (rs == EJ(k).strand && EJ(k).start == newstart && EJ(k).end == newend) { S.rj[RJI(&S, n, i)] = S.ej[k]; searchjunc = 0; break; }
                  if (searchjunc) {
                    junc_t nj2; memset(&nj2, 0, sizeof(nj2));
                    nj2.start = newstart; nj2.end = newend; nj2.strand = rs; nj2.consleft = -1; nj2.consright = -1;
                    J[S.nJ] = nj2; S.ej[S.nE++] = S.nJ; S.rj[RJI(&S, n, i)] = S.nJ; S.nJ++;
                  }
                }
              }
            }
          }
        }
        if (keep) {
          for (int p = 0; p < NPAIR(&S, n); p++) {
            const int np = S.pair[PAIRI(&S, n, p)];
            if (np > -1) {
              S.pair[PAIRI(&S, n, p)] = -1;
              for (int j = 0; j < NPAIR(&S, np); j++)
                if (S.pair[PAIRI(&S, np, j)] == n) { S.pair[PAIRI(&S, np, j)] = -1; break; }
              if (!NJUNC(&S, np)) S.strand[np] = 0;
            }
          }
          keep = 0;
        }
      }
      i++;
    }
    if (!keep) {
      if (S.strand[n]) {
        int keepstrand = 0;
        for (int j = 0; j < nj; j++) if (JSTRAND(&S, n, j)) { keepstrand = 1; break; }
        if (!keepstrand) S.strand[n] = 0;
      }
    } else if (!nj && S.strand[n]) {
      if (NPAIR(&S, n)) {
        int keepstrand = 0;
        for (int p = 0; p < NPAIR(&S, n); p++) {
          const int np = S.pair[PAIRI(&S, n, p)];
          if (np > -1 && n < np) {
            if (!NJUNC(&S, np)) {
              if (S.strand[np] == S.strand[n]) { keepstrand = 1; break; }
            } else {
              for (int j = 0; j < NJUNC(&S, np); j++) {
                const int jx = S.rj[RJI(&S, np, j)];
                if (S.J[jx].strand && good_junc_mut(&S, jx)) { keepstrand = 1; break; }
              }
              if (!keepstrand) S.strand[np] = 0;
            }
          }
          if (keepstrand) break;
        }
        if (!keepstrand) S.strand[n] = 0;
      }
    }

    /* groups: grid[] must cover every slot new_group may hand out while this read is processed */
    {
      const int need = S.ng + 2 * nseg + 16;
      if (need > gridcap) { gridcap = need * 2; grid = (int32_t*)realloc(grid, 4 * (size_t)gridcap); }
    }
    color = add_read_to_group(&S, grid, n, color);

    /* fragment counting: rlink.cpp:15105-15116 */
    if (!RD_UNITIG(&S, n)) frag_len += (double)((float)RD_LEN(&S, n) * RD_COUNT(&S, n));
    double single_count = RD_COUNT(&S, n);
    if (keep)
      for (int p = 0; p < NPAIR(&S, n); p++) {
        const int pi = S.pair[PAIRI(&S, n, p)];
        if (pi != -1 && n > pi && RD_NH(&S, pi)) single_count -= B->pair_cnt[S.pair0 + PAIRI(&S, n, p)];
      }
    if (!RD_UNITIG(&S, n) && single_count > EPSILON) num_fragments += single_count;
  }

  /* merge groups that are close together: rlink.cpp:15134-15164 (no guides: boundaryleft/right are empty) */
  if (S.bundledist) {
    for (int sno = 0; sno < 3; sno++) {
      int lastgroup = -1, procgroup = S.startg[sno];
      while (procgroup != -1) {
        if (lastgroup != -1) {
          if (S.g_start[procgroup] - S.g_end[lastgroup] <= S.bundledist) {
            merge_fwd_groups(&S, lastgroup, procgroup);
            procgroup = S.g_next[lastgroup];
            continue;
          }
        }
        lastgroup = procgroup;
        procgroup = S.g_next[procgroup];
      }
    }
  }

  /* junction table: appended rows force a re-sort (rlink.cpp:15122-15125: setSorted(true) -> GList::Sort with
   * DefaultCompareProc, gclib/GList.hh:85-90, 110-118).  Rows whose strand was zeroed meanwhile have CHANGED keys and
   * may now tie with other rows, so the reference's own (non-stable) quicksort decides the order. */
  int32_t* perm = (int32_t*)malloc(4 * ((size_t)S.nJ + 1));
  int32_t* inv = (int32_t*)malloc(4 * ((size_t)S.nJ + 1));
  for (int j = 0; j < S.nJ; j++) perm[j] = j;
  if (S.nJ > S.njunc0) orc_qsort_ref(perm, S.J, 0, S.nJ - 1, junc_default_cmp);
  for (int j = 0; j < S.nJ; j++) inv[perm[j]] = j;

  /* export */
  o->n_reads = nr; o->n_segs = (int32_t)nseg; o->n_pairs = (int32_t)npair; o->n_juncs = S.nJ; o->n_groups = S.ng; o->n_colors = S.ncol;
  o->r_strand = S.strand; o->pair_idx = S.pair; o->seg_start = S.seg_s; o->seg_end = S.seg_e;
  o->rjunc = S.rj;
  for (uint32_t k = 0; k < nintr; k++) o->rjunc[k] = inv[o->rjunc[k]];
  o->j_start = (uint32_t*)malloc(4 * ((size_t)S.nJ + 1)); o->j_end = (uint32_t*)malloc(4 * ((size_t)S.nJ + 1));
  o->j_strand = (int8_t*)malloc((size_t)S.nJ + 1);
  o->j_nreads = (double*)malloc(8 * ((size_t)S.nJ + 1)); o->j_nreads_good = (double*)malloc(8 * ((size_t)S.nJ + 1));
  o->j_left = (double*)malloc(8 * ((size_t)S.nJ + 1)); o->j_right = (double*)malloc(8 * ((size_t)S.nJ + 1));
  o->j_nm = (double*)malloc(8 * ((size_t)S.nJ + 1)); o->j_mm = (double*)malloc(8 * ((size_t)S.nJ + 1));
  for (int j = 0; j < S.nJ; j++) {
    const junc_t* x = &S.J[perm[j]];
    o->j_start[j] = x->start; o->j_end[j] = x->end; o->j_strand[j] = x->strand; o->j_nreads[j] = x->nreads;
    o->j_nreads_good[j] = x->nreads_good; o->j_left[j] = x->leftsupport; o->j_right[j] = x->rightsupport;
    o->j_nm[j] = x->nm; o->j_mm[j] = x->mm;
  }
  o->g_start = S.g_start; o->g_end = S.g_end; o->g_color = S.g_color; o->g_next = S.g_next; o->g_cov_sum = S.g_cov;
  o->g_multi = S.g_multi; o->g_alive = S.g_alive; o->merge = S.merge; o->eqcol = S.eqcol;
  for (int s = 0; s < 3; s++) o->startgroup[s] = S.startg[s];
  o->rg_off = (uint32_t*)malloc(4 * ((size_t)nr + 1));
  uint32_t tot = 0;
  for (int n = 0; n < nr; n++) { o->rg_off[n] = tot; tot += (uint32_t)S.rgn[n]; }
  o->rg_off[nr] = tot;
  o->rg = (int32_t*)malloc(4 * ((size_t)tot + 1));
  for (int n = 0; n < nr; n++) { memcpy(o->rg + o->rg_off[n], S.rg[n], 4 * (size_t)S.rgn[n]); free(S.rg[n]); }
  o->num_fragments = num_fragments; o->frag_len = frag_len; o->color = color;
  second_half(&S, o);
  free(S.rg); free(S.rgn); free(S.rgcap); free(S.J); free(S.ej); free(perm); free(inv); free(grid);
#undef EJ
}
