// groups_device.cuh — SURVEY.md §8 rows a8 + a9 for ONE bundle: the read loop of build_graphs() (per-read junction
// repair, un-pairing, un-stranding, add_read_to_group, fragment counting), the merging of neighbouring groups, the
// colour pass over the three strand lanes and the creation of bundles / bundlenodes.
//
// Reference (CPU, sequential per bundle; short-read de novo path: no --viral, --mix, -L, guides):
//   read loop                                   rlink.cpp:14845-15127
//   add_read_to_group / merge_read_to_group     rlink.cpp:1633-1727, 1281-1631;  merge_fwd_groups :1246-1278
//   re-sort of the junction table               rlink.cpp:15122-15125 (GList::Sort -> gclib/GVec.hh:896-920, not stable)
//   merging of groups closer than bundledist    rlink.cpp:15134-15164
//   colour pass, set_strandcol, neg_prop        rlink.cpp:15196-15299, 1028-1056, 984-1026
//   create_bundle / add_group_to_bundle         rlink.cpp:15340-15427, 1058-1095
//
// The loop is sequential in the reference and its state flows both ways between the three parts (the repair of read n
// looks at what earlier reads did to the junction table, merge_read_to_group un-pairs reads), so this first CUDA
// version keeps ONE walker per bundle and lets the 60 000 bundles of a batch run side by side (group_kernels.cu).
// Pointers of the reference become indices: group = slot in the bundle's group arrays (== CGroup::grid at creation),
// junction = row of the bundle's table, rows >= nj0 being the ones appended here.
// The functions are __host__ __device__ so that tools/groups_hostcheck.cu can step through them on a CPU-only box
// during development; libstgpu only ever calls them from kernels.
#pragma once
#include "stgpu_internal.h"
#include <math.h>

#ifdef __CUDA_ARCH__
#define STG_FSUB(a, b) __fsub_rn((a), (b))
#else
#define STG_FSUB(a, b) ((a) - (b))
#endif
#define STG_HD __host__ __device__ __forceinline__
#define STG_HDN __host__ __device__

#define STG_GERR_GROUPS 0x100u    // capacity of the group / colour / readgroup / appended-junction arrays exceeded
#define STG_LONGINTRONANCHOR 25u  // rlink.h:32

// get_cov(1, a, b, bpcov), rlink.cpp:1830-1838, on lane 1 of the bundle (blocked clamped prefix, BSIZE = 10000)
STG_HD double cov_all(const float* __restrict__ c, uint32_t a, uint32_t b) {
  const int m = (int)((b + 1) / STG_BSIZE);
  const int k = (int)(a / STG_BSIZE);
  double cov = 0;
  for (int i = k + 1; i <= m; i++) cov += (double)c[(size_t)i * STG_BSIZE - 1];
  cov += (double)STG_FSUB(c[b + 1], c[a]);
  if (cov < 0) cov = 0;
  return cov;
}

// good_junc, rlink.cpp:13967-14070, !longreads, on the fields of one row; returns the verdict and the strand the
// reference would leave behind
STG_HD bool good_junc_row(uint32_t jstart, uint32_t jend, int8_t strand, int8_t guide_match, double nreads, double good,
                                  double nm, double ls, double rs, int64_t refstart, const float* __restrict__ c1, int64_t len,
                                  const JuncParams& p, int8_t* strand_out) {
  *strand_out = strand;
  if (p.eonly && !guide_match) { *strand_out = 0; return false; }
  if (guide_match) return true;
  if (good < (double)p.junctionthr) { *strand_out = 0; return false; }
  bool mismatch = false;
  if (nm != 0.0 && round(nm) == round(nreads)) {
    mismatch = true;
    if (jend - jstart > 100000u /*longintron*/ && nreads < 100 * 0.1 /*CHI_WIN*ERROR_PERC*/) { *strand_out = 0; return false; }
  }
  const float mult = (float)(1 / 0.1);
  const int bw = 5;
  int j = (int)((int64_t)jstart - refstart);
  double lleft = 0, lright = 0;
  if (j + bw + 1 < len && j >= bw) {
    lleft = cov_all(c1, (uint32_t)(j - bw + 1), (uint32_t)j);
    lright = cov_all(c1, (uint32_t)(j + 1), (uint32_t)(j + bw));
  }
  if (lleft > 1 / 0.1 && ls * mult < 0.1 * lleft && (mismatch || lright > lleft * (1 - 0.1))) { *strand_out = 0; return false; }
  j = (int)((int64_t)jend - refstart - 1);
  double rleft = 0, rright = 0;
  if (j + bw + 1 < len && j >= bw) {
    rleft = cov_all(c1, (uint32_t)(j - bw + 1), (uint32_t)j);
    rright = cov_all(c1, (uint32_t)(j + 1), (uint32_t)(j + bw));
  }
  if (rright > 1 / 0.1 && rs * mult < rright * 0.1 && (mismatch || rleft > rright * (1 - 0.1))) { *strand_out = 0; return false; }
  if (nreads * 10 < 0.1 * ls || nreads * 10 < 0.1 * rs) { *strand_out = 0; return false; }
  return true;
}

// three values selected by strand lane: kept as scalars (no indexed local array) so that the walker's state stays in
// registers
struct Tri {
  int32_t a, b, c;
  STG_HD int32_t get(int i) const { return i == 0 ? a : (i == 1 ? b : c); }
  STG_HD void set(int i, int32_t v) { if (i == 0) a = v; else if (i == 1) b = v; else c = v; }
  STG_HD void fill(int32_t v) { a = b = c = v; }
};

// ---- state of one bundle's walk --------------------------------------------------------------------------------
struct Grp {
  // reads: pointers already offset to the first read / segment / pair link / intron of the bundle
  const uint32_t *r_start, *r_end, *r_len, *seg_off, *pair_off;
  const float *r_count, *pair_cnt;
  const int16_t* r_nh;
  const uint8_t* r_flags;
  int8_t* strand; uint32_t *seg_s, *seg_e; int32_t *pair, *rj;   // mutable copies
  uint32_t seg0, pair0;
  int nr;
  // junction table: rows [0, nj0) are the bundle's rows after the junction pass, rows [nj0, nJ) are appended here
  const uint32_t *j_start, *j_end;
  const double *j_nreads, *j_good;
  const int32_t* j_ej;
  const int8_t *jp_good, *jp_good_strand;
  int8_t* jt_strand; double* jt_mm;
  uint32_t *ap_start, *ap_end; int8_t *ap_strand, *ap_mm;
  int nj0, nJ, apcap;
  // groups, colours, readgroup
  uint32_t *g_start, *g_end; int32_t *g_color, *g_next, *g_merge, *g_grid; float *g_cov, *g_multi; int8_t* g_alive;
  int ng, gcap;
  int32_t* eqcol; int ncol, ccap;
  int32_t* rg; uint32_t* rg_cnt; const uint32_t* rcap_off;   // rg is batch-wide; rg_cnt / rcap_off offset to the bundle
  Tri curr, startg;
  uint32_t junctionsupport, bundledist;
  int64_t refstart, len;
  const float* c1;
  JuncParams prm;
  uint32_t err;

  STG_HD int nseg(int n) const { return (int)(seg_off[n + 1] - seg_off[n]); }
  STG_HD uint32_t segi(int n, int i) const { return seg_off[n] - seg0 + (uint32_t)i; }
  STG_HD int njunc(int n) const { return nseg(n) - 1; }
  STG_HD uint32_t rji(int n, int i) const { return seg_off[n] - seg0 - (uint32_t)n + (uint32_t)i; }
  STG_HD int npair(int n) const { return (int)(pair_off[n + 1] - pair_off[n]); }
  STG_HD uint32_t pairi(int n, int p) const { return pair_off[n] - pair0 + (uint32_t)p; }
  STG_HD uint32_t slen(int n, int i) const { const uint32_t k = segi(n, i); return seg_e[k] - seg_s[k] + 1u; }
  STG_HD bool unitig(int n) const { return (r_flags[n] & STG_RF_UNITIG) != 0; }
  STG_HD int8_t jstrand(int k) const { return k < nj0 ? jt_strand[k] : ap_strand[k - nj0]; }
  STG_HD void set_jstrand(int k, int8_t s) { if (k < nj0) jt_strand[k] = s; else ap_strand[k - nj0] = s; }
  STG_HD uint32_t jstart(int k) const { return k < nj0 ? j_start[k] : ap_start[k - nj0]; }
  STG_HD uint32_t jend(int k) const { return k < nj0 ? j_end[k] : ap_end[k - nj0]; }
  STG_HD double jnreads(int k) const { return k < nj0 ? j_nreads[k] : 0.0; }
  STG_HD double jgood(int k) const { return k < nj0 ? j_good[k] : 0.0; }
  STG_HD bool jmm_neg(int k) const { return k < nj0 ? jt_mm[k] < 0 : ap_mm[k - nj0] != 0; }
  STG_HD void set_jmm_neg(int k) { if (k < nj0) jt_mm[k] = -1; else ap_mm[k - nj0] = 1; }
  STG_HD int ej(int k) const { return k < nj0 ? j_ej[k] : k; }
  STG_HD int8_t rjstrand(int n, int i) const { return jstrand(rj[rji(n, i)]); }   // CReadAln::juncs[i]->strand
};

// good_junc, mutating like the reference (rlink.cpp:13967): on failure the junction loses its strand.  For the rows of
// the junction pass the verdict was computed once per row by junc_pass_kernel (it only depends on the row and bpcov,
// and a row is only evaluated while its strand is still the one the pass left); appended rows are evaluated here.
STG_HD bool good_junc_mut(Grp& S, int j) {
  int8_t s;
  bool g;
  if (j < S.nj0) { g = S.jp_good[j] == 1; s = S.jp_good_strand[j]; }
  else g = good_junc_row(S.jstart(j), S.jend(j), S.jstrand(j), 0, 0.0, 0.0, 0.0, 0.0, 0.0, S.refstart, S.c1, S.len, S.prm, &s);
  S.set_jstrand(j, s);
  return g;
}

STG_HD int new_color(Grp& S, int* color) {   // eqcol.Add(color); color++
  if (S.ncol < S.ccap) S.eqcol[S.ncol++] = *color;
  else S.err |= STG_GERR_GROUPS;
  return (*color)++;
}

STG_HD int new_group(Grp& S, uint32_t start, uint32_t end, int color, float cov, float multi) {
  int g = S.ng;
  if (g < S.gcap) S.ng++;
  else { S.err |= STG_GERR_GROUPS; g = S.gcap - 1; }
  S.g_start[g] = start; S.g_end[g] = end; S.g_color[g] = color; S.g_next[g] = -1;
  S.g_cov[g] = cov; S.g_multi[g] = multi; S.g_alive[g] = 1; S.g_merge[g] = g;
  return g;
}

STG_HD void rg_add(Grp& S, int n, int g) {
  const uint32_t base = S.rcap_off[n], cap = S.rcap_off[n + 1] - base, k = S.rg_cnt[n];
  if (k < cap) { S.rg[(size_t)base + k] = g; S.rg_cnt[n] = k + 1; }
  else S.err |= STG_GERR_GROUPS;
}

STG_HD int find_color(const Grp& S, int c) {
  while (S.eqcol[c] != c) c = S.eqcol[c];
  return c;
}

// rlink.cpp:1246-1278
STG_HD void merge_fwd_groups(Grp& S, int g1, int g2) {
  S.g_end[g1] = S.g_end[g2];
  int c1 = find_color(S, S.g_color[g1]), c2 = find_color(S, S.g_color[g2]);
  S.g_color[g2] = c2;
  if (c1 < c2) S.eqcol[c2] = c1;
  else if (c1 > c2) { S.eqcol[c1] = c2; c1 = c2; }
  S.g_color[g1] = c1;
  S.g_cov[g1] += S.g_cov[g2];
  S.g_next[g1] = S.g_next[g2];
  S.g_merge[g2] = g1;
  S.g_multi[g1] += S.g_multi[g2];
  S.g_alive[g2] = 0;
}

STG_HD void unpair(Grp& S, int n, int p, int np) {   // readlist[n]->pair_idx[p] = -1 and the mirror entry of np
  S.pair[S.pairi(n, p)] = -1;
  const int m = S.npair(np);
  for (int j = 0; j < m; j++)
    if (S.pair[S.pairi(np, j)] == n) { S.pair[S.pairi(np, j)] = -1; break; }
}

STG_HD bool keep_exon(const Grp& S, int n, int i) {   // the small-exon rule, rlink.cpp:1327-1342 (short reads)
  if (S.slen(n, i) >= S.junctionsupport) return true;
  const int nj = S.njunc(n);
  if (i < nj) {
    if (!S.rjstrand(n, i) && (!i || !S.rjstrand(n, i - 1))) return false;
  } else if (nj) {
    if (!S.rjstrand(n, i - 1)) return false;
  }
  return true;
}

// the mate np (> n) continues the fragment of read n: rlink.cpp:1561-1567 / 1604-1608
STG_HD bool mate_continues(const Grp& S, int n, int np, uint32_t localdist) {
  return np != -1 && np > n && S.r_end[n] < S.r_start[np] && S.r_end[n] + localdist > S.r_start[np] &&
         (S.slen(np, 0) >= S.junctionsupport || (S.njunc(np) && S.rjstrand(np, 0)));
}

// rlink.cpp:1281-1631
STG_HD int merge_read_to_group(Grp& S, int n, int np, int p, float readcov, int sno, int readcol, int color, Tri& usedcol) {
  const uint32_t localdist = S.bundledist + STG_LONGINTRONANCHOR;   // !longreads && !mixedMode
  if (np > -1 && np < n && S.r_end[np] < S.r_start[n] && S.r_end[np] + localdist > S.r_start[n] &&
      (S.slen(np, S.nseg(np) - 1) >= S.junctionsupport || (S.njunc(np) && S.rjstrand(np, S.njunc(np) - 1))) &&
      (S.slen(n, 0) >= S.junctionsupport || (S.njunc(n) && S.rjstrand(n, 0))))
    return color;   // the pair was placed as one fragment when np was processed
  bool first = true;
  int currgroup = S.curr.get(sno);
  if (currgroup != -1) {
    int lastgroup = -1;
    const uint32_t rstart = S.r_start[n];
    while (currgroup != -1 && rstart > S.g_end[currgroup]) { lastgroup = currgroup; currgroup = S.g_next[currgroup]; }
    if (currgroup == -1 || S.seg_e[S.segi(n, 0)] < S.g_start[currgroup]) currgroup = lastgroup;
    int thisgroup = currgroup;
    int ncoord = S.nseg(n);
    int lastpushedgroup = -1;
    int i = 0;
    while (i < ncoord) {
      bool keep = true;
      if (S.slen(n, i) < S.junctionsupport) {
        const int nj = S.njunc(n);
        if (i < nj) {
          if (!S.rjstrand(n, i)) {
            if (!i || !S.rjstrand(n, i - 1)) {
              keep = false;
              if (!i && lastgroup != -1) currgroup = lastgroup;
            }
          }
        } else if (nj) {
          if (!S.rjstrand(n, i - 1)) keep = false;
        }
      }
      if (keep) {
        const uint32_t ss = S.seg_s[S.segi(n, i)], se = S.seg_e[S.segi(n, i)];
        while (thisgroup != -1 && ss > S.g_end[thisgroup]) { lastgroup = thisgroup; thisgroup = S.g_next[thisgroup]; }
        if (thisgroup != -1 && se + localdist >= S.g_start[thisgroup]) {
          if (!i) {
            int gr = S.g_grid[thisgroup];
            while (S.g_merge[gr] != gr) gr = S.g_merge[gr];
            S.g_grid[thisgroup] = gr;
            const uint32_t base = S.rcap_off[n], cnt = S.rg_cnt[n];
            for (uint32_t g = 0; g < cnt; g++) {
              int x = S.rg[(size_t)base + g];
              while (S.g_merge[x] != x) x = S.g_merge[x];
              if (x == gr) { first = false; break; }
            }
            if (np > -1 && S.r_nh[np] && np < n) {
              const int thiscol = find_color(S, S.g_color[thisgroup]);
              S.g_color[thisgroup] = thiscol;
              if (thiscol != readcol) {
                unpair(S, n, p, np);
                readcol = usedcol.get(sno);
                if (readcol < 0) { usedcol.set(sno, color); readcol = new_color(S, &color); }
              }
            }
          }
          if (S.r_nh[n] > 1) S.g_multi[thisgroup] += readcov * (float)(se - ss + 1u);
          if (ss < S.g_start[thisgroup]) S.g_start[thisgroup] = ss;
          int nextgroup = S.g_next[thisgroup];
          while (nextgroup != -1 && se >= S.g_start[nextgroup]) {
            merge_fwd_groups(S, thisgroup, nextgroup);
            nextgroup = S.g_next[thisgroup];
          }
          if (se > S.g_end[thisgroup]) S.g_end[thisgroup] = se;
          int gc = find_color(S, S.g_color[thisgroup]);
          S.g_color[thisgroup] = gc;
          if (readcol != gc) {
            if (i && !S.rjstrand(n, i - 1)) readcol = gc;
            else if (readcol < gc) { S.eqcol[gc] = readcol; S.g_color[thisgroup] = readcol; }
            else { S.eqcol[readcol] = gc; readcol = gc; }
          }
          if (S.g_grid[thisgroup] != lastpushedgroup) {
            if (first) rg_add(S, n, S.g_grid[thisgroup]);
            lastpushedgroup = S.g_grid[thisgroup];
          }
          S.g_cov[thisgroup] += (float)(se - ss + 1u) * readcov;
        } else {
          if (!i && np > -1 && S.r_nh[np] && np < n) {
            unpair(S, n, p, np);
            readcol = usedcol.get(sno);
            if (readcol < 0) { usedcol.set(sno, color); readcol = new_color(S, &color); }
          } else if (i && !S.rjstrand(n, i - 1)) {
            usedcol.set(sno, color); readcol = new_color(S, &color);
          }
          float multi = 0;
          if (S.r_nh[n] > 1) multi = readcov * (float)(se - ss + 1u);
          const int ngroup = new_group(S, ss, se, readcol, (float)(se - ss + 1u) * readcov, multi);
          S.g_grid[ngroup] = ngroup;
          if (lastgroup != -1) S.g_next[lastgroup] = ngroup;
          S.g_next[ngroup] = thisgroup;
          rg_add(S, n, ngroup);
          lastpushedgroup = ngroup;
          thisgroup = ngroup;
        }
        if (i == ncoord - 1 && mate_continues(S, n, np, localdist)) {
          n = np; np = -1;
          ncoord = S.nseg(n);
          first = true;
          if (S.r_start[n] > S.g_end[thisgroup]) {
            int nextgroup = S.g_next[thisgroup];
            while (nextgroup != -1 && S.r_start[n] >= S.g_start[nextgroup]) {
              merge_fwd_groups(S, thisgroup, nextgroup);
              nextgroup = S.g_next[thisgroup];
            }
            if (S.r_start[n] > S.g_end[thisgroup]) S.g_end[thisgroup] = S.r_start[n];
          }
          lastpushedgroup = -1;
          i = -1;
        }
      }
      i++;
    }
  } else {   // first group of this strand lane
    if (np > -1 && S.r_nh[np] && np < n) {
      unpair(S, n, p, np);
      readcol = usedcol.get(sno);
      if (readcol < 0) { usedcol.set(sno, color); readcol = new_color(S, &color); }
    }
    int ncoord = S.nseg(n);
    int lastgroup = -1;
    float multi = 0;
    if (S.r_nh[n] > 1) multi = readcov;
    int i = 0;
    while (i < ncoord) {
      const bool keep = keep_exon(S, n, i);
      if (!i || keep) {
        if (i && !S.rjstrand(n, i - 1)) { usedcol.set(sno, color); readcol = new_color(S, &color); }
        const uint32_t ss = S.seg_s[S.segi(n, i)], se = S.seg_e[S.segi(n, i)];
        const float flen = (float)(se - ss + 1u);
        const int ngroup = new_group(S, ss, se, readcol, flen * readcov, flen * multi);
        S.g_grid[ngroup] = ngroup;
        if (lastgroup != -1) S.g_next[lastgroup] = ngroup;
        else currgroup = ngroup;
        lastgroup = ngroup;
        if (keep) {
          rg_add(S, n, ngroup);
          if (i == ncoord - 1 && mate_continues(S, n, np, localdist)) {
            n = np; np = -1;
            ncoord = S.nseg(n);
            if (S.seg_e[S.segi(n, 0)] > S.g_end[lastgroup]) {
              S.g_end[lastgroup] = S.seg_e[S.segi(n, 0)];
              if (first) rg_add(S, n, ngroup);
              S.g_cov[lastgroup] += (float)S.slen(n, 0) * readcov;
            }
            i = 0;
          }
        }
      }
      i++;
    }
  }
  S.curr.set(sno, currgroup);
  if (S.startg.get(sno) == -1) S.startg.set(sno, currgroup);
  return color;
}

// rlink.cpp:1633-1727: one pass per pair link, then one for the unpaired share of the read (p == npair)
STG_HD int add_read_to_group(Grp& S, int n, int color) {
  Tri usedcol;
  usedcol.fill(-1);
  float single_count = S.r_count[n];
  const int npr = S.npair(n);
  for (int p = 0; p <= npr; p++) {
    int sno = (int)S.strand[n] + 1, np = -1, readcol = -1;
    double readcov = 0;
    bool go = false;
    if (p < npr) {
      np = S.pair[S.pairi(n, p)];
      if (np > -1 && S.r_nh[np]) {
        const int snop = (int)S.strand[np] + 1;
        if (sno != snop) {
          if (sno == 1) sno = snop;
          else if (snop != 1) { unpair(S, n, p, np); np = -1; }
        }
        if (np > -1) {
          const float pc = S.pair_cnt[S.pairi(n, p)];
          single_count -= pc;
          readcol = usedcol.get(sno);
          if (np < n) {
            int grouppair = S.rg[(size_t)S.rcap_off[np]];
            while (S.g_merge[grouppair] != grouppair) grouppair = S.g_merge[grouppair];
            S.rg[(size_t)S.rcap_off[np]] = grouppair;
            readcol = find_color(S, S.g_color[grouppair]);
            S.g_color[grouppair] = readcol;
          } else if (usedcol.get(sno) < 0) { usedcol.set(sno, color); readcol = new_color(S, &color); }
          if (!S.unitig(n)) readcov = pc;
          go = true;
        }
      }
    } else if (single_count > 0.000001 /*epsilon, rlink.h:25*/) {
      readcol = usedcol.get(sno);
      if (readcol < 0) { usedcol.set(sno, color); readcol = new_color(S, &color); }
      if (!S.unitig(n)) readcov = single_count;
      go = true;
    }
    if (go) color = merge_read_to_group(S, n, np, p < npr ? p : -1, (float)readcov, sno, readcol, color, usedcol);
  }
  return color;
}

// CJunction::operator< (bundle.h:155-162) through GList::DefaultCompareProc (gclib/GList.hh:85-90)
STG_HD int junc_default_cmp(const Grp& S, int a, int b) {
  const uint32_t sa = S.jstart(a), sb = S.jstart(b);
  if (sa != sb) return sa < sb ? -1 : 1;
  const uint32_t ea = S.jend(a), eb = S.jend(b);
  if (ea != eb) return ea < eb ? -1 : 1;
  const int8_t ta = S.jstrand(a), tb = S.jstrand(b);
  if (ta != tb) return ta > tb ? -1 : 1;
  return 0;
}

// the reference's quicksort (gclib/GVec.hh:896-920: Hoare partition, middle pivot, not stable) on a row list; the
// sub-ranges of a partition are disjoint, so the order in which they are finished does not change the permutation
STG_HD void sort_rows_ref(const Grp& S, int32_t* list, int n) {
  int stack[2 * 48];
  int sp = 0;
  if (n < 2) return;
  stack[sp++] = 0; stack[sp++] = n - 1;
  while (sp) {
    const int R = stack[--sp], L = stack[--sp];
    int I = L, Jx = R;
    const int P = list[L + ((R - L) >> 1)];
    do {
      while (junc_default_cmp(S, list[I], P) < 0) I++;
      while (junc_default_cmp(S, list[Jx], P) > 0) Jx--;
      if (I <= Jx) {
        const int t = list[I]; list[I] = list[Jx]; list[Jx] = t;
        I++; Jx--;
      }
    } while (I <= Jx);
    const bool left_todo = L < Jx, right_todo = I < R;
    if (left_todo && right_todo) {
      if (Jx - L > R - I) { stack[sp++] = L; stack[sp++] = Jx; stack[sp++] = I; stack[sp++] = R; }
      else { stack[sp++] = I; stack[sp++] = R; stack[sp++] = L; stack[sp++] = Jx; }
    } else if (left_todo) { stack[sp++] = L; stack[sp++] = Jx; }
    else if (right_todo) { stack[sp++] = I; stack[sp++] = R; }
  }
}

// the repair of the junctions of read n, its un-pairing and un-stranding: rlink.cpp:14845-15092.  Returns `keep`.
STG_HD bool repair_read(Grp& S, int n) {
  bool keep = true;
  const int nj = S.njunc(n);
  for (int i = 0; i < nj; i++) {
    const int jd = S.rj[S.rji(n, i)];
    const bool changeright = S.jgood(jd) < 0;
    const bool changeleft = S.jnreads(jd) < 0;
    if (S.jstrand(jd)) {
      if (!good_junc_mut(S, jd)) S.set_jmm_neg(jd);
      if (changeleft || changeright) S.set_jstrand(jd, 0);
      if (S.jmm_neg(jd)) S.set_jstrand(jd, 0);
    }
    if (!S.jstrand(jd)) {
      if (changeleft || changeright) {
        const uint32_t k0 = S.segi(n, i), k1 = S.segi(n, i + 1);
        uint32_t newstart = S.seg_e[k0], newend = S.seg_s[k1];
        int jk = -1, ek = -1;
        if (changeleft) {
          jk = abs((int)S.jnreads(jd));
          if (S.jnreads(jk) < 0) newstart = S.seg_s[k0] - 1;
          else { newstart = S.jstart(jk); if (S.jstrand(jk)) jk = -1; }
        }
        if (changeright) {
          ek = abs((int)S.jgood(jd));
          const int er = S.ej(ek);
          if (S.jgood(er) < 0) newend = S.seg_e[k1] + 1;
          else { newend = S.jend(er); if (S.jstrand(er)) ek = -1; }
        }
        if (newstart >= S.seg_s[k0] && newend <= S.seg_e[k1] && newstart <= newend) {
          S.seg_e[k0] = newstart;
          S.seg_s[k1] = newend;
          if (!S.jmm_neg(jd)) {
            const int8_t rs = S.strand[n];
            int found = -1;
            bool search = false;
            if (jk > 0) {   // neighbours of jk in start order, then the appended rows
              search = true;
              for (int k = jk; found < 0 && k > 0 && S.jstart(k) == newstart; k--)
                if (rs == S.jstrand(k) && S.jend(k) == newend) found = k;
              for (int k = jk + 1; found < 0 && k < S.nJ && S.jstart(k) == newstart; k++)
                if (rs == S.jstrand(k) && S.jend(k) == newend) found = k;
            } else if (ek > 0) {   // neighbours of ek in end order
              search = true;
              for (int k = ek; found < 0 && k > 0 && S.jend(S.ej(k)) == newend; k--) {
                const int r = S.ej(k);
                if (rs == S.jstrand(r) && S.jstart(r) == newstart) found = r;
              }
              for (int k = ek + 1; found < 0 && k < S.nJ && S.jend(S.ej(k)) == newend; k++) {
                const int r = S.ej(k);
                if (rs == S.jstrand(r) && S.jstart(r) == newstart) found = r;
              }
            }
            if (search) {
              for (int k = S.nj0; found < 0 && k < S.nJ; k++)
                if (rs == S.jstrand(k) && S.jstart(k) == newstart && S.jend(k) == newend) found = k;
              if (found < 0) {   // new CJunction(newstart, newend, strand): rlink.cpp:14958-14969, 15000-15011
                const int a = S.nJ - S.nj0;
                if (a < S.apcap) {
                  S.ap_start[a] = newstart; S.ap_end[a] = newend; S.ap_strand[a] = rs; S.ap_mm[a] = 0;
                  found = S.nJ++;
                } else S.err |= STG_GERR_GROUPS;
              }
              if (found >= 0) S.rj[S.rji(n, i)] = found;
            }
          }
        }
      }
      if (keep) {   // a bad junction: the read loses its mates (rlink.cpp:15027-15043)
        const int npr = S.npair(n);
        for (int p = 0; p < npr; p++) {
          const int np = S.pair[S.pairi(n, p)];
          if (np > -1) {
            unpair(S, n, p, np);
            if (!S.njunc(np)) S.strand[np] = 0;
          }
        }
        keep = false;
      }
    }
  }
  if (!keep) {
    if (S.strand[n]) {
      bool keepstrand = false;
      for (int j = 0; j < nj; j++) if (S.rjstrand(n, j)) { keepstrand = true; break; }
      if (!keepstrand) S.strand[n] = 0;
    }
  } else if (!nj && S.strand[n]) {   // unspliced stranded read: it keeps its strand only if a later mate supports it
    const int npr = S.npair(n);
    if (npr) {
      bool keepstrand = false;
      for (int p = 0; p < npr && !keepstrand; p++) {
        const int np = S.pair[S.pairi(n, p)];
        if (np > -1 && n < np) {
          const int njp = S.njunc(np);
          if (!njp) {
            if (S.strand[np] == S.strand[n]) keepstrand = true;
          } else {
            for (int j = 0; j < njp; j++) {
              const int jx = S.rj[S.rji(np, j)];
              if (S.jstrand(jx) && good_junc_mut(S, jx)) { keepstrand = true; break; }
            }
            if (!keepstrand) S.strand[np] = 0;
          }
        }
      }
      if (!keepstrand) S.strand[n] = 0;
    }
  }
  return keep;
}

// Walks the reads of the bundle in order.  jt_perm / jt_inv: [nj0 + apcap] scratch for the re-sort of the junction
// table; on return jt_perm[k] = working row of the k-th row of the re-sorted table and rj[] holds re-sorted rows.
// one read of the loop, the reference's way: repair -> add_read_to_group -> fragment counting (rlink.cpp:15105-15116)
STG_HD int walk_read(Grp& S, int n, int color, double& num_fragments, double& frag_len) {
  const bool keep = repair_read(S, n);
  color = add_read_to_group(S, n, color);
  if (!S.unitig(n)) frag_len += (double)((float)S.r_len[n] * S.r_count[n]);
  double single_count = S.r_count[n];
  if (keep) {
    const int npr = S.npair(n);
    for (int p = 0; p < npr; p++) {
      const int pi = S.pair[S.pairi(n, p)];
      if (pi != -1 && n > pi && S.r_nh[pi]) single_count -= S.pair_cnt[S.pairi(n, p)];
    }
  }
  if (!S.unitig(n) && single_count > 0.000001) num_fragments += single_count;
  return color;
}

// after the last read: jt_perm / jt_inv: [nj0 + apcap] scratch for the re-sort of the junction table; on return
// jt_perm[k] = working row of the k-th row of the re-sorted table and rj[] holds re-sorted rows
STG_HD void walk_finish(Grp& S, int32_t* jt_perm, int32_t* jt_inv) {
  // merge groups that are close together: rlink.cpp:15134-15164 (no guides: no boundaries)
  if (S.bundledist) {
    for (int sno = 0; sno < 3; sno++) {
      int lastgroup = -1, procgroup = S.startg.get(sno);
      while (procgroup != -1) {
        if (lastgroup != -1 && S.g_start[procgroup] - S.g_end[lastgroup] <= S.bundledist) {
          merge_fwd_groups(S, lastgroup, procgroup);
          procgroup = S.g_next[lastgroup];
          continue;
        }
        lastgroup = procgroup;
        procgroup = S.g_next[procgroup];
      }
    }
  }
  // appended rows force a re-sort of the table (rlink.cpp:15122-15125); rows whose strand was zeroed meanwhile have
  // CHANGED keys and may tie with other rows, so the reference's own quicksort decides the order
  for (int j = 0; j < S.nJ; j++) jt_perm[j] = j;
  if (S.nJ > S.nj0) {
    sort_rows_ref(S, jt_perm, S.nJ);
    for (int j = 0; j < S.nJ; j++) jt_inv[jt_perm[j]] = j;
    const uint32_t nintr = (S.seg_off[S.nr] - S.seg0) - (uint32_t)S.nr;
    for (uint32_t k = 0; k < nintr; k++) S.rj[k] = jt_inv[S.rj[k]];
  }
}

struct NoHook { STG_HD void operator()(Grp&, int) const {} };

// The whole loop, one read after the other.  `before_read(S, n)` runs ahead of every read (the kernel uses it to move
// the group arrays out of shared memory before they can overflow).
template <class Hook = NoHook>
STG_HD void groups_first_half(Grp& S, int32_t* jt_perm, int32_t* jt_inv, double* num_fragments_out, double* frag_len_out,
                              Hook before_read = Hook()) {
  int color = 0;
  double num_fragments = 0, frag_len = 0;
  for (int n = 0; n < S.nr && !S.err; n++) {
    before_read(S, n);
    color = walk_read(S, n, color, num_fragments, frag_len);
  }
  *num_fragments_out = num_fragments; *frag_len_out = frag_len;
  if (S.err) return;
  walk_finish(S, jt_perm, jt_inv);
}

// ---- the parallel form of the loop ------------------------------------------------------------------------------
// In a deep bundle almost every read is "simple": its junctions are all good, every exon it brings lies INSIDE a group
// that already exists, and the colours it meets are already united.  Such a read changes no structure — it adds
// coverage to groups, records its readgroup entries, may mint one fresh colour that is immediately attached to the
// group's colour, and moves the lane cursor — so what it does can be worked out from a snapshot of the state without
// knowing what the simple reads just before it did.  classify_read() does that, read-only, for one read; a warp
// classifies 32 consecutive reads at once, commits the leading run of simple ones (order only matters for the float
// sums, which one lane adds in read order), hands the first non-simple read to walk_read(), and starts over behind it.
// Anything classify_read() is not sure about is "not simple": the sequential code is the definition of the result.
#define STG_FP_MAXV 6   // group visits of one read (its exons + the exons of a mate that continues the fragment)
struct FastRec {
  int32_t grp[STG_FP_MAXV];     // visited groups, in order
  float cov[STG_FP_MAXV];       // coverage added to cov_sum
  float multi[STG_FP_MAXV];     // ... to multi (valid where has_multi has the bit)
  int32_t gcol[STG_FP_MAXV];    // root colour of the visited group (CGroup::color is left pointing at it)
  int32_t rgv[STG_FP_MAXV];     // readgroup pushes: the first nrg0 go to the read, the next nrg1 to the mate npc
  int32_t nv, nrg0, nrg1, has_multi;
  int32_t sno, curr;            // currgroup[sno] = curr afterwards (sno = -1: untouched)
  int32_t newcol, fresh_gc;     // mints a colour, attached to fresh_gc
  int32_t gp_np, gp_root, gp_col;   // earlier mate: readgroup[np][0] = gp_root, color of gp_root = gp_col (gp_np = -1: none)
  int32_t npc;                  // mate that continued the fragment (-1: none)
  int32_t unpair_np;            // earlier mate the read is un-paired from (-1: none)
  int32_t xg; uint32_t xs, xe;  // group the read grows and its new bounds (-1: none); such a read ends the run
  double frag_len, numfrag;     // what the read adds to the bundle's sums
};

// development aid: tools/groups_hostcheck.cu counts why reads were not simple (reason = position of the test below)
#if defined(STG_FP_COUNT_BAILS)
extern long long stg_fp_bails[64];   // [why]: steps whose run of simple reads was cut by test `why`; [0]: steps; [63]: full steps
extern int stg_fp_last;
#endif
#if !defined(__CUDA_ARCH__) && defined(STG_FP_COUNT_BAILS)
inline int fp_bail(int why) { stg_fp_last = why; return 0; }
#else
STG_HD int fp_bail(int) { return 0; }
#endif

// returns 1: simple (R holds what the read does); 3: simple but for growing ONE group (R.xg, no merge): it may be
// committed as the last read of a run, the reads behind it have to be classified against the grown group; 0: not
// simple; 2: simple or not, it has to wait for an earlier read of the same step (its mate), i.e. the step ends in front
// of it and the next one starts with it
// What a read took for granted about the bounds of groups it walked past or stopped in front of: group g[i] must keep
// its END below x[i] (bit i of kind clear) or its START above x[i] (bit set).  With these recorded, reads that grow a
// group need not end a run: a later read of the step waits only if an earlier one moves a bound it depends on.
#define STG_FP_MAXC 6
struct FastDeps { int32_t n; uint32_t kind; int32_t g[STG_FP_MAXC]; uint32_t x[STG_FP_MAXC]; };

// hop_budget: a read that would have to follow the chain of groups for more than this many links (it lies far behind the
// cursor: every link is a dependent load, and the step waits for its slowest thread) gives up and waits for the next
// step, whose cursor is closer — unless it is the first read of the window (progress).
template <bool DEPS>
STG_HD int classify_read(const Grp& S, int n, int lo, int hi, FastRec& __restrict__ R, FastDeps* __restrict__ D = nullptr,
                         int hop_budget = 0x7fffffff) {
  if (n == lo) hop_budget = 0x7fffffff;
  R.nv = 0; R.nrg0 = 0; R.nrg1 = 0; R.has_multi = 0; R.sno = -1; R.newcol = 0; R.fresh_gc = -1; R.gp_np = -1; R.npc = -1; R.unpair_np = -1; R.xg = -1;
  if (DEPS) { D->n = 0; D->kind = 0; }
#ifdef __CUDA_ARCH__
  // the lines this read will need one after the other (its exons and junction rows, its mate's fields) are asked for
  // at once: the classification is a chain of dependent loads, this shortens it
  {
    const uint32_t so = S.seg_off[n] - S.seg0, po = S.pair_off[n];
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.seg_s + so));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.seg_e + so));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.rj + (so - (uint32_t)n)));
    if (S.pair_off[n + 1] != po) {
      const int m = S.pair[po - S.pair0];
      if (m >= 0 && m < S.nr) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.strand + m));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.r_nh + m));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.r_start + m));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.r_end + m));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.seg_off + m));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.rg_cnt + m));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(S.rcap_off + m));
      }
    }
  }
#endif
  const int nj = S.njunc(n);
  // the repair must leave the read alone: every junction stranded, good, not demoted, not flagged
  for (int i = 0; i < nj; i++) {
    const int jd = S.rj[S.rji(n, i)];
    if (jd >= S.nj0) return fp_bail(1);
    const int8_t st = S.jt_strand[jd];
    if (!st || S.jp_good[jd] != 1 || S.jp_good_strand[jd] != st || S.j_nreads[jd] < 0 || S.j_good[jd] < 0 || S.jt_mm[jd] < 0)
      return fp_bail(2);
  }
  const int npr = S.npair(n);
  if (npr > 1) return fp_bail(3);
  const int8_t strand_n = S.strand[n];
  if (!nj && strand_n && npr) return fp_bail(4);   // may lose its strand (rlink.cpp:15056-15092)
  int sno = (int)strand_n + 1;
  float single_count = S.r_count[n];
  int np = -1, readcol = -1;
  bool pass = false, fresh = false;
  double readcov = 0;
  if (npr == 1) {
    const int m = S.pair[S.pairi(n, 0)];
    if (m > -1 && S.r_nh[m]) {
      if (m >= lo && m < n) { fp_bail(5); return 2; }   // the earlier mate is committed in this very step
      const int snop = (int)S.strand[m] + 1;
      if (sno != snop) {
        if (sno == 1) sno = snop;
        else if (snop != 1) return fp_bail(6);    // un-pairing
      }
      const float pc = S.pair_cnt[S.pairi(n, 0)];
      single_count -= pc;
      if (m < n) {
        if (!S.rg_cnt[m]) return fp_bail(7);
        int grouppair = S.rg[(size_t)S.rcap_off[m]];
        while (S.g_merge[grouppair] != grouppair) grouppair = S.g_merge[grouppair];
        readcol = find_color(S, S.g_color[grouppair]);
        R.gp_np = m; R.gp_root = grouppair; R.gp_col = readcol;
      } else fresh = true;
      if (!S.unitig(n)) readcov = pc;
      np = m;
      pass = true;
    }
  }
  if (single_count > 0.000001) {
    if (pass) return fp_bail(8);                  // two passes of the same read see each other's effects
    sno = (int)strand_n + 1;
    fresh = true;
    if (!S.unitig(n)) readcov = single_count;
    np = -1;
    pass = true;
  }
  if (pass) {
    const float rc = (float)readcov;
    const uint32_t localdist = S.bundledist + STG_LONGINTRONANCHOR;
    int cn = n, cnp = np;
    const bool placed = cnp > -1 && cnp < cn && S.r_end[cnp] < S.r_start[cn] && S.r_end[cnp] + localdist > S.r_start[cn] &&
                        (S.slen(cnp, S.nseg(cnp) - 1) >= S.junctionsupport || (S.njunc(cnp) && S.rjstrand(cnp, S.njunc(cnp) - 1))) &&
                        (S.slen(cn, 0) >= S.junctionsupport || (S.njunc(cn) && S.rjstrand(cn, 0)));
    if (!placed) {
      // a read may GROW one group (no merge, no new group): its later tests see the grown bounds
      int xg = -1;
      uint32_t xs = 0, xe = 0;
      auto GS = [&](int gq) -> uint32_t { return gq == xg ? xs : S.g_start[gq]; };
      auto GE = [&](int gq) -> uint32_t { return gq == xg ? xe : S.g_end[gq]; };
      bool full = false;   // more bounds than FastDeps holds
      auto depends = [&](int gq, bool start_above, uint32_t x) {
        if (!DEPS) return;
        if (D->n == STG_FP_MAXC) { full = true; return; }
        D->g[D->n] = gq; D->x[D->n] = x;
        if (start_above) D->kind |= 1u << D->n;
        D->n++;
      };
      int currgroup = S.curr.get(sno);
      if (currgroup == -1) return fp_bail(9);
      int lastgroup = -1;
      const uint32_t rstart = S.r_start[cn];
      while (currgroup != -1 && rstart > GE(currgroup)) {
        if (--hop_budget < 0) { fp_bail(20); return 2; }
        lastgroup = currgroup; currgroup = S.g_next[currgroup];
      }
      if (lastgroup != -1) depends(lastgroup, false, rstart);           // the last group walked past ends below the read
      if (currgroup != -1 && S.seg_e[S.segi(cn, 0)] < GS(currgroup)) depends(currgroup, true, S.seg_e[S.segi(cn, 0)]);
      if (currgroup == -1 || S.seg_e[S.segi(cn, 0)] < GS(currgroup)) currgroup = lastgroup;
      if (currgroup == -1) return fp_bail(10);
      int thisgroup = currgroup, ncoord = S.nseg(cn), lastpushed = -1, i = 0;
      bool first = true;
      while (i < ncoord) {
        if (!keep_exon(S, cn, i)) return fp_bail(11);
        const uint32_t ss = S.seg_s[S.segi(cn, i)], se = S.seg_e[S.segi(cn, i)];
        {
          int skipped = -1;
          while (thisgroup != -1 && ss > GE(thisgroup)) {
            if (--hop_budget < 0) { fp_bail(20); return 2; }
            skipped = thisgroup; thisgroup = S.g_next[thisgroup];
          }
          if (skipped != -1) depends(skipped, false, ss);
        }
        if (thisgroup == -1 || se + localdist < GS(thisgroup)) return fp_bail(12);             // a new group
        if (ss < GS(thisgroup) || se > GE(thisgroup)) {                                          // the group grows
          if (xg != -1 && xg != thisgroup) return fp_bail(13);                                   // ... a second one
          const int nx = S.g_next[thisgroup];
          const uint32_t ns2 = ss < GS(thisgroup) ? ss : GS(thisgroup), ne2 = se > GE(thisgroup) ? se : GE(thisgroup);
          if (nx != -1 && se >= S.g_start[nx]) return fp_bail(13);                               // ... into the next one: a merge
          if (nx != -1 && se > GE(thisgroup)) depends(nx, true, se);                              // the next group starts above it
          xg = thisgroup; xs = ns2; xe = ne2;
        }
        if (S.g_grid[thisgroup] != thisgroup || S.g_merge[thisgroup] != thisgroup) return fp_bail(14);
        if (!i) {
          const uint32_t base = S.rcap_off[cn], cnt = S.rg_cnt[cn];
          for (uint32_t g = 0; g < cnt; g++) {
            int x = S.rg[(size_t)base + g];
            while (S.g_merge[x] != x) x = S.g_merge[x];
            if (x == thisgroup) { first = false; break; }
          }
          if (cnp > -1 && S.r_nh[cnp] && cnp < cn && find_color(S, S.g_color[thisgroup]) != readcol) {
            // the mates lie in groups of different colours: the link is cut and the read carries on with a colour of
            // its own (rlink.cpp:1383-1389).  Nothing else looks at the two link entries during this step.
            R.unpair_np = cnp;
            fresh = true;
          }
        }
        if (R.nv == STG_FP_MAXV) return fp_bail(16);
        const float flen = (float)(se - ss + 1u);
        const int v = R.nv++;
        R.grp[v] = thisgroup;
        R.multi[v] = 0.f;
        if (S.r_nh[cn] > 1) { R.multi[v] = rc * flen; R.has_multi |= 1 << v; }
        const int gc = find_color(S, S.g_color[thisgroup]);
        R.gcol[v] = gc;
        if (fresh) { R.newcol = 1; R.fresh_gc = gc; readcol = gc; fresh = false; }
        else if (readcol != gc) {
          if (i && !S.rjstrand(cn, i - 1)) readcol = gc;
          else return fp_bail(17);                                                              // two colours are united
        }
        if (thisgroup != lastpushed) {
          if (first) { R.rgv[R.nrg0 + R.nrg1] = thisgroup; if (cn == n) R.nrg0++; else R.nrg1++; }
          lastpushed = thisgroup;
        }
        R.cov[v] = flen * rc;
        if (i == ncoord - 1 && mate_continues(S, cn, cnp, localdist)) {
          cn = cnp; cnp = -1;
          ncoord = S.nseg(cn);
          first = true;
          if (S.r_start[cn] > GE(thisgroup)) {                                                   // the group grows up to the mate
            if (xg != -1 && xg != thisgroup) return fp_bail(18);
            const int nx = S.g_next[thisgroup];
            if (nx != -1 && S.r_start[cn] >= S.g_start[nx]) return fp_bail(18);                  // a merge
            if (nx != -1) depends(nx, true, S.r_start[cn]);
            xs = GS(thisgroup); xe = S.r_start[cn]; xg = thisgroup;
          }
          lastpushed = -1;
          i = -1;
          R.npc = cn;
        }
        i++;
      }
      if (full) return fp_bail(19);
      R.sno = sno; R.curr = currgroup;
      R.xg = xg; R.xs = xs; R.xe = xe;
    }
  }
  // fragment counting
  R.frag_len = 0; R.numfrag = 0;
  if (!S.unitig(n)) {
    R.frag_len = (double)((float)S.r_len[n] * S.r_count[n]);
    double sc = S.r_count[n];
    if (npr == 1) {
      const int pi = R.unpair_np >= 0 ? -1 : S.pair[S.pairi(n, 0)];
      if (pi != -1 && n > pi && S.r_nh[pi]) sc -= S.pair_cnt[S.pairi(n, 0)];
    }
    if (sc > 0.000001) R.numfrag = sc;
  }
  return R.xg >= 0 ? 3 : 1;
}

// does the growth of `earlier` (a read of the same step, outcome 3) move a bound that the later read depends on?
STG_HD bool grows_past(const FastDeps& later, const FastRec& earlier) {
  for (int i = 0; i < later.n; i++)
    if (later.g[i] == earlier.xg) {
      if (later.kind >> i & 1u) { if (earlier.xs <= later.x[i]) return true; }
      else if (earlier.xe >= later.x[i]) return true;
    }
  return false;
}

#ifdef __CUDA_ARCH__
#define STG_MIN_INTO(p, v) atomicMin((p), (v))
#define STG_MAX_INTO(p, v) atomicMax((p), (v))
#else
#define STG_MIN_INTO(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define STG_MAX_INTO(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#endif

// the part of a simple read's effect that does not depend on the other reads of the step; fresh_id = the colour it
// mints.  The readgroup pushes belong here: read n's list is only touched by n itself, and a mate's list by the one
// read of the step that continues into it (the caller makes sure of that).
STG_HD void commit_free(Grp& S, const FastRec& R, int n, int fresh_id) {
  if (R.gp_np >= 0) { S.rg[(size_t)S.rcap_off[R.gp_np]] = R.gp_root; S.g_color[R.gp_root] = R.gp_col; }
  for (int v = 0; v < R.nv; v++) S.g_color[R.grp[v]] = R.gcol[v];
  if (R.newcol) {
    if (fresh_id < S.ccap) S.eqcol[fresh_id] = R.fresh_gc;
    else S.err |= STG_GERR_GROUPS;
  }
  if (R.unpair_np >= 0) unpair(S, n, 0, R.unpair_np);
  if (R.xg >= 0) { STG_MIN_INTO(&S.g_start[R.xg], R.xs); STG_MAX_INTO(&S.g_end[R.xg], R.xe); }   // several reads of a run may grow a group
  for (int k = 0; k < R.nrg0; k++) rg_add(S, n, R.rgv[k]);
  for (int k = 0; k < R.nrg1; k++) rg_add(S, R.npc, R.rgv[R.nrg0 + k]);
}

// the ordered part: float sums in read order, the lane cursor, the bundle's fragment sums
STG_HD void commit_ordered(Grp& S, const FastRec& R, double& num_fragments, double& frag_len) {
  for (int v = 0; v < R.nv; v++) {
    if (R.has_multi >> v & 1) S.g_multi[R.grp[v]] += R.multi[v];
    S.g_cov[R.grp[v]] += R.cov[v];
  }
  if (R.sno >= 0) S.curr.set(R.sno, R.curr);
  if (R.frag_len != 0) frag_len += R.frag_len;
  if (R.numfrag != 0) num_fragments += R.numfrag;
}

// Host-side emulation of the warp's schedule (tools/groups_hostcheck.cu): the lanes of a step are run one after the
// other against the same state, which is what the warp does, since classify_read() only reads.
#define STG_STEP_MAX 256
inline void groups_first_half_steps(Grp& S, int32_t* jt_perm, int32_t* jt_inv, double* num_fragments_out, double* frag_len_out,
                                    int64_t* n_simple = nullptr, int W = 128, bool grow_inside = false) {
  int color = 0;
  double num_fragments = 0, frag_len = 0;
  static FastRec R[STG_STEP_MAX];
  static FastDeps D[STG_STEP_MAX];
  int n = 0;
  while (n < S.nr && !S.err) {
    const int hi = n + W < S.nr ? n + W : S.nr;
    int k = 0;
    int ok[STG_STEP_MAX];
#if defined(STG_FP_COUNT_BAILS)
    int why[STG_STEP_MAX];
    for (int l = 0; n + l < hi; l++) {
      ok[l] = grow_inside ? classify_read<true>(S, n + l, n, hi, R[l], &D[l]) : classify_read<false>(S, n + l, n, hi, R[l]);
      why[l] = stg_fp_last;
    }
#else
    for (int l = 0; n + l < hi; l++)
      ok[l] = grow_inside ? classify_read<true>(S, n + l, n, hi, R[l], &D[l]) : classify_read<false>(S, n + l, n, hi, R[l]);
#endif
    // two reads of the step that continue into the SAME mate (a read with several pair links) see each other's
    // readgroup pushes: the later one has to wait
    for (int l = 1; n + l < hi; l++)
      for (int l2 = 0; l2 < l && (ok[l] == 1 || ok[l] == 3); l2++)
        if ((ok[l2] == 1 || (grow_inside && ok[l2] == 3)) && R[l].npc >= 0 && R[l2].npc == R[l].npc) ok[l] = 2;
    int stop;
    if (grow_inside) {
      // reads that grow a group stay inside the run; a read that took for granted a bound which an earlier read of the
      // step moves has to wait
      for (int l = 1; n + l < hi; l++)
        for (int l2 = 0; l2 < l && (ok[l] == 1 || ok[l] == 3); l2++)
          if (ok[l2] == 3 && grows_past(D[l], R[l2])) ok[l] = 2;
      while (n + k < hi && (ok[k] == 1 || ok[k] == 3)) k++;
      stop = n + k < hi ? ok[k] : 1;
    } else {
      while (n + k < hi && ok[k] == 1) k++;
      stop = n + k < hi ? ok[k] : 1;              // what ended the run
      if (stop == 3) k++;                         // a read that grows a group is the run's last read
    }
#if defined(STG_FP_COUNT_BAILS)
    stg_fp_bails[0]++;
    if (stop == 3 && !grow_inside) stg_fp_bails[60]++; else if (n + k < hi) stg_fp_bails[ok[k] == 2 ? 61 : (why[k] > 0 && why[k] < 60 ? why[k] : 62)]++; else stg_fp_bails[63]++;
#endif
    int minted = 0;
    for (int l = 0; l < k; l++) { commit_free(S, R[l], n + l, color + minted); minted += R[l].newcol; }
    for (int l = 0; l < k; l++) commit_ordered(S, R[l], num_fragments, frag_len);
    S.ncol += minted; color += minted;
    if (n_simple) *n_simple += k;
    n += k;
    if (n < S.nr && n < hi && stop == 0 && !S.err) { color = walk_read(S, n, color, num_fragments, frag_len); n++; }
  }
  *num_fragments_out = num_fragments; *frag_len_out = frag_len;
  if (S.err) return;
  walk_finish(S, jt_perm, jt_inv);
}

// ---- second half of a9: colours across the three lanes, then bundles ------------------------------------------
struct Grp2 {
  int ng, ncol;
  uint32_t bundledist;
  const int32_t* startg;
  // groups after the first half (read only)
  const uint32_t *g_start, *g_end; const int32_t* g_next; const float *g_cov, *g_multi;
  // working / output arrays of this bundle
  int32_t* color;       // [ng]   CGroup::color
  float* negp;          // [ng]   CGroup::neg_prop
  int32_t *eq, *eqneg, *eqpos, *bundlecol;   // [ncol]
  int32_t* g2b;         // [3*ng]
  int32_t *bun_len, *bun_start, *bun_last; float *bun_cov, *bun_multi;   // [3][ng]
  uint32_t *bn_s, *bn_e; float* bn_cov; int32_t* bn_next;                // [3][ng]
  int nbun[3], nbn[3];
};

STG_HD int get_min_start(const Grp2& S, const int* cur) {   // rlink.cpp:984-1026
  if (cur[0] != -1) {
    if (cur[1] != -1) {
      if (cur[2] != -1) {
        const int two = S.g_start[cur[0]] < S.g_start[cur[1]] ? 0 : 1;
        return S.g_start[cur[two]] < S.g_start[cur[2]] ? two : 2;
      }
      return S.g_start[cur[0]] < S.g_start[cur[1]] ? 0 : 1;
    }
    if (cur[2] != -1) return S.g_start[cur[0]] < S.g_start[cur[2]] ? 0 : 2;
    return 0;
  }
  if (cur[1] != -1) {
    if (cur[2] != -1) return S.g_start[cur[1]] < S.g_start[cur[2]] ? 1 : 2;
    return 1;
  }
  return 2;
}

// rlink.cpp:1028-1056: U = the unstranded group, G = the stranded one, grcol = G's colour
STG_HD void set_strandcol(int32_t* color, int U, int G, int grcol, int32_t* eqcol, int32_t* equalcolor) {
  int zerocol = eqcol[color[U]];
  if (zerocol > -1) {
    while (eqcol[zerocol] != -1 && eqcol[zerocol] != zerocol) zerocol = eqcol[zerocol];
    const int tmpcol = zerocol;
    while (equalcolor[zerocol] != zerocol) zerocol = equalcolor[zerocol];
    eqcol[color[U]] = zerocol;
    eqcol[tmpcol] = zerocol;
    if (zerocol < grcol) { equalcolor[grcol] = zerocol; color[G] = zerocol; }
    else if (grcol < zerocol) { equalcolor[zerocol] = grcol; eqcol[color[U]] = grcol; }
  } else eqcol[color[U]] = grcol;
}

STG_HD int bn_add(Grp2& S, int s, uint32_t a, uint32_t e, float cov) {
  const int k = S.nbn[s]++, o = s * S.ng + k;
  S.bn_s[o] = a; S.bn_e[o] = e; S.bn_cov[o] = cov; S.bn_next[o] = -1;
  return k;
}

STG_HD int create_bundle(Grp2& S, int s, int g) {   // rlink.cpp:1083-1095
  const int bid = bn_add(S, s, S.g_start[g], S.g_end[g], S.g_cov[g]);
  const int bno = S.nbun[s]++, o = s * S.ng + bno;
  S.bun_len[o] = (int32_t)(S.g_end[g] - S.g_start[g] + 1); S.bun_cov[o] = S.g_cov[g]; S.bun_multi[o] = S.g_multi[g];
  S.bun_start[o] = bid; S.bun_last[o] = bid;
  return bno;
}

STG_HD void add_group_to_bundle(Grp2& S, int s, int g, int bno) {   // rlink.cpp:1058-1081
  const int o = s * S.ng + bno;
  const int last = S.bun_last[o], lo = s * S.ng + last;
  if (S.g_start[g] > S.bn_e[lo] + S.bundledist) {
    const int bid = bn_add(S, s, S.g_start[g], S.g_end[g], S.g_cov[g]);
    S.bn_next[lo] = bid;
    S.bun_last[o] = bid;
    S.bun_len[o] += (int32_t)(S.g_end[g] - S.g_start[g] + 1);
    S.bun_cov[o] += S.g_cov[g];
    S.bun_multi[o] += S.g_multi[g];
  } else {
    if (S.bn_e[lo] < S.g_end[g]) { S.bun_len[o] += (int32_t)(S.g_end[g] - S.bn_e[lo]); S.bn_e[lo] = S.g_end[g]; }
    S.bun_cov[o] += S.g_cov[g];
    S.bun_multi[o] += S.g_multi[g];
    S.bn_cov[lo] += S.g_cov[g];
  }
}

STG_HD float ovl_share(const Grp2& S, int U, int G) {   // share of G's coverage that overlaps U: rlink.cpp:15230-15290
  const uint32_t maxstart = S.g_start[U] > S.g_start[G] ? S.g_start[U] : S.g_start[G];
  uint32_t minend = S.g_end[U] < S.g_end[G] ? S.g_end[U] : S.g_end[G];
  if (minend < maxstart) minend = maxstart;
  return S.g_cov[G] * (float)(minend - maxstart + 1u) / (float)(S.g_end[G] - S.g_start[G] + 1u);
}

// expects color[] = the groups' colours and eq[] = eqcol after the first half, eqneg/eqpos/bundlecol = -1, negp = 0,
// g2b = -1
STG_HDN inline void groups_second_half(Grp2& S) {
  const uint32_t bd = S.bundledist;
  int32_t* color = S.color; int32_t* eq = S.eq;
  int cur[3], prev[3] = {-1, -1, -1};
  for (int i = 0; i < 3; i++) cur[i] = S.startg[i];
  while (cur[0] != -1 || cur[1] != -1 || cur[2] != -1) {
    const int nextgr = get_min_start(S, cur);
    const int g = cur[nextgr];
    int grcol = color[g];
    while (eq[grcol] != grcol) grcol = eq[grcol];
    color[g] = grcol;
    if (nextgr == 1) {
      if (prev[0] != -1 && S.g_start[g] <= S.g_end[prev[0]] + bd) {
        set_strandcol(color, g, prev[0], color[prev[0]], S.eqneg, eq);
        S.negp[g] += ovl_share(S, g, prev[0]);
      }
      while (cur[0] != -1 && S.g_start[g] <= S.g_end[cur[0]] + bd && S.g_start[cur[0]] <= S.g_end[g] + bd) {
        int c0 = color[cur[0]];
        while (eq[c0] != c0) c0 = eq[c0];
        color[cur[0]] = c0;
        S.negp[cur[0]] = 1;
        set_strandcol(color, g, cur[0], color[cur[0]], S.eqneg, eq);
        S.negp[g] += ovl_share(S, g, cur[0]);
        prev[0] = cur[0];
        cur[0] = S.g_next[cur[0]];
      }
      float pos_prop = 0;
      if (prev[2] != -1 && S.g_start[g] <= S.g_end[prev[2]] + bd) {
        set_strandcol(color, g, prev[2], color[prev[2]], S.eqpos, eq);
        if (S.negp[g]) pos_prop += ovl_share(S, g, prev[2]);
      }
      while (cur[2] != -1 && S.g_start[g] <= S.g_end[cur[2]] + bd && S.g_start[cur[2]] <= S.g_end[g] + bd) {
        int c2 = color[cur[2]];
        while (eq[c2] != c2) c2 = eq[c2];
        color[cur[2]] = c2;
        set_strandcol(color, g, cur[2], color[cur[2]], S.eqpos, eq);
        if (S.negp[g]) pos_prop += ovl_share(S, g, cur[2]);
        prev[2] = cur[2];
        cur[2] = S.g_next[cur[2]];
      }
      if (pos_prop) S.negp[g] /= (S.negp[g] + pos_prop);
      else if (S.negp[g]) S.negp[g] = 1;
    } else if (nextgr == 0) S.negp[g] = 1;
    prev[nextgr] = g;
    cur[nextgr] = S.g_next[g];
  }

  for (int i = 0; i < 3; i++) { cur[i] = S.startg[i]; S.nbun[i] = 0; S.nbn[i] = 0; }
  while (cur[0] != -1 || cur[1] != -1 || cur[2] != -1) {
    const int nextgr = get_min_start(S, cur);
    const int g = cur[nextgr];
    int grcol = color[g];
    while (eq[grcol] != grcol) grcol = eq[grcol];
    color[g] = grcol;
    if (nextgr == 0 || nextgr == 2 || (S.eqneg[grcol] == -1 && S.eqpos[grcol] == -1)) {
      int bno = S.bundlecol[grcol];
      if (bno > -1) add_group_to_bundle(S, nextgr, g, bno);
      else { bno = create_bundle(S, nextgr, g); S.bundlecol[grcol] = bno; }
      S.g2b[nextgr * S.ng + g] = S.bun_last[nextgr * S.ng + bno];
    } else {
      if (S.eqneg[grcol] != -1) {
        int negcol = S.eqneg[grcol];
        while (eq[negcol] != negcol) negcol = eq[negcol];
        int bno = S.bundlecol[negcol];
        if (bno > -1) add_group_to_bundle(S, 0, g, bno);
        else { bno = create_bundle(S, 0, g); S.bundlecol[negcol] = bno; }
        S.g2b[0 * S.ng + g] = S.bun_last[0 * S.ng + bno];
      }
      if (S.eqpos[grcol] != -1) {
        int poscol = S.eqpos[grcol];
        while (eq[poscol] != poscol) poscol = eq[poscol];
        int bno = S.bundlecol[poscol];
        if (bno > -1) add_group_to_bundle(S, 2, g, bno);
        else { bno = create_bundle(S, 2, g); S.bundlecol[poscol] = bno; }
        S.g2b[2 * S.ng + g] = S.bun_last[2 * S.ng + bno];
      }
    }
    cur[nextgr] = S.g_next[g];
  }
}
