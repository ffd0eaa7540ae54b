// trims_device.cuh — find_all_trims (rlink.cpp:2118-2419) for ONE region, run by one warp (see trim_kernels.cu for the
// mapping).  Shared by find_trims_kernel (stg_find_trims, the neutral bundles) and by the graph builder (asm_device.cuh:
// trimnode_all, rlink.cpp:2809).  Written against warp_prims.cuh, so tools/asm_hostcheck.cu can step through it on a CPU.
#pragma once
#include "stgpu_internal.h"
#include "warp_prims.cuh"

// get_cov_sign(sno, a, b, bpcov), rlink.cpp:1842-1852; a, b bundle-relative
STG_HD double trims_cov_sign(const float* __restrict__ all, const float* __restrict__ opp, uint32_t a, uint32_t b) {
  const int m = (int)((b + 1) / STG_BSIZE);
  const int k = (int)(a / STG_BSIZE);
  double cov = 0;
  for (int i = k + 1; i <= m; i++) cov += (double)wp_fsub(all[(size_t)i * STG_BSIZE - 1], opp[(size_t)i * STG_BSIZE - 1]);
  float e = wp_fsub(all[b + 1], all[a]);
  e = wp_fsub(e, opp[b + 1]);
  e = wp_fadd(e, opp[a]);
  cov += (double)e;
  if (cov < 0) cov = 0;
  return cov;
}

struct TrimState {   // warp-uniform
  uint32_t cnt, last_pos;
  float last_ab;
  bool last_start;
  double lastdrop;
};

// Trim points of the region [start, end] (genomic) on the strand whose opposite lane is `opp`; up to `cap` points are
// written by lane 0 to t_pos / t_ab / t_st (rejected points keep pos = 0); returns their number (warp-uniform).  The
// caller must wp_sync() before other lanes read the arrays.
STG_HDN inline uint32_t find_trims_warp(const float* __restrict__ all, const float* __restrict__ opp, uint32_t refstart, uint32_t start,
                                 uint32_t end, uint32_t* t_pos, float* t_ab, uint8_t* t_st, uint32_t cap) {
  const int lane = wp_lane();
  #define GCS(a, z) trims_cov_sign(all, opp, (uint32_t)(a) - refstart, (uint32_t)(z) - refstart)
  TrimState S{0u, 0u, 0.f, false, 0.0};
  int len = (int)(end - start + 1);
  if (len < 50 /*CHI_THR*/) return 0u;

  // one candidate (rlink.cpp:2157-2187 and its siblings in the other three loops); arguments are warp-uniform
  auto consider = [&](uint32_t i, int kind, double thisdrop, float ab) {
    if (kind == 1) {   // source: coverage rises to the right of i
      if (!S.cnt || (!S.last_start && i + 1 - (uint32_t)(int)S.last_pos > 50u)) {
        S.cnt++; S.last_pos = i + 1; S.last_ab = ab; S.last_start = true; S.lastdrop = thisdrop;
      } else if (thisdrop < S.lastdrop) {
        S.lastdrop = thisdrop; S.last_pos = i + 1; S.last_ab = ab; S.last_start = true;
      } else return;
    } else {           // sink: coverage falls to the right of i
      if (!S.cnt || (S.last_start && i - (uint32_t)(int)S.last_pos > 50u)) {
        S.cnt++; S.last_pos = i; S.last_ab = ab; S.last_start = false; S.lastdrop = thisdrop;
      } else if (thisdrop < S.lastdrop) {
        S.lastdrop = thisdrop; S.last_pos = i; S.last_ab = ab; S.last_start = false;
      } else return;
    }
    if (lane == 0 && S.cnt <= cap) { t_pos[S.cnt - 1] = S.last_pos; t_ab[S.cnt - 1] = S.last_ab; t_st[S.cnt - 1] = S.last_start ? 1 : 0; }
  };
  // 32 bases per step: each lane evaluates one, the ones past the threshold are replayed in order
  auto phase = [&](uint32_t i0, uint32_t i1, auto&& covs, double localdrop, double norm, bool plus_one) {
    for (uint32_t ib = i0; ib < i1; ib += WP_N) {
      const uint32_t i = ib + (uint32_t)lane;
      int kind = 0;
      double thisdrop = 0.0;
      float ab = 0.f;
      if (i < i1) {
        double cl, cr;
        covs(i, cl, cr);
        if (cl < cr) { thisdrop = cl / cr; kind = 1; ab = (float)((cr - cl) / norm); }
        else if (cr != cl) { thisdrop = plus_one ? (cr + 1) / (cl + 1) : cr / cl; kind = 2; ab = (float)((cl - cr) / norm); }
        if (!(thisdrop < localdrop)) kind = 0;
      }
      uint32_t m = wp_ballot(kind != 0);
      while (m) {
        const int src = wp_ffs(m) - 1;
        m &= m - 1;
        consider(ib + (uint32_t)src, wp_shfl(kind, src), wp_shfl(thisdrop, src), wp_shfl(ab, src));
      }
    }
  };

  if (len < 2 * (100 + 50) + 1) {
    double localdrop = 0.1 / 0.5;                        // ERROR_PERC / DROP
    if (len < 100) localdrop = 0.1 / (10 * 0.5);         // very sharp drop only
    S.lastdrop = localdrop;
    phase(start + 25u, end - 25u, [&](uint32_t i, double& cl, double& cr) {
      cl = GCS(start, i - 1) / (double)(i - start);
      cr = GCS(i, end) / (double)(end - i + 1);
    }, localdrop, 0.5, true);
    return S.cnt < cap ? S.cnt : cap;
  }
  const uint32_t winlen = 150;
  double localdrop = 0.1 / 0.5;
  S.lastdrop = localdrop;
  phase(start + 50u - 1u, start + winlen - 1u, [&](uint32_t i, double& cl, double& cr) {
    cl = GCS(start, i) / (double)(i - start);
    cr = GCS(i + 1, i + winlen) / (double)winlen;
  }, localdrop, 0.5, false);
  if (!S.cnt) S.lastdrop = localdrop;
  const uint32_t len2 = (uint32_t)(len + (int)(start - winlen));   // = end + 1 - winlen
  phase(start + winlen - 1u, len2, [&](uint32_t i, double& cl, double& cr) {
    cl = GCS(i - winlen + 1, i);
    cr = GCS(i + 1, i + winlen);
  }, localdrop, 0.5 * winlen, false);
  phase(len2, end - 50u - 1u, [&](uint32_t i, double& cl, double& cr) {
    cl = GCS(i - winlen + 1, i) / (double)winlen;
    cr = GCS(i + 1, end) / (double)(end - i);
  }, 0.1, 0.5, false);
  wp_sync();
  const int ntp = (int)(S.cnt < cap ? S.cnt : cap);
  if (lane == 0) {
    // global re-validation (rlink.cpp:2351-2404): sequential over the handful of kept points
    localdrop = 0.1 / 0.5;
    uint32_t laststart = start;
    for (int i = 0; i < ntp; i++) {
      uint32_t midpos = t_pos[i];
      if (t_st[i]) midpos--;
      if (midpos - laststart > 300u) laststart = midpos - 300u;
      double covleft = GCS(laststart, midpos) / (double)(midpos - laststart + 1);
      uint32_t endpos = end;
      if (i < ntp - 1) endpos = t_pos[i + 1];
      if (endpos - midpos > 300u) endpos = midpos + 300u;
      double covright = GCS(midpos + 1, endpos) / (double)(endpos - midpos);
      if ((t_st[i] && covleft < localdrop * covright) || (covright < localdrop * covleft)) {
        laststart = t_pos[i] + 1;
      } else {
        t_pos[i] = 0;
        int k = i - 1;
        while (k >= 0) {
          while (k >= 0 && !t_pos[k]) k--;
          if (k >= 0) {
            midpos = t_pos[k];
            if (t_st[k]) midpos--;
            uint32_t leftstart = start;
            int j = k - 1;
            while (j >= 0) { if (t_pos[j]) { leftstart = t_pos[j]; break; } j--; }
            if (midpos - leftstart > 300u) leftstart = midpos - 300u;
            covleft = GCS(leftstart, midpos) / (double)(midpos - leftstart + 1);
            if (endpos - midpos > 300u) endpos = midpos + 300u;
            covright = GCS(midpos + 1, endpos) / (double)(endpos - midpos);
            if ((t_st[k] && covleft < localdrop * covright) || (covright < localdrop * covleft)) break;
            else t_pos[k] = 0;
            k--;
          }
        }
      }
    }
  }
#undef GCS
  return (uint32_t)ntp;
}
