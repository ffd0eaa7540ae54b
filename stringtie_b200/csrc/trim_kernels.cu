// trim_kernels.cu — sm_100a kernel for find_all_trims (rlink.cpp:2118-2419; part of SURVEY.md §8 row a10, and one of the
// range-query consumers of bpcov named in SURVEY.md §7 step 5), short-read path without point features.
//
// What it must reproduce: for every base i of a region the reference compares the mean signed coverage
// (get_cov_sign, rlink.cpp:1842: all reads minus the opposite strand) on the left and on the right of i — halves of the
// whole region when it is shorter than 301 bp, windows of CHI_WIN + CHI_THR = 150 bp otherwise, in four phases with
// their own normalisation and threshold — and keeps, with a small state machine walking the bases in order, the
// sharpest source / sink drop within 50 bp; the kept points are then re-validated against flanks of up to 300 bp
// (rejected points stay in the list with pos = 0).  All of it in f64 on f32 prefix sums, exactly as the reference.
//
// Mapping: the per-base window sums are independent O(1) range queries on the blocked prefix sums — the data-parallel
// bulk (the reference spends O(region length) get_cov_sign calls here).  One warp takes one region: 32 bases per step,
// one per lane; the few bases whose drop passes the threshold are found with a ballot and fed, in base order, to the
// state machine, which every lane runs redundantly on warp-uniform values (only lane 0 writes).  Regions are
// independent, so a batch of them fills the machine.
#include "stgpu_internal.h"

#define FULL 0xffffffffu

namespace {

// get_cov_sign(sno, a, b, bpcov), rlink.cpp:1842-1852; a, b bundle-relative
__device__ __forceinline__ double cov_sign(const float* __restrict__ all, const float* __restrict__ opp, uint32_t a, uint32_t b) {
  const int m = (int)((b + 1) / STG_BSIZE);
  const int k = (int)(a / STG_BSIZE);
  double cov = 0;
  for (int i = k + 1; i <= m; i++) cov += (double)__fsub_rn(all[(size_t)i * STG_BSIZE - 1], opp[(size_t)i * STG_BSIZE - 1]);
  float e = __fsub_rn(all[b + 1], all[a]);
  e = __fsub_rn(e, opp[b + 1]);
  e = __fadd_rn(e, opp[a]);
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

}  // namespace

__global__ void __launch_bounds__(256) find_trims_kernel(DevBatchView v, int64_t n, const int32_t* __restrict__ bundle,
                                                         const int8_t* __restrict__ sno_in, const uint32_t* __restrict__ start_in,
                                                         const uint32_t* __restrict__ end_in, const uint32_t* __restrict__ cap_off,
                                                         uint32_t* __restrict__ o_cnt, uint32_t* o_pos, float* o_ab, uint8_t* o_st) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const int lane = threadIdx.x & 31;
  const int32_t b = bundle[w];
  const int sno = sno_in[w];
  const uint32_t start = start_in[w], end = end_in[w];
  const uint32_t refstart = (uint32_t)v.b_start[b];
  const float* __restrict__ base = v.bpcov + 3 * (size_t)v.b_gbase[b];
  const uint32_t stride = v.b_stride[b];
  const float* __restrict__ all = base + stride;                          // lane 1
  const float* __restrict__ opp = base + (size_t)(2 - sno) * stride;      // the opposite strand
  uint32_t* t_pos = o_pos + cap_off[w];
  float* t_ab = o_ab + cap_off[w];
  uint8_t* t_st = o_st + cap_off[w];
  const uint32_t cap = cap_off[w + 1] - cap_off[w];
#define GCS(a, z) cov_sign(all, opp, (uint32_t)(a) - refstart, (uint32_t)(z) - refstart)
  TrimState S{0u, 0u, 0.f, false, 0.0};
  int len = (int)(end - start + 1);
  if (len < 50 /*CHI_THR*/) { if (lane == 0) o_cnt[w] = 0; return; }

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
    for (uint32_t ib = i0; ib < i1; ib += 32) {
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
      uint32_t m = __ballot_sync(FULL, kind != 0);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        consider(ib + (uint32_t)src, __shfl_sync(FULL, kind, src), __shfl_sync(FULL, thisdrop, src), __shfl_sync(FULL, ab, src));
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
    if (lane == 0) o_cnt[w] = S.cnt < cap ? S.cnt : cap;
    return;
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
  __syncwarp();
  const int ntp = (int)(S.cnt < cap ? S.cnt : cap);
  if (lane == 0) {
    o_cnt[w] = (uint32_t)ntp;
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
}

cudaError_t launch_find_trims(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* sno, const uint32_t* start,
                              const uint32_t* end, const uint32_t* cap_off, uint32_t* o_cnt, uint32_t* o_pos, float* o_ab,
                              uint8_t* o_st, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  find_trims_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(v, n, bundle, sno, start, end, cap_off, o_cnt, o_pos, o_ab, o_st);
  return cudaGetLastError();
}
