// tools/groups_hostcheck.cu — DEVELOPMENT TOOL, not part of libstgpu and never shipped.
// Steps through the per-bundle code of stringtie_b200/csrc/groups_device.cuh (rows a8 + a9) compiled for the HOST and
// compares every field with the oracle (oracle/liborc.so), so that the logic can be debugged on a box without a GPU
// before the kernels go to the B200.  Build + run: python tools/groups_hostcheck.py
#define STG_FP_COUNT_BAILS 1
#include "../stringtie_b200/csrc/groups_device.cuh"
long long stg_fp_bails[64];
int stg_fp_last;
extern "C" long long* hostcheck_bails() { return stg_fp_bails; }
#include "../oracle/stg_oracle.h"
#include <cstdio>
#include <cstring>
#include <vector>

template <class T>
static int cmp(const char* what, int64_t b, const T* got, const T* want, size_t n, const int8_t* mask = nullptr) {
  for (size_t i = 0; i < n; i++) {
    if (mask && !mask[i]) continue;
    if (memcmp(&got[i], &want[i], sizeof(T))) {
      fprintf(stderr, "bundle %lld: %s differs at %zu (got %g want %g)\n", (long long)b, what, i, (double)got[i], (double)want[i]);
      return 1;
    }
  }
  return 0;
}

extern "C" int hostcheck_groups(const stg_packed_batch* B, uint32_t junctionsupport, int32_t junctionthr, uint32_t bundledist,
                                int64_t b_first, int64_t b_last, int steps, int64_t* stats) {
  orc_params P = {junctionsupport, junctionthr, 0};
  JuncParams prm = {junctionsupport, 25, junctionthr, 0};
  int bad = 0;
  for (int64_t b = b_first; b < b_last && b < B->n_bundles; b++) {
    const int64_t len = orc_bundle_len(B, b);
    const uint32_t r0 = B->b_read_off[b], nr = B->b_read_off[b + 1] - r0;
    const uint32_t j0 = B->b_junc_off[b], nj = B->b_junc_off[b + 1] - j0;
    std::vector<float> cov(3 * (size_t)len, 0.f);
    std::vector<double> good(nj + 1, 0.0), left(nj + 1, 0.0), right(nj + 1, 0.0);
    orc_bundle(B, &P, b, cov.data(), good.data(), left.data(), right.data());
    std::vector<int8_t> jp_strand(nj + 1), jp_good(nj + 1), jp_gs(nj + 1);
    std::vector<double> jp_nreads(nj + 1), jp_ngood(nj + 1), jp_left(nj + 1), jp_right(nj + 1), jp_mm(nj + 1);
    std::vector<int32_t> jp_ej(nj + 1);
    if (nj)
      orc_junction_pass(B, &P, b, cov.data(), good.data(), left.data(), right.data(), jp_strand.data(), jp_nreads.data(),
                        jp_ngood.data(), jp_left.data(), jp_right.data(), jp_mm.data(), jp_ej.data(), jp_good.data(), jp_gs.data());
    orc_groups_out o;
    orc_groups_bundle(B, &P, b, cov.data(), good.data(), left.data(), right.data(), bundledist, &o);
    if (nr == 0) { orc_groups_free(&o); continue; }

    const uint32_t seg0 = B->r_seg_off[r0], pair0 = B->r_pair_off[r0];
    const uint32_t nseg = B->r_seg_off[r0 + nr] - seg0, npair = B->r_pair_off[r0 + nr] - pair0, nintr = nseg - nr, i0 = seg0 - r0;
    // capacities, as group_caps_kernel
    std::vector<uint32_t> rcap_off(nr + 1, 0);
    for (uint32_t n = 0; n < nr; n++) {
      const uint32_t ns = B->r_seg_off[r0 + n + 1] - B->r_seg_off[r0 + n];
      const uint32_t p0 = B->r_pair_off[r0 + n], np = B->r_pair_off[r0 + n + 1] - p0;
      uint32_t cap = (2 * np + 1) * ns;
      for (uint32_t p = 0; p < np; p++) {
        const int32_t m = B->pair_idx[p0 + p];
        if (m >= 0 && (uint32_t)m < nr) cap += B->r_seg_off[r0 + m + 1] - B->r_seg_off[r0 + m];
      }
      rcap_off[n + 1] = rcap_off[n] + cap;
    }
    const uint32_t R = rcap_off[nr];
    std::vector<int8_t> strand(B->r_strand + r0, B->r_strand + r0 + nr);
    std::vector<uint32_t> seg_s(B->seg_start + seg0, B->seg_start + seg0 + nseg), seg_e(B->seg_end + seg0, B->seg_end + seg0 + nseg);
    std::vector<int32_t> pair(B->pair_idx + pair0, B->pair_idx + pair0 + npair), rj(nintr + 1);
    for (uint32_t k = 0; k < nintr; k++) rj[k] = B->rjunc_idx[i0 + k] < 0 ? 0 : B->rjunc_idx[i0 + k];
    std::vector<int8_t> jt_strand(jp_strand), ap_strand(nintr + 1), ap_mm(nintr + 1), g_alive(R + 1);
    std::vector<double> jt_mm(jp_mm);
    std::vector<uint32_t> ap_start(nintr + 1), ap_end(nintr + 1), g_start(R + 1), g_end(R + 1), rg_cnt(nr + 1, 0);
    std::vector<int32_t> perm(nj + nintr + 1), inv(nj + nintr + 1), g_color(R + 1), g_next(R + 1), g_merge(R + 1), g_grid(R + 1),
        eqcol(R + 3 * nr + 1), rgv(R + 1);
    std::vector<float> g_cov(R + 1), g_multi(R + 1);
    Grp S;
    S.nr = (int)nr; S.seg0 = seg0; S.pair0 = pair0;
    S.r_start = B->r_start + r0; S.r_end = B->r_end + r0; S.r_len = B->r_len + r0; S.seg_off = B->r_seg_off + r0;
    S.pair_off = B->r_pair_off + r0; S.r_count = B->r_count + r0; S.pair_cnt = B->pair_cnt + pair0; S.r_nh = B->r_nh + r0;
    S.r_flags = B->r_flags + r0;
    S.strand = strand.data(); S.seg_s = seg_s.data(); S.seg_e = seg_e.data(); S.pair = pair.data(); S.rj = rj.data();
    S.j_start = B->j_start + j0; S.j_end = B->j_end + j0; S.j_nreads = jp_nreads.data(); S.j_good = jp_ngood.data();
    S.j_ej = jp_ej.data(); S.jp_good = jp_good.data(); S.jp_good_strand = jp_gs.data();
    S.jt_strand = jt_strand.data(); S.jt_mm = jt_mm.data();
    S.ap_start = ap_start.data(); S.ap_end = ap_end.data(); S.ap_strand = ap_strand.data(); S.ap_mm = ap_mm.data();
    S.nj0 = (int)nj; S.nJ = (int)nj; S.apcap = (int)nintr;
    S.g_start = g_start.data(); S.g_end = g_end.data(); S.g_color = g_color.data(); S.g_next = g_next.data();
    S.g_merge = g_merge.data(); S.g_grid = g_grid.data(); S.g_cov = g_cov.data(); S.g_multi = g_multi.data(); S.g_alive = g_alive.data();
    S.ng = 0; S.gcap = (int)R; S.eqcol = eqcol.data(); S.ncol = 0; S.ccap = (int)(R + 3 * nr);
    S.rg = rgv.data(); S.rg_cnt = rg_cnt.data(); S.rcap_off = rcap_off.data();
    S.curr.fill(-1); S.startg.fill(-1);
    S.junctionsupport = junctionsupport; S.bundledist = bundledist; S.refstart = B->b_start[b]; S.len = len;
    S.c1 = cov.data() + len; S.prm = prm; S.err = 0;
    double nf = 0, fl = 0;
    if (steps) groups_first_half_steps(S, perm.data(), inv.data(), &nf, &fl, stats, steps < 0 ? -steps : steps, steps < 0);   // < 0: growth inside runs   // the warp's schedule, emulated
    else groups_first_half(S, perm.data(), inv.data(), &nf, &fl);
    if (stats) stats[1] += nr;
    int e = 0;
    if (S.err) { fprintf(stderr, "bundle %lld: capacity error\n", (long long)b); e = 1; }
    if (!e && (S.ng != o.n_groups || S.ncol != o.n_colors || S.nJ != o.n_juncs)) {
      fprintf(stderr, "bundle %lld: counts differ: groups %d/%d colours %d/%d juncs %d/%d\n", (long long)b, S.ng, o.n_groups, S.ncol,
              o.n_colors, S.nJ, o.n_juncs);
      e = 1;
    }
    if (!e) {
      e |= cmp("r_strand", b, strand.data(), o.r_strand, nr);
      e |= cmp("pair_idx", b, pair.data(), o.pair_idx, npair);
      e |= cmp("seg_start", b, seg_s.data(), o.seg_start, nseg);
      e |= cmp("seg_end", b, seg_e.data(), o.seg_end, nseg);
      e |= cmp("rjunc", b, rj.data(), o.rjunc, nintr);
      const int32_t sg[3] = {S.startg.a, S.startg.b, S.startg.c};
      e |= cmp("startgroup", b, sg, o.startgroup, 3);
      e |= cmp("num_fragments", b, &nf, &o.num_fragments, 1);
      e |= cmp("frag_len", b, &fl, &o.frag_len, 1);
      const int8_t* al = o.g_alive;
      e |= cmp("g_alive", b, g_alive.data(), o.g_alive, S.ng);
      e |= cmp("g_start", b, g_start.data(), o.g_start, S.ng, al);
      e |= cmp("g_end", b, g_end.data(), o.g_end, S.ng, al);
      e |= cmp("g_color", b, g_color.data(), o.g_color, S.ng, al);
      e |= cmp("g_next", b, g_next.data(), o.g_next, S.ng, al);
      e |= cmp("g_cov", b, g_cov.data(), o.g_cov_sum, S.ng, al);
      e |= cmp("g_multi", b, g_multi.data(), o.g_multi, S.ng, al);
      e |= cmp("merge", b, g_merge.data(), o.merge, S.ng);
      e |= cmp("eqcol", b, eqcol.data(), o.eqcol, S.ncol);
      for (uint32_t n = 0; n < nr && !e; n++) {
        if (rg_cnt[n] != o.rg_off[n + 1] - o.rg_off[n]) {
          fprintf(stderr, "bundle %lld: rg count of read %u: got %u want %u; nseg %d npair %d mate %d strand %d nh %d\n", (long long)b, n, rg_cnt[n],
                  o.rg_off[n + 1] - o.rg_off[n], S.nseg(n), S.npair(n), S.npair(n) ? B->pair_idx[B->r_pair_off[r0 + n]] : -9, B->r_strand[r0 + n], B->r_nh[r0 + n]);
          e = 1; break; }
        e |= cmp("rg", b, rgv.data() + rcap_off[n], o.rg + o.rg_off[n], rg_cnt[n]);
      }
      // junction table in re-sorted order
      for (int k = 0; k < S.nJ && !e; k++) {
        const int r = perm[k];
        if (S.jstart(r) != o.j_start[k] || S.jend(r) != o.j_end[k] || S.jstrand(r) != o.j_strand[k] ||
            (S.jmm_neg(r) != (o.j_mm[k] < 0))) { fprintf(stderr, "bundle %lld: junction row %d differs\n", (long long)b, k); e = 1; }
      }
    }
    if (!e) {   // second half
      const int ng = S.ng, nc = S.ncol;
      std::vector<int32_t> color(g_color.begin(), g_color.begin() + ng), eq(eqcol.begin(), eqcol.begin() + nc), eqneg(nc + 1, -1),
          eqpos(nc + 1, -1), bundlecol(nc + 1, -1), g2b(3 * ng + 1, -1), bun_len(3 * ng + 1), bun_start(3 * ng + 1), bun_last(3 * ng + 1),
          bn_next(3 * ng + 1);
      std::vector<float> negp(ng + 1, 0.f), bun_cov(3 * ng + 1), bun_multi(3 * ng + 1), bn_cov(3 * ng + 1);
      std::vector<uint32_t> bn_s(3 * ng + 1), bn_e(3 * ng + 1);
      Grp2 T;
      T.ng = ng; T.ncol = nc; T.bundledist = bundledist; const int32_t sg2[3] = {S.startg.a, S.startg.b, S.startg.c};
      T.startg = sg2;
      T.g_start = g_start.data(); T.g_end = g_end.data(); T.g_next = g_next.data(); T.g_cov = g_cov.data(); T.g_multi = g_multi.data();
      T.color = color.data(); T.negp = negp.data(); T.eq = eq.data(); T.eqneg = eqneg.data(); T.eqpos = eqpos.data();
      T.bundlecol = bundlecol.data(); T.g2b = g2b.data();
      T.bun_len = bun_len.data(); T.bun_start = bun_start.data(); T.bun_last = bun_last.data(); T.bun_cov = bun_cov.data();
      T.bun_multi = bun_multi.data(); T.bn_s = bn_s.data(); T.bn_e = bn_e.data(); T.bn_cov = bn_cov.data(); T.bn_next = bn_next.data();
      if (ng) groups_second_half(T); else for (int s = 0; s < 3; s++) T.nbun[s] = T.nbn[s] = 0;
      const int8_t* al = o.g_alive;
      e |= cmp("g_color2", b, color.data(), o.g_color2, ng, al);
      e |= cmp("g_neg_prop", b, negp.data(), o.g_neg_prop, ng, al);
      e |= cmp("eqcol2", b, eq.data(), o.eqcol2, nc);
      e |= cmp("eqneg", b, eqneg.data(), o.eqnegcol, nc);
      e |= cmp("eqpos", b, eqpos.data(), o.eqposcol, nc);
      e |= cmp("g2b", b, g2b.data(), o.group2bundle, 3 * (size_t)ng);
      for (int s = 0; s < 3 && !e; s++) {
        if (T.nbun[s] != o.bun_off[s + 1] - o.bun_off[s] || T.nbn[s] != o.bn_off[s + 1] - o.bn_off[s]) {
          fprintf(stderr, "bundle %lld: lane %d bundle counts differ\n", (long long)b, s); e = 1; break;
        }
        e |= cmp("bun_len", b, bun_len.data() + s * ng, o.bun_len + o.bun_off[s], T.nbun[s]);
        e |= cmp("bun_start", b, bun_start.data() + s * ng, o.bun_start + o.bun_off[s], T.nbun[s]);
        e |= cmp("bun_last", b, bun_last.data() + s * ng, o.bun_last + o.bun_off[s], T.nbun[s]);
        e |= cmp("bun_cov", b, bun_cov.data() + s * ng, o.bun_cov + o.bun_off[s], T.nbun[s]);
        e |= cmp("bun_multi", b, bun_multi.data() + s * ng, o.bun_multi + o.bun_off[s], T.nbun[s]);
        e |= cmp("bn_start", b, bn_s.data() + s * ng, o.bn_start + o.bn_off[s], T.nbn[s]);
        e |= cmp("bn_end", b, bn_e.data() + s * ng, o.bn_end + o.bn_off[s], T.nbn[s]);
        e |= cmp("bn_cov", b, bn_cov.data() + s * ng, o.bn_cov + o.bn_off[s], T.nbn[s]);
        e |= cmp("bn_next", b, bn_next.data() + s * ng, o.bn_next + o.bn_off[s], T.nbn[s]);
      }
    }
    bad += e;
    orc_groups_free(&o);
  }
  return bad;
}
