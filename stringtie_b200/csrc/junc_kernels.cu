// junc_kernels.cu — sm_100a kernel of the junction pass that opens build_graphs().
//
// What it must reproduce (reference, CPU, one bundle at a time; short-read path without --viral / --rseq / --ptf):
//   ejunction = junction re-sorted by (end, start)        rlink.cpp:14387-14389 (juncCmpEnd :1236-1244; the reference's
//                                                         own quicksort, gclib/GVec.hh:896-920 — NOT stable, so the
//                                                         order of equal keys is part of the result)
//   per-start / per-end support sums, opposite-strand     rlink.cpp:14391-14531
//   conflicts on equal coordinates, `higherr` detection
//   higherr pass (mm = -1, demotion to a better           rlink.cpp:14538-14787; ok_to_demote :362-370
//   neighbour: its index stored negated in nreads /
//   nreads_good)
//   good_junc on every stranded junction                  rlink.cpp:13967-14070 (the reference evaluates it lazily, per
//                                                         read, in the loop at :14845; the verdict only depends on the
//                                                         junction row and bpcov, so it is computed once per row here)
//
// Mapping: ONE WARP per bundle (the batch supplies 60 000 of them).  Most of the pass is independent per row or per run
// of rows with equal start / equal end, and that part runs with a lane per row or per run:
//   * ejunction: where no two rows share (end, start) the sorted order is unique, so every row finds its rank by counting
//     smaller keys (all lanes read the same key at the same time); only tables with equal keys — whose order under the
//     reference's unstable quicksort is part of the result, because demotions store POSITIONS in ejunction — are sorted by
//     lane 0 with that quicksort;
//   * the opposite-strand conflicts touch the rows of one run of equal (start, end): a lane per run head;
//   * the per-start / per-end sums are added in row order by the lane that owns the run.  The reference interleaves the two
//     sweeps in one loop, so the end-side at step i sees a strand zeroed by the start-side only if that happened at a step
//     <= i: the strand "as of step t" is the final one if the row's run head is <= t, else the input one (strand_at);
//   * higherr: a step of the reference's loop changes state only for rows that are candidates (stranded, all reads with
//     mismatches, no guide) — a predicate of the INPUT columns.  Lanes evaluate it for 32 steps at once; lane 0 replays only
//     the candidate steps, start-side before end-side, in step order (demote_search): the demotion pointers written into
//     nreads / nreads_good are read by later steps, which is the one true sequential dependency of the pass;
//   * good_junc: a lane per row.
// All arithmetic is f64 / f32 exactly as in the reference (the `support` and `tolerance` temporaries are float).
#include "groups_device.cuh"

namespace {

struct JView {   // one bundle's rows of the junction-pass state
  const uint32_t* start; const uint32_t* end;
  int8_t* strand; const int8_t* guide_match;
  double *nreads, *nreads_good, *left, *right, *mm;
  const double* nm;
  int32_t* ej;
};

__device__ __forceinline__ int cmp_end(const JView& J, int a, int b) {   // juncCmpEnd, rlink.cpp:1236-1244
  const uint32_t ea = J.end[a], eb = J.end[b];
  if (ea < eb) return -1;
  if (ea > eb) return 1;
  const uint32_t sa = J.start[a], sb = J.start[b];
  if (sa < sb) return -1;
  if (sa > sb) return 1;
  return 0;
}

// gclib/GVec.hh:896-920 (Hoare partition, middle pivot) with an explicit stack: the sub-ranges produced by a
// partition are disjoint, so the order in which they are processed does not change the resulting permutation.
__device__ void sort_by_end(const JView& J, int n) {
  int stack[2 * 48];
  int sp = 0;
  if (n < 2) return;
  stack[sp++] = 0; stack[sp++] = n - 1;
  while (sp) {
    const int R = stack[--sp], L = stack[--sp];
    int I = L, Jx = R;
    const int P = J.ej[L + ((R - L) >> 1)];
    do {
      while (cmp_end(J, J.ej[I], P) < 0) I++;
      while (cmp_end(J, J.ej[Jx], P) > 0) Jx--;
      if (I <= Jx) {
        const int t = J.ej[I]; J.ej[I] = J.ej[Jx]; J.ej[Jx] = t;
        I++; Jx--;
      }
    } while (I <= Jx);
    // larger part first (processed last): the stack stays within log2(n) entries
    const bool left_todo = L < Jx, right_todo = I < R;
    if (left_todo && right_todo) {
      if (Jx - L > R - I) { stack[sp++] = L; stack[sp++] = Jx; stack[sp++] = I; stack[sp++] = R; }
      else { stack[sp++] = I; stack[sp++] = R; stack[sp++] = L; stack[sp++] = Jx; }
    } else if (left_todo) { stack[sp++] = L; stack[sp++] = Jx; }
    else if (right_todo) { stack[sp++] = I; stack[sp++] = R; }
  }
}

__device__ __forceinline__ double ratio_threshold(double na) {   // rlink.cpp:354-359
  if (na <= 25.0) return 1.0;
  else if (na <= 100.0) return 0.50;
  else if (na <= 200.0) return 0.25;
  else return 0.10;
}

__device__ __forceinline__ bool ok_to_demote(const JView& J, int cur, int better) {   // rlink.cpp:362-370
  const double nbr = J.nreads[better];
  const double nb = nbr > 0.0 ? nbr : 0.0;
  if (nb <= 0.0) return false;
  const double nar = J.nreads[cur];
  const double na = nar > 0.0 ? nar : 0.0;
  return (na / nb) <= ratio_threshold(na);
}

// good_junc, rlink.cpp:13967-14070 (good_junc_row in groups_device.cuh, shared with the read loop of rows a8 + a9)
__device__ bool good_junc(const JView& J, int i, int64_t refstart, const float* __restrict__ c1, int64_t len,
                          const JuncParams& p, int8_t* strand_out) {
  return good_junc_row(J.start[i], J.end[i], J.strand[i], J.guide_match[i], J.nreads[i], J.nreads_good[i], J.nm[i], J.left[i],
                       J.right[i], refstart, c1, len, p, strand_out);
}

}  // namespace

// strand of row r as the reference's interleaved loop sees it at step t (see the header): zeroing happens at the step of
// the row's run head
__device__ __forceinline__ int8_t strand_at(const JView& J, const int8_t* __restrict__ strand_in, int r, int t) {
  int h = r;
  while (h > 0 && J.start[h - 1] == J.start[h]) h--;
  return h <= t ? J.strand[r] : strand_in[r];
}

// One candidate step of the higherr pass (rlink.cpp:14596-14667 start side, :14669-14786 end side).  SIDE 0 looks along the
// table (position k = row k, coordinate = start, support = leftsupport, pointer column = nreads), SIDE 1 along ejunction
// (position k = row ej[k], coordinate = end, support = rightsupport, pointer column = nreads_good); pointers are positions.
template <int SIDE>
__device__ void demote_search(const JView& J, int n, int i, const JuncParams& prm, int64_t refstart, const float* __restrict__ c1) {
  auto row = [&](int k) { return SIDE ? J.ej[k] : k; };
  auto pos = [&](int r) { return SIDE ? J.end[r] : J.start[r]; };
  auto sup = [&](int r) { return SIDE ? J.right[r] : J.left[r]; };
  double* ptr = SIDE ? J.nreads_good : J.nreads;
  const int me = row(i);
  const uint32_t juncsupport = prm.junctionsupport, sserror = prm.sserror;
  const float tolerance = (float)(1 - 0.1);
  if (J.nreads_good[me] >= 0 && J.mm[me] >= 0) {   // all reads mismatched and the coverage does not drop across the site
    if (J.nreads_good[me] < 1.25 * prm.junctionthr) J.mm[me] = -1;
    else {
      const uint32_t point = (uint32_t)((int64_t)pos(me) - (SIDE ? 1 : 0) - refstart);
      const double leftcov = cov_all(c1, point, point), rightcov = cov_all(c1, point + 1, point + 1);
      if (SIDE ? leftcov > tolerance * rightcov : rightcov > tolerance * leftcov) J.mm[me] = -1;
    }
  }
  const auto trusted = [&](int r) { return J.guide_match[r] || J.nm[r] < J.nreads[r]; };
  float support = 0;
  bool searchjunc = true, reliable = false;
  for (int j = i - 1; j > 0; j--) {                // neighbours below, nearest first
    const int r = row(j);
    if (pos(me) - pos(r) >= juncsupport) break;
    if (J.strand[r] != J.strand[me]) continue;
    if (pos(r) == pos(me)) {                       // same site: follow what that row decided
      if (ptr[r] < 0) {
        const int to = (int)(-ptr[r]);
        if (ok_to_demote(J, me, row(to))) { ptr[me] = ptr[r]; searchjunc = false; }
      }
      break;
    }
    if (trusted(r)) {
      if (sup(r) > sup(me) * tolerance && ok_to_demote(J, me, r)) { reliable = true; ptr[me] = -j; support = (float)sup(r); break; }
    } else if (sup(r) > support && pos(me) - pos(r) < sserror && sup(r) * tolerance > sup(me)) {
      if (ok_to_demote(J, me, r)) { ptr[me] = -j; support = (float)sup(r); }
    }
  }
  if (!searchjunc) return;
  int dist = (int)juncsupport;
  if (ptr[me] < 0) dist = (int)(pos(me) - pos(row(abs((int)ptr[me]))));
  for (int j = i + 1; j < n; j++) {                // neighbours above
    const int r = row(j);
    if (pos(r) - pos(me) >= juncsupport) break;
    if (J.strand[r] != J.strand[me] || pos(r) == pos(me)) continue;
    const int d = (int)(pos(r) - pos(me));
    if (trusted(r)) {
      if ((d < dist || (d == dist && sup(r) > support)) && sup(r) > sup(me) * tolerance) {
        if (ok_to_demote(J, me, r)) { ptr[me] = -j; support = (float)sup(r); break; }
      }
      // (the end side's second clause, rlink.cpp:14775, needs trusted(r) false: it can never hold here)
    } else if (!reliable && sup(r) > support && (uint32_t)d < sserror && sup(r) * tolerance > sup(me)) {
      if (ok_to_demote(J, me, r)) { ptr[me] = -j; support = (float)sup(r); }
    }
  }
}

#define JUNC_WARPS 4
#define JUNC_RANK_MAX 2048   // tables up to this many rows are ranked by counting; longer ones go to the quicksort
__global__ void __launch_bounds__(32 * JUNC_WARPS) junc_pass_kernel(DevBatchView v, JuncParams prm) {
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int64_t b = (int64_t)blockIdx.x * JUNC_WARPS + (threadIdx.x >> 5);
  if (b >= v.n_bundles) return;                     // (uniform over the warp)
  const uint32_t j0 = v.b_junc_off[b];
  const int n = (int)(v.b_junc_off[b + 1] - j0);
  if (n == 0) return;
  const int64_t refstart = v.b_start[b];
  const int64_t len = v.b_len[b];
  const float* __restrict__ c1 = v.bpcov + v.cov_pitch + v.b_gbase[b];   // lane 1
  JView J;
  J.start = v.j_start + j0; J.end = v.j_end + j0; J.strand = v.jp_strand + j0; J.guide_match = v.j_guide_match + j0;
  J.nreads = v.jp_nreads + j0; J.nreads_good = v.jp_nreads_good + j0; J.left = v.jp_left + j0; J.right = v.jp_right + j0;
  J.mm = v.jp_mm + j0; J.nm = v.j_nm + j0; J.ej = v.jp_ej + j0;
  const int8_t* __restrict__ strand_in = v.j_strand + j0;
  const double* __restrict__ nreads_in = v.j_nreads + j0;

  // ---- working copy of the table; ejunction --------------------------------------------------------------------------
  for (int i = lane; i < n; i += 32) {
    J.strand[i] = strand_in[i]; J.nreads[i] = nreads_in[i]; J.mm[i] = v.j_mm[j0 + i];
    J.nreads_good[i] = v.j_nreads_good[j0 + i]; J.left[i] = v.j_leftsupport[j0 + i]; J.right[i] = v.j_rightsupport[j0 + i];
  }
  bool tie = n > JUNC_RANK_MAX;
  if (!tie) {
    for (int r0 = 0; r0 < n; r0 += 32) {
      const int r = r0 + lane;
      const uint64_t key = r < n ? ((uint64_t)J.end[r] << 32 | J.start[r]) : ~0ull;
      int rank = 0;
      bool same = false;
      for (int q = 0; q < n; q++) {
        const uint64_t kq = (uint64_t)J.end[q] << 32 | J.start[q];
        rank += kq < key;
        same |= kq == key && q != r;
      }
      tie = __any_sync(FULL, same && r < n);
      if (tie) break;
      if (r < n) J.ej[rank] = r;
    }
  }
  __syncwarp();
  if (tie) {
    for (int i = lane; i < n; i += 32) J.ej[i] = i;
    __syncwarp();
    if (lane == 0) sort_by_end(J, n);
  }
  __syncwarp();

  // ---- opposite-strand conflicts on equal coordinates (rlink.cpp:14441-14452): a lane per run head ----------------------
  for (int i = lane; i < n; i += 32) {
    const bool head = i == 0 ? J.start[0] != 0 : J.start[i] != J.start[i - 1];
    if (!head || !J.strand[i]) continue;
    for (int j = i + 1; j < n && J.start[j] == J.start[i] && J.end[j] == J.end[i]; j++)
      if (J.strand[j] && J.strand[i] != J.strand[j]) {
        if (J.nreads[i] > J.nreads[j] && J.nreads_good[i] > J.nreads_good[j]) J.strand[j] = 0;
        if (J.nreads[i] < J.nreads[j] && J.nreads_good[i] < J.nreads_good[j]) J.strand[i] = 0;
      }
  }
  __syncwarp();

  // ---- higherr (the reference tests row i at the top of step i: its own run head has not been through the conflicts yet) --
  bool he = false;
  for (int i = lane; i < n; i += 32) {
    const int8_t s = strand_at(J, strand_in, i, i - 1);
    he |= s && J.nm[i] == nreads_in[i] && !J.guide_match[i];
  }
  const bool higherr = __any_sync(FULL, he);

  // ---- per-start sums of leftsupport (rlink.cpp:14391-14470): the lane that owns a run adds it in row order and hands the
  //      sum to the run's rows — except the last run, which the reference's loop leaves as it is -------------------------
  for (int i = lane; i < n; i += 32) {
    if (i != 0 && J.start[i] == J.start[i - 1]) continue;
    int e = i + 1;
    while (e < n && J.start[e] == J.start[i]) e++;
    if (e == n) continue;
    double ls[2] = {0, 0};
    for (int j = i; j < e; j++) if (J.strand[j]) ls[(1 + J.strand[j]) / 2] += J.left[j];
    for (int j = i; j < e; j++) if (J.strand[j]) J.left[j] = ls[(1 + J.strand[j]) / 2];
  }
  // ---- per-end sums of rightsupport along ejunction (rlink.cpp:14472-14531) ---------------------------------------------
  for (int i = lane; i < n; i += 32) {
    if (i != 0 && J.end[J.ej[i]] == J.end[J.ej[i - 1]]) continue;
    int e = i + 1;
    while (e < n && J.end[J.ej[e]] == J.end[J.ej[i]]) e++;
    if (e == n) continue;
    double rs[2] = {0, 0};
    for (int k = i; k < e; k++) {
      const int r = J.ej[k];
      const int8_t s = strand_at(J, strand_in, r, k);
      if (s) rs[(1 + s) / 2] += J.right[r];
    }
    for (int k = i; k < e; k++) {
      const int r = J.ej[k];
      const int8_t s = strand_at(J, strand_in, r, e);
      if (s) J.right[r] = rs[(1 + s) / 2];
    }
  }
  __syncwarp();

  // ---- higherr pass (rlink.cpp:14538-14787, non-viral branch): the candidate steps, in step order ------------------------
  if (higherr) {
    auto cand = [&](int r) { return J.strand[r] && J.nm[r] != 0.0 && !J.guide_match[r] && J.nm[r] >= nreads_in[r]; };
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const bool in = i >= 1 && i < n;
      const unsigned ms = __ballot_sync(FULL, in && cand(i)), me = __ballot_sync(FULL, in && cand(J.ej[i]));
      if (lane == 0) {
        unsigned m = ms | me;
        while (m) {
          const int p = __ffs(m) - 1;
          m &= m - 1;
          if (ms >> p & 1u) demote_search<0>(J, n, base + p, prm, refstart, c1);
          if (me >> p & 1u) demote_search<1>(J, n, base + p, prm, refstart, c1);
        }
      }
      __syncwarp();
    }
  }

  // ---- good_junc verdict of every stranded row: a lane per row ---------------------------------------------------------
  for (int i = lane; i < n; i += 32) {
    int8_t gs = J.strand[i], g = -1;
    if (J.strand[i]) g = good_junc(J, i, refstart, c1, len, prm, &gs) ? 1 : 0;
    v.jp_good[j0 + i] = g;
    v.jp_good_strand[j0 + i] = gs;
  }
}

cudaError_t launch_junc_pass(const DevBatchView& v, const JuncParams& p, cudaStream_t st) {
  if (v.n_bundles == 0 || v.n_juncs == 0) return cudaSuccess;
  junc_pass_kernel<<<(unsigned)((v.n_bundles + JUNC_WARPS - 1) / JUNC_WARPS), 32 * JUNC_WARPS, 0, st>>>(v, p);
  return cudaGetLastError();
}
