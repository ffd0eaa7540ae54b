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
// Mapping: the pass is a chain of sequential sweeps over ONE bundle's junction table (10 - 10^4 rows) whose state
// (running sums per start / end, "support so far", demotion pointers) carries from row to row, while the 60 000
// bundles of a batch are independent — so one thread walks one bundle and the batch supplies the parallelism.  All
// arithmetic is f64 / f32 exactly as in the reference (the `support` and `tolerance` temporaries are float).
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

__global__ void __launch_bounds__(128) junc_pass_kernel(DevBatchView v, JuncParams prm) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.n_bundles) return;
  const uint32_t j0 = v.b_junc_off[b];
  const int n = (int)(v.b_junc_off[b + 1] - j0);
  if (n == 0) return;
  const int64_t refstart = v.b_start[b];
  const int64_t len = v.b_len[b];
  const float* __restrict__ c1 = v.bpcov + 3 * (size_t)v.b_gbase[b] + v.b_stride[b];   // lane 1
  JView J;
  J.start = v.j_start + j0; J.end = v.j_end + j0; J.strand = v.jp_strand + j0; J.guide_match = v.j_guide_match + j0;
  J.nreads = v.jp_nreads + j0; J.nreads_good = v.jp_nreads_good + j0; J.left = v.jp_left + j0; J.right = v.jp_right + j0;
  J.mm = v.jp_mm + j0; J.nm = v.j_nm + j0; J.ej = v.jp_ej + j0;
  for (int i = 0; i < n; i++) {
    J.strand[i] = v.j_strand[j0 + i]; J.nreads[i] = v.j_nreads[j0 + i]; J.mm[i] = v.j_mm[j0 + i];
    J.nreads_good[i] = v.j_nreads_good[j0 + i]; J.left[i] = v.j_leftsupport[j0 + i]; J.right[i] = v.j_rightsupport[j0 + i];
    J.ej[i] = i;
  }
  sort_by_end(J, n);
#define E(k) J.ej[k]

  // ---- per-start / per-end support: rlink.cpp:14391-14531 (no genomic sequence: consleft = consright = -1) ----
  uint32_t start = 0, end = 0;
  double lsum[2] = {0, 0}, rsum[2] = {0, 0};
  bool higherr = false;
  for (int i = 0; i < n; i++) {
    if (!higherr && J.strand[i] && J.nm[i] == J.nreads[i] && !J.guide_match[i]) higherr = true;
    if (J.start[i] != start) {
      int j = i - 1;
      while (j >= 0 && J.start[j] == start) {
        if (J.strand[j]) J.left[j] = lsum[(1 + J.strand[j]) / 2];
        j--;
      }
      j = i + 1;
      if (J.strand[i])
        while (j < n && J.start[j] == J.start[i] && J.end[j] == J.end[i]) {
          if (J.strand[j] && J.strand[i] != J.strand[j]) {
            if (J.nreads[i] > J.nreads[j] && J.nreads_good[i] > J.nreads_good[j]) J.strand[j] = 0;
            if (J.nreads[i] < J.nreads[j] && J.nreads_good[i] < J.nreads_good[j]) J.strand[i] = 0;
          }
          j++;
        }
      lsum[0] = 0; lsum[1] = 0;
      if (J.strand[i]) lsum[(1 + J.strand[i]) / 2] = J.left[i];
      start = J.start[i];
    } else if (J.strand[i]) lsum[(1 + J.strand[i]) / 2] += J.left[i];

    const int ei = E(i);
    if (J.end[ei] != end) {
      int j = i - 1;
      while (j >= 0 && J.end[E(j)] == end) {
        const int ejj = E(j);
        if (J.strand[ejj]) J.right[ejj] = rsum[(1 + J.strand[ejj]) / 2];
        j--;
      }
      rsum[0] = 0; rsum[1] = 0;
      if (J.strand[ei]) rsum[(1 + J.strand[ei]) / 2] = J.right[ei];
      end = J.end[ei];
    } else if (J.strand[ei]) rsum[(1 + J.strand[ei]) / 2] += J.right[ei];
  }

  // ---- higherr: rlink.cpp:14538-14787, non-viral branch ----
  if (higherr) {
    const uint32_t juncsupport = prm.junctionsupport, sserror = prm.sserror;
    const float tolerance = (float)(1 - 0.1);
    for (int i = 1; i < n; i++) {
      if (J.strand[i]) {
        if (J.nm[i] != 0.0 && !J.guide_match[i] && J.nm[i] >= J.nreads[i]) {
          if (J.nreads_good[i] >= 0 && J.mm[i] >= 0) {
            if (J.nreads_good[i] < 1.25 * prm.junctionthr) J.mm[i] = -1;
            else {
              const uint32_t point = (uint32_t)((int64_t)J.start[i] - refstart);
              const double leftcov = cov_all(c1, point, point);
              const double rightcov = cov_all(c1, point + 1, point + 1);
              if (rightcov > tolerance * leftcov) J.mm[i] = -1;
            }
          }
          int j = i - 1;
          float support = 0;
          bool searchjunc = true, reliable = false;
          while (j > 0 && J.start[i] - J.start[j] < juncsupport) {
            if (J.strand[j] == J.strand[i]) {
              if (J.start[j] == J.start[i]) {
                if (J.nreads[j] < 0) {
                  const int point = (int)(-J.nreads[j]);
                  if (ok_to_demote(J, i, point)) { J.nreads[i] = J.nreads[j]; searchjunc = false; }
                }
                break;
              } else if (J.guide_match[j] || J.nm[j] < J.nreads[j]) {
                if (J.left[j] > J.left[i] * tolerance) {
                  if (ok_to_demote(J, i, j)) { reliable = true; J.nreads[i] = -j; support = (float)J.left[j]; break; }
                }
              } else if (J.left[j] > support && J.start[i] - J.start[j] < sserror && J.left[j] * tolerance > J.left[i]) {
                if (ok_to_demote(J, i, j)) { J.nreads[i] = -j; support = (float)J.left[j]; }
              }
            }
            j--;
          }
          if (searchjunc) {
            j = i + 1;
            int dist = (int)juncsupport;
            if (J.nreads[i] < 0) dist = (int)(J.start[i] - J.start[abs((int)J.nreads[i])]);
            while (j < n && J.start[j] - J.start[i] < juncsupport) {
              if (J.strand[j] == J.strand[i] && J.start[j] != J.start[i]) {
                const int d = (int)(J.start[j] - J.start[i]);
                if (J.guide_match[j] || J.nm[j] < J.nreads[j]) {
                  if ((d < dist || (d == dist && J.left[j] > support)) && J.left[j] > J.left[i] * tolerance) {
                    if (ok_to_demote(J, i, j)) { J.nreads[i] = -j; support = (float)J.left[j]; break; }
                  }
                } else if (!reliable && J.left[j] > support && (uint32_t)d < sserror && J.left[j] * tolerance > J.left[i]) {
                  if (ok_to_demote(J, i, j)) { J.nreads[i] = -j; support = (float)J.left[j]; }
                }
              }
              j++;
            }
          }
        }
      }
      const int ei = E(i);
      if (J.strand[ei]) {
        if (J.nm[ei] != 0.0 && !J.guide_match[ei] && J.nm[ei] >= J.nreads[ei]) {
          if (J.nreads_good[ei] >= 0 && J.mm[ei] >= 0) {
            if (J.nreads_good[ei] < 1.25 * prm.junctionthr) J.mm[ei] = -1;
            else {
              const uint32_t point = (uint32_t)((int64_t)J.end[ei] - 1 - refstart);
              const double leftcov = cov_all(c1, point, point);
              const double rightcov = cov_all(c1, point + 1, point + 1);
              if (leftcov > tolerance * rightcov) J.mm[ei] = -1;
            }
          }
          int j = i - 1;
          float support = 0;
          bool searchjunc = true, reliable = false;
          while (j > 0 && J.end[ei] - J.end[E(j)] < juncsupport) {
            const int ejj = E(j);
            if (J.strand[ejj] == J.strand[ei]) {
              if (J.end[ejj] == J.end[ei]) {
                if (J.nreads_good[ejj] < 0) {
                  const int point = (int)(-J.nreads_good[ejj]);
                  if (ok_to_demote(J, ei, E(point))) { J.nreads_good[ei] = J.nreads_good[ejj]; searchjunc = false; }
                }
                break;
              } else if (J.guide_match[ejj] || J.nm[ejj] < J.nreads[ejj]) {
                if (J.right[ejj] > J.right[ei] * tolerance) {
                  if (ok_to_demote(J, ei, ejj)) { reliable = true; J.nreads_good[ei] = -j; support = (float)J.right[ejj]; break; }
                }
              } else if (J.right[ejj] > support && J.end[ei] - J.end[ejj] < sserror && J.right[ejj] * tolerance > J.right[ei]) {
                if (ok_to_demote(J, ei, ejj)) { J.nreads_good[ei] = -j; support = (float)J.right[ejj]; }
              }
            }
            j--;
          }
          if (searchjunc) {
            j = i + 1;
            int dist = (int)juncsupport;
            if (J.nreads_good[ei] < 0) dist = (int)(J.end[ei] - J.end[E(abs((int)J.nreads_good[ei]))]);
            while (j < n && J.end[E(j)] - J.end[ei] < juncsupport) {
              const int ejj = E(j);
              if (J.strand[ejj] == J.strand[ei] && J.end[ejj] != J.end[ei]) {
                const int d = (int)(J.end[ejj] - J.end[ei]);
                if (J.guide_match[ejj] || J.nm[ejj] < J.nreads[ejj]) {
                  if ((d < dist || (d == dist && J.right[ejj] > support)) && J.right[ejj] > J.right[ei] * tolerance) {
                    if (ok_to_demote(J, ei, ejj)) { J.nreads_good[ei] = -j; support = (float)J.right[ejj]; break; }
                  }
                } else if ((!reliable && J.right[ejj] > support && J.end[ejj] - J.end[ei] < sserror &&
                            J.right[ejj] * tolerance > J.right[ei]) ||
                           ((int)(J.end[ejj] - J.end[ei]) < dist && (J.guide_match[ejj] || J.nm[ejj] < J.nreads[ejj]))) {
                  if (ok_to_demote(J, ei, ejj)) { J.nreads_good[ei] = -j; support = (float)J.right[ejj]; }
                }
              }
              j++;
            }
          }
        }
      }
    }
  }
#undef E

  // ---- good_junc verdict of every stranded row ----
  for (int i = 0; i < n; i++) {
    int8_t gs = J.strand[i], g = -1;
    if (J.strand[i]) g = good_junc(J, i, refstart, c1, len, prm, &gs) ? 1 : 0;
    v.jp_good[j0 + i] = g;
    v.jp_good_strand[j0 + i] = gs;
  }
}

cudaError_t launch_junc_pass(const DevBatchView& v, const JuncParams& p, cudaStream_t st) {
  if (v.n_bundles == 0 || v.n_juncs == 0) return cudaSuccess;
  junc_pass_kernel<<<(unsigned)((v.n_bundles + 127) / 128), 128, 0, st>>>(v, p);
  return cudaGetLastError();
}
