// neutral_kernels.cu — SURVEY.md §8 row a19, de novo part: the single-exon predictions build_graphs makes for its neutral
// bundles (rlink.cpp:15500-15634; no guides, rawreads or long reads).  A neutral bundle — strand lane 1: groups that
// touch no stranded group — whose coverage is not mostly multi-mapped is cut, bundlenode by bundlenode, at the trim
// points of find_all_trims (rlink.cpp:2118; kernel find_trims, trim_kernels.cu); every piece longer than
// mintranscriptlen becomes a prediction with the mean coverage get_cov (rlink.cpp:1830) gives on the all-strands lane.
// Everything is read from what stg_build_groups left on the device:
//   neutral_regions   one thread per bundle of the batch: its neutral bundlenodes, in the order the reference visits them
//   find_trims        one warp per bundlenode (existing kernel)
//   neutral_preds     one thread per bundle: pieces, coverage, gene numbers (a counter that runs through the bundle)
//   neutral_compact   one thread per bundle: predictions into dense arrays
#include "groups_device.cuh"

namespace {

// the lane-1 bundles / bundlenodes of batch bundle b in the dense arrays of GroupsView
struct Lane1 { size_t base; int nbun; };
__device__ __forceinline__ Lane1 lane1_of(const GroupsView& g, int64_t b) {
  const size_t g0 = g.b_ng[b];
  const int ng = (int)(g.b_ng[b + 1] - g.b_ng[b]);
  return Lane1{3 * g0 + (size_t)ng, (int)g.b_nbun[3 * b + 1]};
}

__global__ void __launch_bounds__(128) neutral_count_kernel(DevBatchView v, GroupsView g, uint32_t* reg_off) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > v.n_bundles) return;
  reg_off[b] = b < v.n_bundles ? g.b_nbn[3 * b + 1] : 0u;   // upper bound: every neutral bundlenode of the bundle
}

// reg_off: exclusive prefix of the bounds.  Regions of bundles that do not qualify are not emitted; the unused slots of
// a bundle's range are made empty (end < start: find_trims returns nothing for them).
__global__ void __launch_bounds__(128) neutral_regions_kernel(DevBatchView v, GroupsView g, NeutralView nv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.n_bundles) return;
  const Lane1 L = lane1_of(g, b);
  uint32_t r = nv.reg_off[b];
  const uint32_t r_end = nv.reg_off[b + 1];
  for (int q = 0; q < L.nbun; q++) {
    const float bc = g.bun_cov[L.base + q], bm = g.bun_multi[L.base + q];
    if (!(bc != 0.f && (double)__fdiv_rn(bm, bc) <= (double)nv.mcov * (1 - 0.1 /*ERROR_PERC*/))) continue;   // rlink.cpp:15509
    bool first = true;
    for (int k = g.bun_start[L.base + q]; k != -1 && r < r_end; k = g.bn_next[L.base + k], r++) {
      const uint32_t s = g.bn_start[L.base + k], e = g.bn_end[L.base + k];
      nv.reg_bundle[r] = (int32_t)b; nv.reg_sno[r] = 0; nv.reg_start[r] = s; nv.reg_end[r] = e;
      nv.reg_first[r] = first ? 1 : 0;
      nv.cap_off[r] = e < s ? 0u : (e - s + 1u) / 50u + 2u;   // stg_trims_capacity
      first = false;
    }
  }
  nv.reg_cnt[b] = r - nv.reg_off[b];
  for (; r < r_end; r++) {
    nv.reg_bundle[r] = (int32_t)b; nv.reg_sno[r] = 0; nv.reg_start[r] = 1; nv.reg_end[r] = 0; nv.reg_first[r] = 0; nv.cap_off[r] = 0;
  }
  if (b == 0) nv.cap_off[nv.n_regions] = 0;
}

// cap_off: exclusive prefix.  A region with c trim points yields at most c + 1 predictions: region r writes its
// predictions from slot cap_off[r] + r on; the predictions of a bundle are packed at the slot of its first region.
__global__ void __launch_bounds__(128) neutral_preds_kernel(DevBatchView v, GroupsView g, NeutralView nv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.n_bundles) return;
  const uint32_t r0 = nv.reg_off[b], nr = nv.reg_cnt[b];
  const int64_t refstart = v.b_start[b];
  const float* __restrict__ c1 = v.bpcov + v.cov_pitch + v.b_gbase[b];
  const size_t out0 = (size_t)nv.cap_off[r0] + r0;
  uint32_t np = 0;
  int geneno = 0, t = 1;
  auto emit = [&](uint32_t a, uint32_t z, int len) {
    const size_t o = out0 + np;
    nv.p_start[o] = a; nv.p_end[o] = z; nv.p_tlen[o] = len; nv.p_geneno[o] = geneno - 1;
    nv.p_cov[o] = cov_all(c1, (uint32_t)((int64_t)a - refstart), (uint32_t)((int64_t)z - refstart)) / len;
    np++; t++;
  };
  for (uint32_t r = r0; r < r0 + nr; r++) {
    if (nv.reg_first[r]) t = 1;
    if (t == 1) geneno++;
    const uint32_t ns = nv.reg_start[r], ne = nv.reg_end[r], c0 = nv.cap_off[r], nt = nv.trim_cnt[r];
    uint32_t predstart = ns;
    for (uint32_t i = 0; i < nt; i++) {
      const uint32_t pos = nv.trim_pos[c0 + i];
      if (!pos) continue;
      if (nv.trim_st[c0 + i]) {
        const int len = (int)(pos - 100u /*CHI_WIN*/ - predstart + 1u);
        if (len > nv.mintranscriptlen) emit(predstart, pos - 100u, len);
        predstart = pos;
      } else {
        const int len = (int)(pos - predstart + 1u);
        if (len > nv.mintranscriptlen) emit(predstart, pos, len);
        predstart = pos + 100u;
      }
    }
    const int len = (int)(ne - predstart + 1u);
    if (len > nv.mintranscriptlen) emit(predstart, ne, len);
  }
  nv.pred_off[b] = np;
  nv.b_geneno[b] = geneno;
  if (b == 0) nv.pred_off[v.n_bundles] = 0;
}

// pred_off: exclusive prefix of the counts
__global__ void __launch_bounds__(128) neutral_compact_kernel(DevBatchView v, NeutralView nv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.n_bundles) return;
  const uint32_t r0 = nv.reg_off[b];
  const size_t in0 = (size_t)nv.cap_off[r0] + r0, out0 = nv.pred_off[b];
  const uint32_t np = nv.pred_off[b + 1] - nv.pred_off[b];
  for (uint32_t i = 0; i < np; i++) {
    nv.o_start[out0 + i] = nv.p_start[in0 + i]; nv.o_end[out0 + i] = nv.p_end[in0 + i]; nv.o_cov[out0 + i] = nv.p_cov[in0 + i];
    nv.o_tlen[out0 + i] = nv.p_tlen[in0 + i]; nv.o_geneno[out0 + i] = nv.p_geneno[in0 + i];
  }
}

}  // namespace

cudaError_t launch_neutral_count(const DevBatchView& v, const GroupsView& g, uint32_t* reg_off, cudaStream_t st) {
  neutral_count_kernel<<<(unsigned)((v.n_bundles + 1 + 127) / 128), 128, 0, st>>>(v, g, reg_off);
  return cudaGetLastError();
}
cudaError_t launch_neutral_regions(const DevBatchView& v, const GroupsView& g, const NeutralView& nv, cudaStream_t st) {
  if (v.n_bundles == 0) return cudaSuccess;
  neutral_regions_kernel<<<(unsigned)((v.n_bundles + 127) / 128), 128, 0, st>>>(v, g, nv);
  return cudaGetLastError();
}
cudaError_t launch_neutral_preds(const DevBatchView& v, const GroupsView& g, const NeutralView& nv, cudaStream_t st) {
  if (v.n_bundles == 0) return cudaSuccess;
  neutral_preds_kernel<<<(unsigned)((v.n_bundles + 127) / 128), 128, 0, st>>>(v, g, nv);
  return cudaGetLastError();
}
cudaError_t launch_neutral_compact(const DevBatchView& v, const NeutralView& nv, cudaStream_t st) {
  if (v.n_bundles == 0) return cudaSuccess;
  neutral_compact_kernel<<<(unsigned)((v.n_bundles + 127) / 128), 128, 0, st>>>(v, nv);
  return cudaGetLastError();
}
