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
#include "trims_device.cuh"

__global__ void __launch_bounds__(256) find_trims_kernel(DevBatchView v, int64_t n, const int32_t* __restrict__ bundle,
                                                         const int8_t* __restrict__ sno_in, const uint32_t* __restrict__ start_in,
                                                         const uint32_t* __restrict__ end_in, const uint32_t* __restrict__ cap_off,
                                                         uint32_t* __restrict__ o_cnt, uint32_t* o_pos, float* o_ab, uint8_t* o_st) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const int32_t b = bundle[w];
  const int sno = sno_in[w];
  const float* __restrict__ base = v.bpcov + v.b_gbase[b];
  const size_t stride = v.cov_pitch;
  const uint32_t cnt = find_trims_warp(base + stride, base + (size_t)(2 - sno) * stride, (uint32_t)v.b_start[b], start_in[w], end_in[w],
                                       o_pos + cap_off[w], o_ab + cap_off[w], o_st + cap_off[w], cap_off[w + 1] - cap_off[w]);
  if ((threadIdx.x & 31) == 0) o_cnt[w] = cnt;
}

cudaError_t launch_find_trims(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* sno, const uint32_t* start,
                              const uint32_t* end, const uint32_t* cap_off, uint32_t* o_cnt, uint32_t* o_pos, float* o_ab,
                              uint8_t* o_st, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  find_trims_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(v, n, bundle, sno, start, end, cap_off, o_cnt, o_pos, o_ab, o_st);
  return cudaGetLastError();
}
