// stgpu_internal.h — shared between the C-ABI host code (stgpu_api.cu) and the kernels (cov_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/stgpu.h"

#define STG_TILE_W 1024          // positions per sub-tile (one warp owns one sub-tile of all three strand lanes)
#define STG_TILE_WARPS 12        // warps per CTA of the tile kernel: 12 * 16 KB = 192 KB dynamic shared memory
#define STG_LANE_ALIGN 32        // every strand lane of every bundle starts on a 128-byte boundary

// Device view of one batch: inputs (copies of stg_packed_batch arrays), the planned layout, scratch and outputs.
struct DevBatchView {
  int64_t n_bundles, n_reads, n_segs, n_pairs, n_juncs;
  // inputs
  const int32_t *b_start, *b_end;
  const uint32_t *b_read_off, *b_junc_off;
  const uint32_t *r_start, *r_end, *r_len;
  const float* r_count;
  const int16_t* r_nh;
  const int8_t* r_strand;
  const uint8_t* r_flags;
  const uint32_t *r_seg_off, *r_pair_off, *seg_start, *seg_end;
  const int32_t* rjunc_idx;
  const int32_t* pair_idx;
  const float* pair_cnt;
  const uint32_t *j_start, *j_end;
  const int8_t *j_strand, *j_guide_match;
  const double *j_nreads, *j_nm, *j_mm;
  // layout (planned on the host in stg_upload)
  const uint32_t* b_len;       // [n_bundles]   end-start+3
  const uint32_t* b_stride;    // [n_bundles]   len rounded up to STG_LANE_ALIGN
  const uint32_t* b_gbase;     // [n_bundles+1] prefix sum of b_stride: "global position" of index 0 of bundle b
  const uint32_t* b_tile_off;  // [n_bundles+1] prefix sum of ceil(len/STG_TILE_W)
  int64_t n_tiles;
  int64_t total_pos;           // b_gbase[n_bundles]
  // scratch
  uint32_t* r_ivl_cnt;         // [n_reads+1] per-read interval count, then exclusive prefix (in place)
  int64_t n_ivl;               // filled after the count pass (host reads it back)
  uint32_t* ivl_a;             // [n_ivl] global position of the +w deposit
  uint32_t* ivl_b;             // [n_ivl] global position of the -w deposit
  float* ivl_w;                // [n_ivl]
  uint8_t* ivl_lane;           // [n_ivl] 0,1,2
  uint32_t* ivl_pmax_b;        // [n_ivl] inclusive prefix max of ivl_b
  uint32_t* ivl_smin_a;        // [n_ivl] inclusive suffix min of ivl_a
  uint32_t* scan_partials;     // scratch for the 3-phase scans
  float* tile_carry;           // [n_tiles*8] c[3], bsum[3] handed from sub-tile t to t+1 of the same bundle
  uint32_t* tile_flag;         // [n_tiles]   1 when tile_carry[t] is published
  uint32_t* tile_ticket;       // [1]
  // outputs
  float* bpcov;                // [3*total_pos]: bundle b lane s index i at 3*b_gbase[b] + s*b_stride[b] + i
  double *j_nreads_good, *j_leftsupport, *j_rightsupport;  // [n_juncs]
  double* b_cov_total;         // [n_bundles*3]
};

struct KernelParams {
  uint32_t junctionsupport;
  int32_t junctionthr;
  uint32_t longreads;
};

// kernel launchers (cov_kernels.cu). Each enqueues on `st` and returns the CUDA error of the launch.
cudaError_t launch_expand_count(const DevBatchView& v, const KernelParams& p, cudaStream_t st);
cudaError_t launch_scan_exclusive_add(uint32_t* data, int64_t n_plus_1, uint32_t* partials, cudaStream_t st);
cudaError_t launch_expand_write(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_window_scans(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_cov_tiles(const DevBatchView& v, int num_sms, cudaStream_t st);
cudaError_t launch_cov_totals(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_query_cov(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* lane,
                             const uint32_t* start, const uint32_t* end, int sign, double* out, cudaStream_t st);
int64_t scan_partials_needed(int64_t n);
int cov_tile_smem_bytes();
