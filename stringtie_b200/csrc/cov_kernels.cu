// cov_kernels.cu — sm_100a kernels of the coverage / junction-support stage.
//
// What the stage must reproduce (reference, CPU, one bundle at a time):
//   add_read_to_cov / cov_edge_add   rlink.cpp:133-219     difference-array deposits, mate-merged, float32
//   junction support loop            rlink.cpp:16890-17016 per-read anchors -> CJunction support counters (f64)
//   blocked clamped scan             rlink.cpp:17027-17063 c_i = max(0, c_{i-1} + d_i); b_i = b_{i-1} + c_i,
//                                                          b restarted every BSIZE = 10000 entries
//   get_cov / get_cov_sign           rlink.cpp:1830-1853
//
// Design: thousands of bundles are processed at once.
//   K0  junc_anchor    one thread per junction row: the anchor length it demands (10 or 25).
//   K1  expand_count   one thread per read: counts the coverage intervals the read deposits (mate merge walk)
//                      and adds its junction support with f64 atomics (exact, hence order-free: DESIGN.md §3.4).
//   K2  scan           exclusive prefix sum of the counts (3-phase block scan).
//   K3  expand_write   one thread per read: writes its intervals (global position of +w, of -w, w, lane)
//                      in reference order, so "index order" == the order the reference applies deposits in.
//   K4  window scans   prefix max of the -w positions, suffix min of the +w positions: both are monotone, so
//                      a sub-tile finds the contiguous slice of intervals that can touch it by two searches.
//   K4b tile_plan      one thread per sub-tile: the two searches + everything else the tile kernel needs, as one
//                      32-byte record, in column-major ticket order (all first sub-tiles of all bundles, then
//                      all second ones ...) so that resident warps work on ~3500 different bundles at once.
//   K5  cov_tile       one WARP per sub-tile of 512 positions x 3 strand lanes held in shared memory:
//                      ordered deposit of the slice (float32 adds land on every address in reference order —
//                      bit-exact for any weights), then the two sequential recurrences over the sub-tile with
//                      the (c, b) state handed from the previous sub-tile of the bundle through global memory,
//                      then one coalesced streaming write of the finished bpcov.  The difference array never
//                      exists in HBM.
// All float arithmetic is compiled without FMA contraction (-fmad=false): the reference is plain x86-64 SSE.
#include "stgpu_internal.h"

#define FULL 0xffffffffu

// --------------------------------------------------------------------------------------------------------------
// small helpers
// --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_bundle_of_read(const uint32_t* __restrict__ b_read_off, int64_t nb,
                                                        uint32_t n) {
  // last b with b_read_off[b] <= n
  uint32_t lo = 0, hi = (uint32_t)nb;  // answer in [lo, hi)
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(b_read_off + mid) <= n) lo = mid; else hi = mid;
  }
  return lo;
}

// same, but the warp first tries the bundle of its first lane (reads of a warp almost always share a bundle)
__device__ __forceinline__ uint32_t find_bundle_of_read_warp(const uint32_t* __restrict__ b_read_off, int64_t nb,
                                                             uint32_t n, bool active) {
  const uint32_t mask = __activemask();
  const int leader = __ffs(mask) - 1;
  uint32_t b = 0;
  if ((int)(threadIdx.x & 31) == leader) b = find_bundle_of_read(b_read_off, nb, n);
  b = __shfl_sync(mask, b, leader);
  if (active) {
    // lanes beyond the leader's bundle walk forward (bundles may be empty: skip while the next one starts <= n)
    while (b + 1 < (uint32_t)nb && __ldg(b_read_off + b + 1) <= n) b++;
  }
  return b;
}

// Walk the coverage intervals read n deposits, in the reference's order (rlink.cpp:144-219):
// for every mate link i with pair_idx[i] <= n the union of both mates' segments with weight pair_count[i],
// then the unpaired remainder on the read's own segments.
template <class Emit>
__device__ __forceinline__ void walk_read_intervals(const DevBatchView& v, uint32_t r0, uint32_t n, Emit&& emit) {
  if (v.r_flags[n] & STG_RF_UNITIG) return;  // super-read alignments add no coverage (rlink.cpp:16887)
  const uint32_t* __restrict__ ss = v.seg_start;
  const uint32_t* __restrict__ se = v.seg_end;
  const uint32_t rs0 = v.r_seg_off[n], rs1 = v.r_seg_off[n + 1];
  const int sno = (int)v.r_strand[n] + 1;
  float single_count = v.r_count[n];
  const uint32_t p0 = v.r_pair_off[n], p1 = v.r_pair_off[n + 1];
  for (uint32_t k = p0; k < p1; k++) {
    const float pc = v.pair_cnt[k];
    single_count = __fsub_rn(single_count, pc);
    const int32_t pl = v.pair_idx[k];
    if (pl < 0) continue;
    const uint32_t np = r0 + (uint32_t)pl;
    if (np > n) continue;
    const uint32_t ps0 = v.r_seg_off[np], ps1 = v.r_seg_off[np + 1];
    const int snop = (int)v.r_strand[np] + 1;
    int lane = sno;
    if (sno != snop) {
      if (sno == 1) lane = snop;
      else if (snop != 1) lane = 1;
    }
    uint32_t p = ps0, r = rs0;
    while (p < ps1 || r < rs1) {
      uint32_t start, end;
      if (r < rs1) {
        const uint32_t rs = ss[r], re = se[r];
        if (p == ps1) { start = rs; end = re; r++; }
        else {
          const uint32_t qs = ss[p], qe = se[p];
          if (qs > re) { start = rs; end = re; r++; }
          else if (qe < rs) { start = qs; end = qe; p++; }
          else {
            start = qs < rs ? qs : rs; end = qe; p++;
            bool grew = true;
            while (grew) {
              grew = false;
              if (r < rs1 && ss[r] <= end) { if (se[r] > end) end = se[r]; r++; grew = true; }
              if (p < ps1 && ss[p] <= end) { if (se[p] > end) end = se[p]; p++; grew = true; }
            }
          }
        }
      } else { start = ss[p]; end = se[p]; p++; }
      emit(start, end, pc, lane);
    }
  }
  if ((double)single_count > 0.000001) {
    for (uint32_t s = rs0; s < rs1; s++) emit(ss[s], se[s], single_count, sno);
  }
}

// --------------------------------------------------------------------------------------------------------------
// K0: per-junction anchor requirement (rlink.cpp:16991-16997): junctionsupport, raised to longintronanchor for
// introns longer than longintron or (short reads) junctions whose non-mismatch support is below junctionthr
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) junc_anchor_kernel(DevBatchView v, KernelParams prm) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= v.n_juncs) return;
  uint32_t anchor = prm.junctionsupport;
  const uint32_t jlen = v.j_end[j] - v.j_start[j] + 1;
  if (jlen > 100000u /*longintron, rlink.h:31*/ ||
      (!prm.longreads && v.j_nreads[j] - v.j_nm[j] < (double)prm.junctionthr && v.j_mm[j] == 0.0)) {
    if (anchor < 25u /*longintronanchor, rlink.h:32*/) anchor = 25u;
  }
  v.junc_anchor[j] = (uint8_t)(anchor > 255u ? 255u : anchor);
}

cudaError_t launch_junc_anchor(const DevBatchView& v, const KernelParams& p, cudaStream_t st) {
  if (v.n_juncs == 0) return cudaSuccess;
  junc_anchor_kernel<<<(unsigned)((v.n_juncs + 255) / 256), 256, 0, st>>>(v, p);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K1: count intervals per read + junction support
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_count_kernel(DevBatchView v, KernelParams prm) {
  const int64_t n64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n64 >= v.n_reads) return;
  const uint32_t n = (uint32_t)n64;
  const uint32_t b = find_bundle_of_read_warp(v.b_read_off, v.n_bundles, n, true);
  const uint32_t r0 = v.b_read_off[b];
  uint32_t cnt = 0;
  walk_read_intervals(v, r0, n, [&](uint32_t, uint32_t, float, int) { cnt++; });
  v.r_ivl_cnt[n] = cnt;
  if (n64 == v.n_reads - 1) v.r_ivl_cnt[v.n_reads] = 0;

  // junction support (rlink.cpp:16890-17016): left anchor = longest exon at or before the intron,
  // right anchor = longest exon after it.
  const uint32_t s0 = v.r_seg_off[n], s1 = v.r_seg_off[n + 1];
  const uint32_t nex = s1 - s0;
  if (nex < 2) return;
  float rdcount = v.r_count[n];
  if ((v.r_flags[n] & STG_RF_UNITIG) && rdcount > 1.f) rdcount = 1.f;
  const int32_t* __restrict__ rj = v.rjunc_idx + (s0 - n);
  const uint32_t j0 = v.b_junc_off[b];
  constexpr uint32_t kLocal = 24;
  uint32_t rsup[kLocal];  // rsup[i] = longest exon among segs[i..nex-1], for i < kLocal
  {
    uint32_t m = 0;
    for (uint32_t i = nex; i-- > 0;) {
      const uint32_t l = v.seg_end[s0 + i] - v.seg_start[s0 + i] + 1;
      if (l > m) m = l;
      if (i < kLocal) rsup[i] = m;
    }
  }
  const double w = (double)rdcount;
  uint32_t lmax = 0;
  for (uint32_t i = 1; i < nex; i++) {
    const uint32_t l = v.seg_end[s0 + i - 1] - v.seg_start[s0 + i - 1] + 1;
    if (l > lmax) lmax = l;
    const int32_t jl = rj[i - 1];
    if (jl < 0) continue;
    uint32_t rmax;
    if (i < kLocal) rmax = rsup[i];
    else {  // very long reads: recompute the suffix maximum
      rmax = 0;
      for (uint32_t q = i; q < nex; q++) {
        const uint32_t lq = v.seg_end[s0 + q] - v.seg_start[s0 + q] + 1;
        if (lq > rmax) rmax = lq;
      }
    }
    const uint32_t j = j0 + (uint32_t)jl;
    const uint32_t anchor = v.junc_anchor[j];
    if (lmax >= anchor) {
      atomicAdd(v.j_leftsupport + j, w);
      if (rmax >= anchor) { atomicAdd(v.j_rightsupport + j, w); atomicAdd(v.j_nreads_good + j, w); }
    } else if (rmax >= anchor) {
      atomicAdd(v.j_rightsupport + j, w);
    }
  }
}

// --------------------------------------------------------------------------------------------------------------
// K3: write intervals
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_write_kernel(DevBatchView v) {
  const int64_t n64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n64 >= v.n_reads) return;
  const uint32_t n = (uint32_t)n64;
  uint32_t k = v.r_ivl_cnt[n];
  const bool any = k != v.r_ivl_cnt[n + 1];
  const uint32_t b = find_bundle_of_read_warp(v.b_read_off, v.n_bundles, n, any);
  if (!any) return;
  const uint32_t r0 = v.b_read_off[b];
  const uint32_t gb = v.b_gbase[b];
  const uint32_t refstart = (uint32_t)v.b_start[b];
  const uint32_t len = v.b_len[b];
  walk_read_intervals(v, r0, n, [&](uint32_t start, uint32_t end, float w, int lane) {
    // cov_edge_add(lane, start-refstart, end-refstart+1, w): +w at index start-refstart+1, -w at end-refstart+2
    uint32_t ia = start - refstart + 1, ib = end - refstart + 2;
    if (start < refstart || ib >= len) { ia = 0; ib = 0; w = 0.f; }  // read outside its bundle: dropped (the
                                                                    // reference would write out of bounds)
    v.ivl[k] = make_uint4(gb + ia, gb + ib, __float_as_uint(w), (uint32_t)lane);
    v.ivl_smin_a[k] = gb + ia;
    v.ivl_pmax_b[k] = gb + ib;
    k++;
  });
}

// --------------------------------------------------------------------------------------------------------------
// K2/K4: 3-phase block scans over uint32 (add / max / min; forward or reversed index order)
// --------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

enum { OP_ADD = 0, OP_MAX = 1, OP_MIN = 2 };
template <int OP> __device__ __forceinline__ uint32_t scan_op(uint32_t a, uint32_t b) {
  if (OP == OP_ADD) return a + b;
  if (OP == OP_MAX) return a > b ? a : b;
  return a < b ? a : b;
}
template <int OP> __device__ __forceinline__ uint32_t scan_identity() { return OP == OP_MIN ? 0xffffffffu : 0u; }

int64_t scan_partials_needed(int64_t n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 1; }

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total gets the block total
template <int OP> __device__ __forceinline__ uint32_t block_exclusive(uint32_t x, uint32_t* total) {
  __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
  __shared__ uint32_t block_tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = x;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl = scan_op<OP>(y, incl);
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : scan_identity<OP>();
    uint32_t ti = t;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(FULL, ti, o);
      if (lane >= o) ti = scan_op<OP>(y, ti);
    }
    uint32_t excl = __shfl_up_sync(FULL, ti, 1);
    if (lane == 0) excl = scan_identity<OP>();
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = excl;
    if (lane == 31) block_tot = ti;
  }
  __syncthreads();
  uint32_t excl_lane = __shfl_up_sync(FULL, incl, 1);
  if (lane == 0) excl_lane = scan_identity<OP>();
  uint32_t res = scan_op<OP>(warp_tot[wid], excl_lane);
  *total = block_tot;
  __syncthreads();
  return res;
}

template <int OP, bool REVERSE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                    uint32_t* __restrict__ partials) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
  uint32_t acc = scan_identity<OP>();
  for (int it = 0; it < SCAN_ITEMS; it++) {
    const int64_t i = base + (int64_t)it * SCAN_THREADS + threadIdx.x;
    if (i < n) acc = scan_op<OP>(acc, in[REVERSE ? n - 1 - i : i]);
  }
  uint32_t tot;
  block_exclusive<OP>(acc, &tot);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// single block: exclusive scan of the per-chunk totals, in place
template <int OP>
__global__ void __launch_bounds__(SCAN_THREADS) scan_partials_kernel(uint32_t* __restrict__ partials, int64_t nb) {
  uint32_t carry = scan_identity<OP>();
  for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
    const int64_t i = base + threadIdx.x;
    const uint32_t x = i < nb ? partials[i] : scan_identity<OP>();
    uint32_t tot;
    const uint32_t ex = block_exclusive<OP>(x, &tot);
    if (i < nb) partials[i] = scan_op<OP>(carry, ex);
    carry = scan_op<OP>(carry, tot);
  }
}

// out[i] = scan over logical positions; EXCLUSIVE only for OP_ADD (in place allowed)
template <int OP, bool REVERSE, bool EXCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* in, uint32_t* out, int64_t n,
                                                                   const uint32_t* __restrict__ partials) {
  // each thread owns SCAN_ITEMS consecutive logical elements
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t x[SCAN_ITEMS];
  uint32_t acc = scan_identity<OP>();
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; it++) {
    const int64_t i = base + it;
    x[it] = i < n ? in[REVERSE ? n - 1 - i : i] : scan_identity<OP>();
    acc = scan_op<OP>(acc, x[it]);
  }
  uint32_t tot;
  uint32_t run = scan_op<OP>(partials[blockIdx.x], block_exclusive<OP>(acc, &tot));
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; it++) {
    const int64_t i = base + it;
    uint32_t nx = scan_op<OP>(run, x[it]);
    if (i < n) out[REVERSE ? n - 1 - i : i] = EXCLUSIVE ? run : nx;
    run = nx;
  }
}

template <int OP, bool REVERSE, bool EXCLUSIVE>
static cudaError_t run_scan(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* partials, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  scan_reduce_kernel<OP, REVERSE><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, partials);
  scan_partials_kernel<OP><<<1, SCAN_THREADS, 0, st>>>(partials, nb);
  scan_apply_kernel<OP, REVERSE, EXCLUSIVE><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, partials);
  return cudaGetLastError();
}

cudaError_t launch_scan_exclusive_add(uint32_t* data, int64_t n_plus_1, uint32_t* partials, cudaStream_t st) {
  return run_scan<OP_ADD, false, true>(data, data, n_plus_1, partials, st);
}

cudaError_t launch_window_scans(const DevBatchView& v, cudaStream_t st) {
  cudaError_t e = run_scan<OP_MAX, false, false>(v.ivl_pmax_b, v.ivl_pmax_b, v.n_ivl, v.scan_partials, st);
  if (e != cudaSuccess) return e;
  return run_scan<OP_MIN, true, false>(v.ivl_smin_a, v.ivl_smin_a, v.n_ivl, v.scan_partials, st);
}

cudaError_t launch_expand_count(const DevBatchView& v, const KernelParams& p, cudaStream_t st) {
  if (v.n_reads == 0) return cudaSuccess;
  expand_count_kernel<<<(unsigned)((v.n_reads + 255) / 256), 256, 0, st>>>(v, p);
  return cudaGetLastError();
}

cudaError_t launch_expand_write(const DevBatchView& v, cudaStream_t st) {
  if (v.n_reads == 0) return cudaSuccess;
  expand_write_kernel<<<(unsigned)((v.n_reads + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K4b: per-tile plan records, in ticket (column-major) order
// --------------------------------------------------------------------------------------------------------------
// first index in [lo, hi) with arr[idx] >= key (arr non-decreasing)
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ arr, uint32_t lo, uint32_t hi,
                                                    uint32_t key) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(arr + mid) >= key) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) tile_plan_kernel(DevBatchView v) {
  const int64_t t64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t64 >= v.n_tiles) return;
  const uint32_t t = (uint32_t)t64;
  // column j = last j with col_off[j] <= t; rank within the column -> bundle
  uint32_t lo = 0, hi = (uint32_t)v.n_cols;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(v.col_off + mid) <= t) lo = mid; else hi = mid;
  }
  const uint32_t jt = lo;
  const uint32_t b = v.b_order[t - v.col_off[jt]];
  const uint32_t len = v.b_len[b];
  const uint32_t i0 = jt * STG_TILE_W;
  TileRec r;
  r.gbase = v.b_gbase[b];
  r.g0 = r.gbase + i0;
  r.npos = min((uint32_t)STG_TILE_W, len - i0);
  r.stride = v.b_stride[b];
  r.tile_id = v.b_tile_off[b] + jt;
  r.pad = 0;
  const uint32_t K0 = v.r_ivl_cnt[v.b_read_off[b]], K1 = v.r_ivl_cnt[v.b_read_off[b + 1]];
  r.k0 = lower_bound_u32(v.ivl_pmax_b, K0, K1, r.g0);             // first interval whose -w lands at or after g0
  r.k1 = lower_bound_u32(v.ivl_smin_a, K0, K1, r.g0 + r.npos);    // first interval from which every +w is past the tile
  if (r.k1 < r.k0) r.k1 = r.k0;
  v.tile_plan[t] = r;
}

cudaError_t launch_tile_plan(const DevBatchView& v, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  tile_plan_kernel<<<(unsigned)((v.n_tiles + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K5: the tile kernel
// --------------------------------------------------------------------------------------------------------------
constexpr int TW = STG_TILE_W;          // positions per sub-tile
constexpr int TWS = TW + 4;             // lane stride in shared memory (+4 floats: the three chain threads hit
                                        // different banks with their 128-bit accesses)
constexpr int TNW = TW / 32;            // 32-position words per sub-tile
constexpr int TILE_SMEM_WORDS = 3 * TWS + TW + 3 * TNW + 3 * TNW;  // d, stamp, masks, fills

int cov_tile_smem_bytes() { return STG_TILE_WARPS * TILE_SMEM_WORDS * 4; }

__device__ __forceinline__ void deposit_one(float* d, uint32_t sl, uint32_t p, float w, bool add) {
  float* q = d + sl * TWS + p;
  *q = add ? __fadd_rn(*q, w) : __fsub_rn(*q, w);
  if (sl != 1) {
    float* q1 = d + TWS + p;
    *q1 = add ? __fadd_rn(*q1, w) : __fsub_rn(*q1, w);
  }
}

__global__ void __launch_bounds__(STG_TILE_WARPS * 32, 1) cov_tile_kernel(DevBatchView v) {
  extern __shared__ __align__(16) uint32_t smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* d = reinterpret_cast<float*>(smem_raw) + (size_t)wid * TILE_SMEM_WORDS;  // d[3][TWS]
  uint32_t* stamp = reinterpret_cast<uint32_t*>(d + 3 * TWS);                     // stamp[TW]
  uint32_t* masks = stamp + TW;                                                   // masks[3][TNW]
  float* fills = reinterpret_cast<float*>(masks + 3 * TNW);                       // fills[3][TNW]

  // stamps persist across the sub-tiles of this warp: keys only ever decrease, so stale ones never match
  for (int q = lane; q < TW; q += 32) stamp[q] = FULL;
  uint32_t epoch = 0x03ffffffu;
  __syncwarp();

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(v.tile_ticket, 1u);
    t = __shfl_sync(FULL, t, 0);
    if ((int64_t)t >= v.n_tiles) break;
    // one 32-byte record per sub-tile
    uint32_t recw = 0;
    if (lane < 8) recw = __ldg(reinterpret_cast<const uint32_t*>(v.tile_plan + t) + lane);
    const uint32_t g0 = __shfl_sync(FULL, recw, 0), gbase = __shfl_sync(FULL, recw, 1);
    const uint32_t k0 = __shfl_sync(FULL, recw, 2), k1 = __shfl_sync(FULL, recw, 3);
    const uint32_t stride = __shfl_sync(FULL, recw, 4), npos = __shfl_sync(FULL, recw, 5);
    const uint32_t tile_id = __shfl_sync(FULL, recw, 6);
    const uint32_t i0 = g0 - gbase;

    // state handed over by the previous sub-tile of this bundle: issue the loads now, consume them after the deposit
    float c = 0.f, bs = 0.f;
    bool have_carry = (i0 == 0);
    if (!have_carry) {
      uint32_t f = 0;
      if (lane == 0) f = atomicAdd(v.tile_flag + (tile_id - 1), 0u);
      f = __shfl_sync(FULL, f, 0);
      if (f) {
        __threadfence();
        if (lane < 3) { c = __ldcg(v.tile_carry + (size_t)(tile_id - 1) * 8 + lane); bs = __ldcg(v.tile_carry + (size_t)(tile_id - 1) * 8 + 4 + lane); }
        have_carry = true;
      }
    }

    if (epoch < 0x10000u) {  // (practically never) restart the key space
      for (int q = lane; q < TW; q += 32) stamp[q] = FULL;
      epoch = 0x03ffffffu;
    }
    // clear the sub-tile
    {
      float4* d4 = reinterpret_cast<float4*>(d);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = lane; q < 3 * TWS / 4; q += 32) d4[q] = z;
    }
    __syncwarp();

    // ordered deposit of intervals [k0, k1). `epoch` decreases, so keys of later rounds are smaller than stale stamps.
    uint4 cur = make_uint4(FULL, FULL, 0u, 1u);
    if (k0 + lane < k1) cur = __ldcs(v.ivl + k0 + lane);
    for (uint32_t kb = k0; kb < k1; kb += 32) {
      uint4 nxt = make_uint4(FULL, FULL, 0u, 1u);
      if (kb + 32 + lane < k1) nxt = __ldcs(v.ivl + kb + 32 + lane);
      const uint32_t pa = cur.x - g0, pz = cur.y - g0;   // unsigned: out-of-tile positions become >= npos
      const bool ina = pa < npos, inz = pz < npos;
      const float w = __uint_as_float(cur.z);
      const uint32_t sl = cur.w;
      // round 0: plain stores detect whether any two lanes share a position
      uint32_t key = (epoch << 5) | (uint32_t)lane; epoch--;
      if (ina) stamp[pa] = key;
      if (inz) stamp[pz] = key;
      __syncwarp();
      const bool mine = (!ina || stamp[pa] == key) && (!inz || stamp[pz] == key);
      if (__all_sync(FULL, mine)) {
        if (ina) deposit_one(d, sl, pa, w, true);
        if (inz) deposit_one(d, sl, pz, w, false);
        __syncwarp();
      } else {
        // some position is hit by several intervals of this group: apply in lane (= reference) order
        bool pending = ina || inz;
        while (__any_sync(FULL, pending)) {
          key = (epoch << 5) | (uint32_t)lane; epoch--;
          if (pending) { if (ina) atomicMin(stamp + pa, key); if (inz) atomicMin(stamp + pz, key); }
          __syncwarp();
          const bool win = pending && (!ina || stamp[pa] == key) && (!inz || stamp[pz] == key);
          if (win) {
            if (ina) deposit_one(d, sl, pa, w, true);
            if (inz) deposit_one(d, sl, pz, w, false);
            pending = false;
          }
          __syncwarp();
        }
      }
      cur = nxt;
    }
    __syncwarp();

    // which 32-position words received any deposit, per strand lane
    const int nwords = (int)((npos + 31) / 32);
    for (int q = 0; q < nwords; q++) {
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const uint32_t m = __ballot_sync(FULL, d[s * TWS + q * 32 + lane] != 0.f);
        if (lane == 0) masks[s * TNW + q] = m;
      }
    }
    __syncwarp();

    if (!have_carry) {
      if (lane == 0) { while (atomicAdd(v.tile_flag + (tile_id - 1), 0u) == 0u) __nanosleep(40); }
      __syncwarp();
      __threadfence();
      if (lane < 3) { c = __ldcg(v.tile_carry + (size_t)(tile_id - 1) * 8 + lane); bs = __ldcg(v.tile_carry + (size_t)(tile_id - 1) * 8 + 4 + lane); }
    }

    // the two recurrences (rlink.cpp:17037-17063), one strand lane per thread:
    //   c = max(0, c + d[i]);  bs = (i % BSIZE == 0 ? 0 : bs) + c;  out[i] = bs
    uint32_t fillmask = 0;
    if (lane < 3) {
      float* dl = d + lane * TWS;
      const uint32_t* ml = masks + lane * TNW;
      float* fl = fills + lane * TNW;
      const uint32_t to_reset = (STG_BSIZE - i0 % STG_BSIZE) % STG_BSIZE;  // in-tile position where idx % BSIZE == 0
      for (int q = 0; q < nwords; q++) {
        const uint32_t base = (uint32_t)q * 32;
        const bool full = base + 32 <= npos;
        const bool has_reset = to_reset >= base && to_reset < base + 32;
        if (full && !has_reset) {
          if (ml[q] == 0 && __fadd_rn(bs, c) == bs) {  // nothing deposited and c too small to move bs: constant word
            fl[q] = bs;
            fillmask |= 1u << q;
            continue;
          }
          float4 x[8];
#pragma unroll
          for (int u = 0; u < 8; u++) x[u] = *reinterpret_cast<const float4*>(dl + base + u * 4);
#pragma unroll
          for (int u = 0; u < 8; u++) {
            c = fmaxf(__fadd_rn(x[u].x, c), 0.f); bs = __fadd_rn(c, bs); x[u].x = bs;
            c = fmaxf(__fadd_rn(x[u].y, c), 0.f); bs = __fadd_rn(c, bs); x[u].y = bs;
            c = fmaxf(__fadd_rn(x[u].z, c), 0.f); bs = __fadd_rn(c, bs); x[u].z = bs;
            c = fmaxf(__fadd_rn(x[u].w, c), 0.f); bs = __fadd_rn(c, bs); x[u].w = bs;
          }
#pragma unroll
          for (int u = 0; u < 8; u++) *reinterpret_cast<float4*>(dl + base + u * 4) = x[u];
        } else {
          const uint32_t lim = min(base + 32, npos);
          for (uint32_t i = base; i < lim; i++) {
            if (i == to_reset) bs = 0.f;
            c = fmaxf(__fadd_rn(dl[i], c), 0.f);
            bs = __fadd_rn(c, bs);
            dl[i] = bs;
          }
        }
      }
      __stcg(v.tile_carry + (size_t)tile_id * 8 + lane, c);
      __stcg(v.tile_carry + (size_t)tile_id * 8 + 4 + lane, bs);
      __threadfence();
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); atomicExch(v.tile_flag + tile_id, 1u); }

    // constant words
#pragma unroll
    for (int s = 0; s < 3; s++) {
      uint32_t fm = __shfl_sync(FULL, fillmask, s);
      while (fm) {
        const int q = __ffs(fm) - 1;
        fm &= fm - 1;
        d[s * TWS + q * 32 + lane] = fills[s * TNW + q];
      }
    }
    __syncwarp();

    // finished bpcov of the sub-tile: one streaming, coalesced write per strand lane (incl. the zero padding)
    {
      float* out = v.bpcov + 3 * (size_t)gbase + i0;
      const uint32_t n4 = min((uint32_t)TW, stride - i0) / 4;
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const float4* src = reinterpret_cast<const float4*>(d + s * TWS);
        float4* dst = reinterpret_cast<float4*>(out + (size_t)s * stride);
        for (uint32_t q = lane; q < n4; q += 32) __stcs(dst + q, src[q]);
      }
    }
    __syncwarp();
  }
}

cudaError_t launch_cov_tiles(const DevBatchView& v, int num_sms, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  static bool attr_set = false;
  const int smem = cov_tile_smem_bytes();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cov_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int64_t want = (v.n_tiles + STG_TILE_WARPS - 1) / STG_TILE_WARPS;
  int grid = (int)(want < num_sms ? want : num_sms);
  cov_tile_kernel<<<grid, STG_TILE_WARPS * 32, smem, st>>>(v);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// a5: get_cov / get_cov_sign (rlink.cpp:1830-1853) on the device-resident bpcov
// --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dev_get_cov(const float* __restrict__ c, uint32_t start, uint32_t end) {
  const int m = (int)((end + 1) / STG_BSIZE);
  const int k = (int)(start / STG_BSIZE);
  double cov = 0;
  for (int i = k + 1; i <= m; i++) cov += (double)c[(size_t)i * STG_BSIZE - 1];
  cov += (double)__fsub_rn(c[end + 1], c[start]);
  if (cov < 0) cov = 0;
  return cov;
}

__device__ __forceinline__ double dev_get_cov_sign(const float* __restrict__ all, const float* __restrict__ opp,
                                                   uint32_t start, uint32_t end) {
  const int m = (int)((end + 1) / STG_BSIZE);
  const int k = (int)(start / STG_BSIZE);
  double cov = 0;
  for (int i = k + 1; i <= m; i++) cov += (double)__fsub_rn(all[(size_t)i * STG_BSIZE - 1], opp[(size_t)i * STG_BSIZE - 1]);
  float e = __fsub_rn(all[end + 1], all[start]);
  e = __fsub_rn(e, opp[end + 1]);
  e = __fadd_rn(e, opp[start]);
  cov += (double)e;
  if (cov < 0) cov = 0;
  return cov;
}

__global__ void query_cov_kernel(DevBatchView v, int64_t n, const int32_t* __restrict__ bundle,
                                 const int8_t* __restrict__ lane, const uint32_t* __restrict__ start,
                                 const uint32_t* __restrict__ end, int sign, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t b = bundle[i];
  const float* base = v.bpcov + 3 * (size_t)v.b_gbase[b];
  const uint32_t stride = v.b_stride[b];
  const int s = lane[i];
  if (!sign) out[i] = dev_get_cov(base + (size_t)s * stride, start[i], end[i]);
  else out[i] = dev_get_cov_sign(base + stride, base + (size_t)(2 - s) * stride, start[i], end[i]);
}

cudaError_t launch_query_cov(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* lane,
                             const uint32_t* start, const uint32_t* end, int sign, double* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  query_cov_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, bundle, lane, start, end, sign, out);
  return cudaGetLastError();
}

// per (bundle, lane): get_cov(s, 0, len-2) — the summed clamped coverage of the whole bundle
__global__ void cov_totals_kernel(DevBatchView v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_bundles * 3) return;
  const int64_t b = i / 3;
  const int s = (int)(i % 3);
  const float* base = v.bpcov + 3 * (size_t)v.b_gbase[b] + (size_t)s * v.b_stride[b];
  v.b_cov_total[i] = dev_get_cov(base, 0u, v.b_len[b] - 2u);
}

cudaError_t launch_cov_totals(const DevBatchView& v, cudaStream_t st) {
  if (v.n_bundles == 0) return cudaSuccess;
  cov_totals_kernel<<<(unsigned)((v.n_bundles * 3 + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}
