// cov_kernels.cu — sm_100a kernels of the coverage / junction-support stage.
//
// What the stage must reproduce (reference, CPU, one bundle at a time):
//   add_read_to_cov / cov_edge_add   rlink.cpp:133-219     difference-array deposits, mate-merged, float32
//   junction support loop            rlink.cpp:16890-17016 per-read anchors -> CJunction support counters (f64)
//   blocked clamped scan             rlink.cpp:17027-17063 c_i = max(0, c_{i-1} + d_i); b_i = b_{i-1} + c_i,
//                                                          b restarted every BSIZE = 10000 entries
//   get_cov / get_cov_sign           rlink.cpp:1830-1853
//
// Design: thousands of bundles are processed at once; float32 deposits must land on every address in
// the reference's order (read order), so the events are BINNED, not scattered with atomics:
//   K0  junc_anchor    one thread per junction row: the anchor length it demands (10 or 25).
//   K1  expand_count   one thread per read: counts the coverage intervals the read deposits (mate merge walk),
//                      validates offsets / mate / junction indices, and adds its junction support: the warp groups
//                      its updates by (junction row, weight) and issues one f64 atomic of count * w per group
//                      (f64 sums of float weights are exact, hence order-free: DESIGN.md §5).
//   K2  scan           exclusive prefix sum of the counts (3-phase block scan).
//   K3  expand_write   one thread per read: writes its intervals {tile(+w), tile(-w), positions, lane, w}
//                      in reference order, so "index order" == the order the reference applies deposits in.
//   K4  chunk_hull     one warp per chunk of 512 intervals: the range of tiles its events land in; scans give
//                      each chunk a row of per-tile counters and make the ranges searchable.
//   K5  bin_count      one warp per chunk: events per (chunk, tile).
//   K6  bin_columns    one thread per tile: exclusive prefix down the chunks that reach it -> every chunk's first
//                      slot in the tile's event list; per-tile totals -> scan -> list offsets.
//   K7  bin_scatter    one warp per chunk, walking its intervals in order: a STABLE counting sort of the events by
//                      tile (rank = slot of chunk + rank inside the warp by ballot) -> tight per-tile event lists.
//   K8  tile_plan      one thread per sub-tile: one 32-byte record; sub-tiles with events get tickets in a
//                      proportional schedule (sub-tile j of nt runs in round j*R/nt), deep ones are listed for
//                      K8b, event-free ones are listed for K10.
//   K8b deep_deposit   one CTA per deep sub-tile (>= 2048 events): stable counting sort of its events by position,
//                      short per-position lists summed by one thread each; K8c deep_long: one warp per long list
//                      (a splice site of a deep locus) runs the bare chain of dependent adds.
//   K9  cov_tile       one WARP per sub-tile with events, 512 positions x 3 strand lanes in shared memory:
//                      ordered deposit of its event list (same-position events of a 32-group are chained in lane
//                      order inside one thread: bit-exact for any weights), then the two sequential recurrences
//                      with the (c, b) state taken from the nearest earlier sub-tile that had events (bridged over
//                      the event-free gap in closed form, advance_bs), then one coalesced streaming write of the
//                      finished bpcov.  The difference array never exists in HBM.
//   K10 cov_fill       one warp per event-free sub-tile: state from the carry + closed form, streaming write.
// All float arithmetic is compiled without FMA contraction (-fmad=false): the reference is plain x86-64 SSE.
#include "stgpu_internal.h"
#include "exact_chain.cuh"

#define FULL 0xffffffffu

// --------------------------------------------------------------------------------------------------------------
// small helpers
// --------------------------------------------------------------------------------------------------------------
// bundle of read n: start from the bundle of the first read of this CTA (planned on the host) and walk forward;
// empty bundles are skipped (the answer is the LAST b with b_read_off[b] <= n)
__device__ __forceinline__ uint32_t bundle_of_read(const DevBatchView& v, uint32_t n) {
  uint32_t b = __ldg(v.cta_bundle + blockIdx.x);
  const uint32_t nb = (uint32_t)v.n_bundles;
  while (b + 1 < nb && __ldg(v.b_read_off + b + 1) <= n) b++;
  return b;
}

// Walk the coverage intervals read n deposits, in the reference's order (rlink.cpp:144-219):
// for every mate link i with pair_idx[i] <= n the union of both mates' segments with weight pair_count[i],
// then the unpaired remainder on the read's own segments.
// Malformed input (offsets that run backwards or past their arrays, a mate index outside the bundle) must not turn
// into wild reads on the device: such reads / links are skipped — identically in the count and in the write pass —
// and reported through v.err_flag; stg_run then fails with STG_EINVAL.  (The reference has no such checks: it would
// index out of bounds.)
#define STG_ERR_SEG_OFF 1u
#define STG_ERR_PAIR_OFF 2u
#define STG_ERR_PAIR_IDX 4u
#define STG_ERR_JUNC_IDX 8u

// CHECK = true in the count pass only: stg_run aborts on a raised flag before the write pass is launched.
template <bool CHECK, class Emit>
__device__ __forceinline__ void walk_read_intervals(const DevBatchView& v, uint32_t r0, uint32_t r1, uint32_t n, Emit&& emit) {
  if (v.r_flags[n] & STG_RF_UNITIG) return;  // super-read alignments add no coverage (rlink.cpp:16887)
  const uint32_t* __restrict__ ss = v.seg_start;
  const uint32_t* __restrict__ se = v.seg_end;
  const uint32_t rs0 = v.r_seg_off[n], rs1 = v.r_seg_off[n + 1];
  if (CHECK && (rs1 < rs0 || (int64_t)rs1 > v.n_segs)) { atomicOr(v.err_flag, STG_ERR_SEG_OFF); return; }
  const int sno = (int)v.r_strand[n] + 1;
  float single_count = v.r_count[n];
  const uint32_t p0 = v.r_pair_off[n], p1 = v.r_pair_off[n + 1];
  if (CHECK && (p1 < p0 || (int64_t)p1 > v.n_pairs)) { atomicOr(v.err_flag, STG_ERR_PAIR_OFF); return; }
  for (uint32_t k = p0; k < p1; k++) {
    const float pc = v.pair_cnt[k];
    single_count = __fsub_rn(single_count, pc);
    const int32_t pl = v.pair_idx[k];
    if (pl < 0) continue;
    const uint32_t np = r0 + (uint32_t)pl;
    if (CHECK && np >= r1) { atomicOr(v.err_flag, STG_ERR_PAIR_IDX); continue; }
    if (np > n) continue;
    const uint32_t ps0 = v.r_seg_off[np], ps1 = v.r_seg_off[np + 1];
    if (CHECK && (ps1 < ps0 || (int64_t)ps1 > v.n_segs)) { atomicOr(v.err_flag, STG_ERR_SEG_OFF); continue; }
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
// K1: interval count of every read (the mate-merge walk of add_read_to_cov, rlink.cpp:144-219) and
// junction support (rlink.cpp:16890-17016): for intron i of a read, left anchor = longest exon at or before
// it, right anchor = longest exon after it; the junction row gets leftsupport / rightsupport / nreads_good += w.
// The sums are float weights accumulated in f64: every partial sum is exactly representable, so the order of the
// additions is free.  One thread per read; the 32 reads of a warp usually cross the same few junctions with the
// same few weights, so each round the warp groups its updates by (junction row, weight) with one 64-bit match and
// issues ONE f64 atomic of count * w per group and counter (the product is exact: count <= 32, w has 24 bits).
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(STG_READ_BLOCK, 8) expand_count_kernel(DevBatchView v) {
  const int64_t n64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = n64 < v.n_reads;
  const uint32_t n = live ? (uint32_t)n64 : 0u;
  uint32_t s0 = 0, nex = 0, b = 0;
  if (live) {
    b = bundle_of_read(v, n);
    uint32_t cnt = 0;
    walk_read_intervals<true>(v, v.b_read_off[b], v.b_read_off[b + 1], n, [&](uint32_t, uint32_t, float, int) { cnt++; });
    v.r_ivl_cnt[n] = cnt;
    if (n64 == v.n_reads - 1) v.r_ivl_cnt[v.n_reads] = 0;
    s0 = v.r_seg_off[n]; nex = v.r_seg_off[n + 1] - s0;
    if (v.r_seg_off[n + 1] < s0 || (int64_t)v.r_seg_off[n + 1] > v.n_segs) nex = 0;   // flagged by the walk above
  }
  const uint32_t max_nex = __reduce_max_sync(FULL, nex);
  if (max_nex < 2 || v.n_juncs == 0) return;
  uint32_t j0 = 0, nj = 0;
  float w = 0.f;
  const int32_t* __restrict__ rj = v.rjunc_idx + (s0 - n);
  if (nex >= 2) {
    j0 = v.b_junc_off[b]; nj = v.b_junc_off[b + 1] - j0;
    w = v.r_count[n];
    if ((v.r_flags[n] & STG_RF_UNITIG) && w > 1.f) w = 1.f;
  }
  uint32_t lmax = 0, rmax = 0, rmax_at = 0;  // rmax = max len[i..nex-1], attained (last) at rmax_at
  for (uint32_t i = 1; i < max_nex; i++) {
    uint32_t key = 0x80000000u | (uint32_t)lane;  // lanes without an update this round never match anyone
    bool fl = false, fr = false;
    if (i < nex) {
      const uint32_t l = v.seg_end[s0 + i - 1] - v.seg_start[s0 + i - 1] + 1;
      if (l > lmax) lmax = l;
      if (i == 1 || rmax_at < i) {
        rmax = 0;
        for (uint32_t q = i; q < nex; q++) {
          const uint32_t lq = v.seg_end[s0 + q] - v.seg_start[s0 + q] + 1;
          if (lq >= rmax) { rmax = lq; rmax_at = q; }
        }
      }
      const int32_t jl = rj[i - 1];
      if (jl >= 0 && (uint32_t)jl >= nj) atomicOr(v.err_flag, STG_ERR_JUNC_IDX);
      else if (jl >= 0) {
        const uint32_t j = j0 + (uint32_t)jl;
        const uint32_t anchor = v.junc_anchor[j];
        fl = lmax >= anchor; fr = rmax >= anchor;
        if (fl || fr) key = j;
      }
    }
    // updates with the same row AND the same weight: their sum is (count * w), exact in f64
    const uint32_t grp = __match_any_sync(FULL, ((unsigned long long)key << 32) | (unsigned long long)__float_as_uint(w));
    const uint32_t bl = __ballot_sync(FULL, fl), br = __ballot_sync(FULL, fr);
    if ((fl || fr) && (__ffs(grp) - 1 == lane)) {
      const double wd = (double)w;
      const int nl = __popc(grp & bl), nr = __popc(grp & br), ng = __popc(grp & bl & br);
      if (nl) atomicAdd(v.j_leftsupport + key, (double)nl * wd);
      if (nr) atomicAdd(v.j_rightsupport + key, (double)nr * wd);
      if (ng) atomicAdd(v.j_nreads_good + key, (double)ng * wd);
    }
  }
}

// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(STG_READ_BLOCK) expand_write_kernel(DevBatchView v) {
  const int64_t n64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n64 >= v.n_reads) return;
  const uint32_t n = (uint32_t)n64;
  uint32_t k = v.r_ivl_cnt[n];
  if (k == v.r_ivl_cnt[n + 1]) return;
  const uint32_t b = bundle_of_read(v, n);
  const uint32_t r0 = v.b_read_off[b];
  const uint32_t toff = v.b_tile_off[b];
  const uint32_t refstart = (uint32_t)v.b_start[b];
  const uint32_t len = v.b_len[b];
  walk_read_intervals<false>(v, r0, 0u, n, [&](uint32_t start, uint32_t end, float w, int lane) {
    // cov_edge_add(lane, start-refstart, end-refstart+1, w): +w at index start-refstart+1, -w at end-refstart+2
    uint32_t ia = start - refstart + 1, ib = end - refstart + 2;
    if (start < refstart || ib >= len) { ia = 0; ib = 0; w = 0.f; }  // read outside its bundle: dropped (the
                                                                    // reference would write out of bounds)
    v.ivl[k] = make_uint4(toff + (ia >> STG_TILE_SHIFT), toff + (ib >> STG_TILE_SHIFT),
                          (ia & (STG_TILE_W - 1)) | ((ib & (STG_TILE_W - 1)) << 9) | ((uint32_t)lane << 18),
                          __float_as_uint(w));
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

// Single-pass exclusive prefix sum, in place (chained scan with decoupled look-back): every block scans its chunk,
// publishes its total, and its first warp sums the totals of the blocks before it, 32 at a time, until it meets one
// that already knows its inclusive prefix.  Block ids come from a ticket so that every earlier block is running or
// done; the chunks are uniform streaming work, so nobody waits long (unlike a chained scan inside the read walk).
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p);
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long x);

__global__ void __launch_bounds__(SCAN_THREADS) scan_lookback_kernel(uint32_t* __restrict__ data, int64_t n,
                                                                      unsigned long long* state, uint32_t* ticket) {
  __shared__ uint32_t s_blk, s_prefix;
  if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t blk = s_blk;
  const int lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blk * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t x[SCAN_ITEMS];
  uint32_t acc = 0;
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
      const uint4 t = *reinterpret_cast<const uint4*>(data + base + 4 * q);
      x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) x[it] = base + it < n ? data[base + it] : 0u;
  }
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; it++) acc += x[it];
  uint32_t total;
  const uint32_t excl = block_exclusive<OP_ADD>(acc, &total);
  if (threadIdx.x < 32) {
    uint32_t prefix = 0;   // (flag << 32) | value; flag 1: this block's total, 2: inclusive prefix
    if (blk > 0) {
      if (lane == 0) st_relaxed_u64(state + blk, (1ull << 32) | total);
      int64_t b0 = (int64_t)blk - 1;
      for (;;) {
        const int64_t idx = b0 - lane;
        unsigned long long w = 2ull << 32;               // before the first block: inclusive prefix 0
        if (idx >= 0) {
          w = ld_relaxed_u64(state + idx);
          while ((w >> 32) == 0ull) { __nanosleep(20); w = ld_relaxed_u64(state + idx); }
        }
        const uint32_t incl = __ballot_sync(FULL, (w >> 32) == 2ull);
        const int first = incl ? __ffs(incl) - 1 : 32;
        prefix += __reduce_add_sync(FULL, lane <= first ? (uint32_t)w : 0u);
        if (incl) break;
        b0 -= 32;
      }
    }
    if (lane == 0) { st_relaxed_u64(state + blk, (2ull << 32) | (unsigned long long)(prefix + total)); s_prefix = prefix; }
  }
  __syncthreads();
  uint32_t run = s_prefix + excl;
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
      uint4 t;
      t.x = run; run += x[4 * q]; t.y = run; run += x[4 * q + 1]; t.z = run; run += x[4 * q + 2]; t.w = run; run += x[4 * q + 3];
      *reinterpret_cast<uint4*>(data + base + 4 * q) = t;
    }
  } else {
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; it++) { if (base + it < n) data[base + it] = run; run += x[it]; }
  }
}

cudaError_t launch_scan_exclusive_add(uint32_t* data, int64_t n_plus_1, unsigned long long* state, uint32_t* ticket,
                                      cudaStream_t st) {
  if (n_plus_1 <= 0) return cudaSuccess;
  const int64_t nb = (n_plus_1 + SCAN_CHUNK - 1) / SCAN_CHUNK;
  cudaError_t e = cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (size_t)nb, st);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st)) != cudaSuccess) return e;
  scan_lookback_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(data, n_plus_1, state, ticket);
  return cudaGetLastError();
}


cudaError_t launch_expand_count(const DevBatchView& v, cudaStream_t st) {
  if (v.n_reads == 0) return cudaSuccess;
  expand_count_kernel<<<(unsigned)((v.n_reads + STG_READ_BLOCK - 1) / STG_READ_BLOCK), STG_READ_BLOCK, 0, st>>>(v);
  return cudaGetLastError();
}



cudaError_t launch_expand_write(const DevBatchView& v, cudaStream_t st) {
  if (v.n_reads == 0) return cudaSuccess;
  expand_write_kernel<<<(unsigned)((v.n_reads + STG_READ_BLOCK - 1) / STG_READ_BLOCK), STG_READ_BLOCK, 0, st>>>(v);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K4..K7: stable binning of the deposit events by sub-tile
//
// Interval k contributes two events: +w at (ta, pa) and -w at (tb, pb); events are ordered by (k, endpoint).
// A chunk is STG_CHUNK consecutive intervals, owned by one warp.  For every chunk c and every tile t in the
// chunk's hull [ch_tlo[c], ch_thi[c]] there is one counter cell.  After bin_count a cell holds the number of
// events of chunk c in tile t; bin_columns turns every column into its exclusive prefix over chunks, so that
// bin_scatter can place event e of chunk c at  tile_evoff[t] + cell(c, t) + (rank of e among the events of c in t),
// which is exactly the position of e in the reference-ordered event list of tile t.
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_hull_kernel(DevBatchView v) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= v.n_chunks) return;
  const int lane = threadIdx.x & 31;
  const int64_t k0 = c * STG_CHUNK;
  const int64_t k1 = k0 + STG_CHUNK < v.n_ivl ? k0 + STG_CHUNK : v.n_ivl;
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (int64_t k = k0 + lane; k < k1; k += 32) {
    const uint4 iv = __ldg(v.ivl + k);
    lo = min(lo, iv.x);  // ta <= tb always (the -w deposit lies to the right of the +w one)
    hi = max(hi, iv.y);
  }
  lo = __reduce_min_sync(FULL, lo);
  hi = __reduce_max_sync(FULL, hi);
  if (lane == 0) {
    v.ch_tlo[c] = lo; v.ch_thi[c] = hi;
    v.ch_smin_tlo[c] = lo; v.ch_pmax_thi[c] = hi;
    v.ch_row_off[c] = hi - lo + 1;
    if (c == v.n_chunks - 1) v.ch_row_off[v.n_chunks] = 0;
  }
}

cudaError_t launch_chunk_hull(const DevBatchView& v, cudaStream_t st) {
  if (v.n_chunks == 0) return cudaSuccess;
  chunk_hull_kernel<<<(unsigned)((v.n_chunks * 32 + 255) / 256), 256, 0, st>>>(v);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if ((e = run_scan<OP_ADD, false, true>(v.ch_row_off, v.ch_row_off, v.n_chunks + 1, v.scan_partials, st)) != cudaSuccess) return e;
  if ((e = run_scan<OP_MAX, false, false>(v.ch_pmax_thi, v.ch_pmax_thi, v.n_chunks, v.scan_partials, st)) != cudaSuccess) return e;
  return run_scan<OP_MIN, true, false>(v.ch_smin_tlo, v.ch_smin_tlo, v.n_chunks, v.scan_partials, st);
}

// One warp walks its chunk in order, 32 intervals (64 events) per step; inside a step the events of one tile are
// ranked by (lane, endpoint) with two ballots.  SCATTER = false: count only.
template <bool SCATTER>
__global__ void __launch_bounds__(256) bin_kernel(DevBatchView v) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= v.n_chunks) return;
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t k0 = c * STG_CHUNK;
  const int64_t k1 = k0 + STG_CHUNK < v.n_ivl ? k0 + STG_CHUNK : v.n_ivl;
  const uint32_t tlo = __ldg(v.ch_tlo + c);
  uint32_t* __restrict__ row = v.cells + __ldg(v.ch_row_off + c);
  uint4 cur = make_uint4(0u, 0u, 0u, 0u);
  if (k0 + lane < k1) cur = __ldcs(v.ivl + k0 + lane);
  for (int64_t kb = k0; kb < k1; kb += 32) {
    uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
    if (kb + 32 + lane < k1) nxt = __ldcs(v.ivl + kb + 32 + lane);
    const bool valid = kb + lane < k1;
    const uint32_t ta = cur.x, tb = cur.y;
    bool todo_a = valid, todo_b = valid;
    for (;;) {
      const uint32_t pending = __ballot_sync(FULL, todo_a || todo_b);
      if (!pending) break;
      const int src = __ffs(pending) - 1;
      const uint32_t T = __shfl_sync(FULL, todo_a ? ta : tb, src);
      const bool ina = todo_a && ta == T, inb = todo_b && tb == T;
      const uint32_t ba = __ballot_sync(FULL, ina), bb = __ballot_sync(FULL, inb);
      const uint32_t cntT = __popc(ba) + __popc(bb);
      if (!SCATTER) {
        if (lane == src) atomicAdd(row + (T - tlo), cntT);   // the row is private to this warp: no contention
      } else {
        uint32_t base = 0;
        if (lane == src) base = atomicAdd(row + (T - tlo), cntT) + __ldg(v.tile_evoff + T);
        base = __shfl_sync(FULL, base, src);
        const uint32_t before = base + __popc(ba & lt) + __popc(bb & lt);
        const uint32_t lane_bits = (cur.z >> 18) << 9;
        if (ina) v.events[before] = make_uint2((cur.z & 511u) | lane_bits, cur.w);
        if (inb) v.events[before + (ina ? 1u : 0u)] = make_uint2(((cur.z >> 9) & 511u) | lane_bits, cur.w ^ 0x80000000u);
      }
      todo_a = todo_a && !ina;
      todo_b = todo_b && !inb;
    }
    cur = nxt;
  }
}

cudaError_t launch_bin_count(const DevBatchView& v, cudaStream_t st) {
  if (v.n_chunks == 0) return cudaSuccess;
  bin_kernel<false><<<(unsigned)((v.n_chunks * 32 + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}

cudaError_t launch_bin_scatter(const DevBatchView& v, cudaStream_t st) {
  if (v.n_chunks == 0) return cudaSuccess;
  bin_kernel<true><<<(unsigned)((v.n_chunks * 32 + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}

// first index in [lo, hi) with arr[idx] >= key (arr non-decreasing)
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ arr, uint32_t lo, uint32_t hi,
                                                    uint32_t key) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(arr + mid) >= key) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// One thread per tile: exclusive prefix of its column over the chunks whose hull contains it.
__global__ void __launch_bounds__(256) bin_columns_kernel(DevBatchView v) {
  const int64_t t64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t64 >= v.n_tiles) return;
  const uint32_t t = (uint32_t)t64;
  uint32_t run = 0;
  if (v.n_chunks > 0) {
    const uint32_t nc = (uint32_t)v.n_chunks;
    const uint32_t c0 = lower_bound_u32(v.ch_pmax_thi, 0, nc, t);       // first chunk reaching up to t
    const uint32_t c1 = lower_bound_u32(v.ch_smin_tlo, 0, nc, t + 1);   // from here on every chunk starts past t
    // the loads of a batch of chunks are independent of the running prefix: issue them together
    for (uint32_t cb = c0; cb < c1; cb += 8) {
      uint32_t* cell[8];
      uint32_t x[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const uint32_t c = cb + u;
        cell[u] = nullptr;
        if (c < c1) {
          const uint32_t lo = __ldg(v.ch_tlo + c);
          if (t >= lo && t <= __ldg(v.ch_thi + c)) cell[u] = v.cells + __ldg(v.ch_row_off + c) + (t - lo);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) x[u] = cell[u] ? *cell[u] : 0u;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (cell[u]) { *cell[u] = run; run += x[u]; }
      }
    }
  }
  v.tile_evoff[t] = run;
  v.tile_lastne[t] = run ? t + 1 : 0u;
  if (t64 == v.n_tiles - 1) v.tile_evoff[v.n_tiles] = 0;
}

cudaError_t launch_bin_columns(const DevBatchView& v, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  bin_columns_kernel<<<(unsigned)((v.n_tiles + 255) / 256), 256, 0, st>>>(v);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if ((e = run_scan<OP_ADD, false, true>(v.tile_evoff, v.tile_evoff, v.n_tiles + 1, v.scan_partials, st)) != cudaSuccess) return e;
  return run_scan<OP_MAX, false, false>(v.tile_lastne, v.tile_lastne, v.n_tiles, v.scan_partials, st);
}

// --------------------------------------------------------------------------------------------------------------
// K8: per-tile plan records in ticket order.
// Schedule: sub-tile j of a bundle with nt sub-tiles runs in round floor(j * n_rounds / nt), n_rounds >= max nt.
// Every bundle then advances at the same relative pace from the first ticket to the last (long bundles are not
// left over at the end), a bundle has at most one sub-tile per round, and sub-tile j-1 always holds an earlier
// ticket than sub-tile j — which is what makes waiting on a predecessor's carry deadlock-free.
// --------------------------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void __launch_bounds__(256) tile_plan_kernel(DevBatchView v) {
  const int64_t t64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = t64 < v.n_tiles;
  const uint32_t tid = live ? (uint32_t)t64 : 0u;   // bundle-major tile id
  // key: the round of a sub-tile with events; one common key for all event-free sub-tiles; none for dead lanes
  uint32_t key = 0x80000000u | (uint32_t)lane, b = 0, jt = 0, e0 = 0, nev = 0;
  if (live) {
    e0 = v.tile_evoff[tid];
    nev = v.tile_evoff[tid + 1] - e0;
    uint32_t lo = 0, hi = (uint32_t)v.n_bundles;     // bundle = last b with b_tile_off[b] <= tid
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(v.b_tile_off + mid) <= tid) lo = mid; else hi = mid;
    }
    b = lo;
    const uint32_t first = __ldg(v.b_tile_off + b);
    jt = tid - first;
    const uint32_t nt = __ldg(v.b_tile_off + b + 1) - first;
    key = nev ? (uint32_t)(((uint64_t)jt * (uint64_t)v.n_rounds) / nt) : 0x40000000u;
  }
  // one atomic per distinct key in the warp
  const uint32_t grp = __match_any_sync(FULL, key);
  const int leader = __ffs(grp) - 1;
  uint32_t base = 0;
  if (live && lane == leader) {
    uint32_t* ctr;
    if (key == 0x40000000u) ctr = v.round_cur + v.n_rounds;          // the fill list has its own cursor
    else ctr = (WRITE ? v.round_cur : v.round_off) + key;
    if (WRITE || key != 0x40000000u) base = atomicAdd(ctr, (uint32_t)__popc(grp));
  }
  base = __shfl_sync(FULL, base, leader);
  if (!WRITE || !live) return;
  // tickets of the main kernel: [0, n_main) by round; the event-free sub-tiles follow in any order
  const uint32_t n_main = __ldg(v.round_off + v.n_rounds);
  const uint32_t slot = (nev ? __ldg(v.round_off + key) : n_main) + base + __popc(grp & ((1u << lane) - 1u));
  const uint32_t len = v.b_len[b];
  TileRec r;
  r.gbase = v.b_gbase[b];
  r.i0 = jt * STG_TILE_W;
  r.npos = min((uint32_t)STG_TILE_W, len - r.i0);
  r.stride = v.b_stride[b];
  r.tile_id = tid;
  r.e0 = e0;
  r.nev = nev;
  if (nev >= STG_DEEP_EVENTS) {
    // deep sub-tile: its difference array is prepared by deep_deposit_kernel; the tile kernel loads slot e0
    const uint32_t ds = atomicAdd(v.deep_count, 1u);
    v.deep_list[ds] = make_uint4(tid, e0, nev, 0u);
    r.e0 = ds;
    r.nev = nev | 0x80000000u;
  }
  r.wait_tile = STG_NO_TILE;
  if (jt > 0) {
    const uint32_t last = v.tile_lastne[tid - 1];   // 1 + last tile at or before tid-1 that has events
    if (last > tid - jt) r.wait_tile = last - 1;
  }
  v.tile_plan[slot] = r;
}

cudaError_t launch_tile_plan(const DevBatchView& v, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(v.round_off, 0, sizeof(uint32_t) * (size_t)(v.n_rounds + 1), st);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(v.round_cur, 0, sizeof(uint32_t) * (size_t)(v.n_rounds + 1), st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(v.deep_count, 0, 4 * sizeof(uint32_t), st)) != cudaSuccess) return e;
  tile_plan_kernel<false><<<(unsigned)((v.n_tiles + 255) / 256), 256, 0, st>>>(v);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = run_scan<OP_ADD, false, true>(v.round_off, v.round_off, v.n_rounds + 1, v.scan_partials, st)) != cudaSuccess) return e;
  tile_plan_kernel<true><<<(unsigned)((v.n_tiles + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K8b: deep sub-tiles (>= STG_DEEP_EVENTS events; in RNA-seq a third of all events sit in a fraction of a percent
// of the sub-tiles, and a single position — a splice site of a highly expressed gene — can receive 10^5 of them).
// Chaining such a position through warp shuffles costs ~12 instructions per event on one warp: milliseconds.
// Instead one CTA sorts the sub-tile's event list by position (stable: 16 warps x 512 positions counting sort in
// shared memory, each warp owning a contiguous slice of the list), and then every position's events are summed in
// a register by one thread reading them in reference order — the bare dependent-add chain the arithmetic needs.
// Output: the finished difference array d[3][512] of the sub-tile (deep_d), consumed by cov_tile_kernel.
// --------------------------------------------------------------------------------------------------------------
constexpr int DEEP_WARPS = 16;
static_assert(DEEP_WARPS * 32 == STG_TILE_W, "phase 2 of deep_deposit_kernel maps one thread to one position");

__global__ void __launch_bounds__(DEEP_WARPS * 32) deep_deposit_kernel(DevBatchView v) {
  __shared__ uint32_t cnt[DEEP_WARPS][STG_TILE_W];       // per warp slice: events per position, then first slot
  __shared__ uint32_t poff[STG_TILE_W + 1];              // first slot of every position in the sorted list
  __shared__ uint32_t wtot[DEEP_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  __shared__ uint32_t next_tile;
  const uint32_t ndeep = *v.deep_count;
  for (;;) {
    if (threadIdx.x == 0) next_tile = atomicAdd(v.tile_ticket + 1, 1u);
    __syncthreads();
    const uint32_t di = next_tile;
    if (di >= ndeep) break;
    const uint4 rec = v.deep_list[di];
    const uint32_t e0 = rec.y, nev = rec.z;
    const uint2* __restrict__ ev = v.events + e0;
    uint2* __restrict__ sorted = v.events2 + e0;
    const uint32_t steps = (nev + 31) / 32, per = (steps + DEEP_WARPS - 1) / DEEP_WARPS;
    const uint32_t s_lo = min(steps, (uint32_t)wid * per), s_hi = min(steps, s_lo + per);
    // phase 1: per-slice histogram
    for (int q = lane; q < STG_TILE_W; q += 32) cnt[wid][q] = 0u;
    __syncwarp();
    for (uint32_t st = s_lo; st < s_hi; st += 4) {
      uint32_t px[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t e = (st + u) * 32 + lane;
        px[u] = (st + u < s_hi && e < nev) ? (__ldg(&ev[e].x) & 511u) : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) if (px[u] != 0xffffffffu) atomicAdd(&cnt[wid][px[u]], 1u);
    }
    __syncthreads();
    // phase 2: thread p owns position p: exclusive prefix down the slices, then across positions
    {
      const int p = threadIdx.x;
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < DEEP_WARPS; w++) { const uint32_t x = cnt[w][p]; cnt[w][p] = run; run += x; }
      uint32_t incl = run;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
      if (lane == 31) wtot[wid] = incl;
      __syncthreads();
      uint32_t wbase = 0;
      for (int w = 0; w < wid; w++) wbase += wtot[w];
      const uint32_t excl = wbase + incl - run;
      poff[p] = excl;
      if (p == STG_TILE_W - 1) poff[STG_TILE_W] = excl + run;
#pragma unroll
      for (int w = 0; w < DEEP_WARPS; w++) cnt[w][p] += excl;
    }
    __syncthreads();
    // phase 3: stable scatter of this warp's slice, 32 events per step in order
    {
      uint2 x = make_uint2(0u, 0u);
      if (s_lo < s_hi && s_lo * 32 + lane < nev) x = __ldg(&ev[s_lo * 32 + lane]);
      for (uint32_t st = s_lo; st < s_hi; st++) {
        const bool valid = st * 32 + lane < nev;
        uint2 xn = make_uint2(0u, 0u);
        if (st + 1 < s_hi && (st + 1) * 32 + lane < nev) xn = __ldg(&ev[(st + 1) * 32 + lane]);
        const uint32_t pos = x.x & 511u;
        const uint32_t grp = __match_any_sync(FULL, valid ? pos : (0x8000u | (uint32_t)lane));
        const int leader = __ffs(grp) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) { base = cnt[wid][pos]; cnt[wid][pos] = base + __popc(grp); }
        base = __shfl_sync(FULL, base, leader);
        if (valid) sorted[base + __popc(grp & lt)] = make_uint2(x.x >> 9, x.y);
        __syncwarp();
        x = xn;
      }
    }
    __syncthreads();   // also orders the global writes of `sorted` before the reads below (same CTA)
    // phase 4: every position's events summed in reference order (cov_edge_add, rlink.cpp:133-142: the strand's own
    // lane, and lane 1 as well).  Short lists: one thread per position walks its list.  Long lists (a splice site of
    // a deep locus): the warp loads 32 events at a time and lanes 0..2 run the bare chain of dependent adds.
    float* dd = v.deep_d + (size_t)di * (3 * STG_TILE_W);
    {
      const int p = threadIdx.x;
      const uint32_t l0 = poff[p], n = poff[p + 1] - l0;
      const bool is_long = n > 64u;
      if (!is_long) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (uint32_t i = 0; i < n; i += 8) {   // 8 loads in flight, then the chain
          uint2 x[8];
#pragma unroll
          for (int u = 0; u < 8; u++) x[u] = i + u < n ? sorted[l0 + i + u] : make_uint2(1u, 0u);  // pad: +0.0, a no-op
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const float val = __uint_as_float(x[u].y);
            if (x[u].x == 0u) a0 = __fadd_rn(a0, val);
            if (x[u].x == 2u) a2 = __fadd_rn(a2, val);
            a1 = __fadd_rn(a1, val);
          }
        }
        dd[p] = a0; dd[STG_TILE_W + p] = a1; dd[2 * STG_TILE_W + p] = a2;
      }
      else {
        // handed to deep_long_kernel: the CTA must not idle behind one position's chain of 10^4..10^5 dependent adds
        // (lists long enough for a whole CTA are collected from the far end of the array)
        const bool big = n >= (uint32_t)EC_CTA_MIN;
        const uint32_t k = atomicAdd(v.deep_count + (big ? 2 : 1), 1u);
        v.deep_long[big ? v.deep_long_cap - 1u - k : k] = make_uint4(di, (uint32_t)p, e0 + l0, n);
      }
    }
    __syncthreads();
  }
}

// Long per-position lists of deep sub-tiles (> 64 events on one position: a splice site of a deep locus).  The three
// strand lanes of a position are three chains of dependent float adds in reference order (cov_edge_add,
// rlink.cpp:133-142: the strand's own lane and lane 1; the other lanes get +0.0, a no-op since no accumulator is ever
// -0); each (list, strand lane) gets a warp, which adds the list through warp_chain_list (exact_chain.cuh): inside a
// binade of the running sum the chain is an integer sum of per-event increments that the warp takes 256 at a time with
// one scan; a block that leaves the binade is added the plain way.
struct EcEventList {   // strand lane `c` of a position's event list {tag, value}
  const uint2* list; int n; uint32_t c;
  __device__ __forceinline__ void load(int rel, float (&x)[8]) const {
    uint2 t[8];
    ec_load8(list, n, rel, t, make_uint2(1u, 0u));
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = (c == 1u || t[e].x == c) ? __uint_as_float(t[e].y) : 0.f;
  }
  __device__ __forceinline__ void prefetch(int rel) const { if (rel < n) ec_prefetch_l2(list + rel); }
};

__global__ void __launch_bounds__(256) deep_long_kernel(DevBatchView v) {
  const int lane = threadIdx.x & 31;
  // the few very long lists first, a CTA per (list, strand lane): every warp prepares blocks (cta_chain_list)
  {
    __shared__ EcCtaState cst;
    __shared__ uint32_t ticket;
    const uint32_t nbig = v.deep_count[2] * 3u;
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) ticket = atomicAdd(v.tile_ticket + 4, 1u);
      __syncthreads();
      const uint32_t wi = ticket;
      if (wi >= nbig) break;
      const uint4 rec = v.deep_long[v.deep_long_cap - 1u - wi / 3u];
      const EcEventList ld{v.events2 + rec.z, (int)rec.w, wi % 3u};
      const float acc = cta_chain_list(0.f, ld.n, ec_misalign8(ld.list), ld, cst);
      if (threadIdx.x == 0) v.deep_d[(size_t)rec.x * (3 * STG_TILE_W) + ld.c * STG_TILE_W + rec.y] = acc;
    }
  }
  const uint32_t nwork = v.deep_count[1] * 3u;
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(v.tile_ticket + 2, 1u);
    wi = __shfl_sync(FULL, wi, 0);
    if (wi >= nwork) break;
    const uint4 rec = v.deep_long[wi / 3u];
    const EcEventList ld{v.events2 + rec.z, (int)rec.w, wi % 3u};
    const float acc = warp_chain_list(0.f, ld.n, ec_misalign8(ld.list), ld, lane);
    if (lane == 0) v.deep_d[(size_t)rec.x * (3 * STG_TILE_W) + ld.c * STG_TILE_W + rec.y] = acc;
  }
}

// Self-test of warp_chain_add (stg_selftest_chain): one warp per list of a CSR, the list added to s0 in order; PLAIN = the
// bare chain of dependent adds (the timing baseline).
template <bool PLAIN>
__global__ void __launch_bounds__(128) chain_selftest_kernel(int64_t nl, const uint32_t* __restrict__ off, const float* __restrict__ x,
                                                             const float* __restrict__ s0, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t li = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (li >= nl) return;
  const float* list = x + off[li];
  const int n = (int)(off[li + 1] - off[li]);
  float s = s0[li];
  if (PLAIN) {
    for (int j0 = 0; j0 < n; j0 += 32) {
      const float xv = j0 + lane < n ? list[j0 + lane] : 0.f;
      const int m = min(32, n - j0);
      for (int l = 0; l < m; l++) s = __fadd_rn(s, __shfl_sync(FULL, xv, l));
    }
  } else {
    const EcFloatList ld{list, n};
    s = warp_chain_list(s, n, ec_misalign8(list), ld, lane);
  }
  if (lane == 0) out[li] = s;
}

// ... and of cta_chain_list: one CTA per list
__global__ void __launch_bounds__(256) chain_selftest_cta_kernel(int64_t nl, const uint32_t* __restrict__ off, const float* __restrict__ x,
                                                                 const float* __restrict__ s0, float* __restrict__ out) {
  __shared__ EcCtaState cst;
  for (int64_t li = blockIdx.x; li < nl; li += gridDim.x) {
    const EcFloatList ld{x + off[li], (int)(off[li + 1] - off[li])};
    const float s = cta_chain_list(s0[li], ld.n, ec_misalign8(ld.list), ld, cst);
    if (threadIdx.x == 0) out[li] = s;
  }
}

cudaError_t launch_chain_selftest(int64_t n, const uint32_t* off, const float* x, const float* s0, float* out, int plain, cudaStream_t st) {
  if (n && plain == 2) {
    chain_selftest_cta_kernel<<<(unsigned)(n < 4096 ? n : 4096), 256, 0, st>>>(n, off, x, s0, out);
    return cudaGetLastError();
  }
  if (n == 0) return cudaSuccess;
#ifdef EC_STATS
  {
    unsigned long long h[4];
    cudaMemcpyFromSymbol(h, ec_stats, sizeof h);
    fprintf(stderr, "ec_stats so far: integer %llu plain %llu (right binade %llu, bad lanes %llu)\n", h[0], h[1], h[2], h[3]);
  }
#endif
  const unsigned grid = (unsigned)((n * 32 + 127) / 128);
  if (plain) chain_selftest_kernel<true><<<grid, 128, 0, st>>>(n, off, x, s0, out);
  else chain_selftest_kernel<false><<<grid, 128, 0, st>>>(n, off, x, s0, out);
  return cudaGetLastError();
}

cudaError_t launch_deep_deposit(const DevBatchView& v, int num_sms, cudaStream_t st) {
  if (v.n_tiles == 0 || v.n_ivl == 0) return cudaSuccess;
  deep_deposit_kernel<<<num_sms * 4, DEEP_WARPS * 32, 0, st>>>(v);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  deep_long_kernel<<<num_sms * 8, 256, 0, st>>>(v);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K9: the tile kernel
// --------------------------------------------------------------------------------------------------------------
constexpr int TW = STG_TILE_W;          // positions per sub-tile
constexpr int TWS = TW + 4;             // lane stride in shared memory (+4 floats: the three chain threads hit
                                        // different banks with their 128-bit accesses)
constexpr int TNW = TW / 32;            // 32-position words per sub-tile
constexpr int TILE_SMEM_WORDS = 3 * TWS + 3 * TNW + 3 * TNW;  // d, masks, fills

int cov_tile_smem_bytes() { return STG_TILE_WARPS * TILE_SMEM_WORDS * 4; }

constexpr unsigned long long CARRY_EMPTY = 0xffffffffffffffffull;  // never a valid (c, bs) pair: both are >= 0

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long x;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long x) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory");
}

// m repeated additions bs = fl(bs + c), c >= 0, in O(binades crossed) instead of O(m).
// Inside one binade a float's bit pattern is linear in its value, and from the second addition inside the binade
// on, every addition of the same c moves the pattern by the same integer (round-to-nearest-even settles after one
// step: the tie case leaves an even mantissa behind).  So: step explicitly until three consecutive sums share a
// binade, then jump to the last sum that still does.  Checked against the plain loop on 3*10^7 random cases
// (tests/test_advance.c) including ties, subnormals and binade crossings.
__device__ __forceinline__ float advance_bs(float bs, float c, uint32_t m) {
  while (m) {
    const float y1 = __fadd_rn(bs, c); m--; if (!m) return y1;
    const float y2 = __fadd_rn(y1, c); m--; if (!m) return y2;
    const uint32_t b1 = __float_as_uint(y1), b2 = __float_as_uint(y2);
    if ((b1 ^ b2) & 0xff800000u) { bs = y2; continue; }
    const float y3 = __fadd_rn(y2, c); m--;
    const uint32_t b3 = __float_as_uint(y3);
    if ((b3 ^ b2) & 0xff800000u) { bs = y3; continue; }
    const uint32_t inc = b3 - b2;
    if (inc == 0) return y3;
    const uint32_t top = (b3 & 0xff800000u) + 0x00800000u;
    uint32_t t = (top - b3) / inc;
    if (t > m) t = m;
    bs = __uint_as_float(b3 + t * inc); m -= t;
  }
  return bs;
}

// The same, writing every intermediate sum: dl[k] for k in [k, k1) with the whole warp (arguments warp-uniform).
__device__ __forceinline__ void fill_run(float* dl, uint32_t k, uint32_t k1, float c, float& bs, int lane) {
  while (k < k1) {
    const float y1 = __fadd_rn(bs, c); if (lane == 0) dl[k] = y1; k++; bs = y1; if (k == k1) break;
    const float y2 = __fadd_rn(y1, c); if (lane == 0) dl[k] = y2; k++; bs = y2; if (k == k1) break;
    const uint32_t b1 = __float_as_uint(y1), b2 = __float_as_uint(y2);
    if ((b1 ^ b2) & 0xff800000u) continue;
    const float y3 = __fadd_rn(y2, c); if (lane == 0) dl[k] = y3; k++; bs = y3;
    const uint32_t b3 = __float_as_uint(y3);
    if ((b3 ^ b2) & 0xff800000u) continue;
    const uint32_t inc = b3 - b2;
    uint32_t t = k1 - k;
    if (inc) {
      const uint32_t top = (b3 & 0xff800000u) + 0x00800000u;
      const uint32_t tm = (top - b3) / inc;
      if (tm < t) t = tm;
    }
    for (uint32_t j = lane + 1; j <= t; j += 32) dl[k + j - 1] = __uint_as_float(b3 + j * inc);
    bs = __uint_as_float(b3 + t * inc); k += t;
  }
}

__global__ void __launch_bounds__(STG_TILE_WARPS * 32, 1) cov_tile_kernel(DevBatchView v) {
  extern __shared__ __align__(16) uint32_t smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* d = reinterpret_cast<float*>(smem_raw) + (size_t)wid * TILE_SMEM_WORDS;  // d[3][TWS]
  uint32_t* masks = reinterpret_cast<uint32_t*>(d + 3 * TWS);                     // masks[3][TNW]
  float* fills = reinterpret_cast<float*>(masks + 3 * TNW);                       // fills[3][TNW]
  const uint2 no_event = make_uint2(0u, 0u);
  const uint32_t n_main = __ldg(v.round_off + v.n_rounds);   // sub-tiles with events: tickets [0, n_main)

  // A warp takes a ticket only when it is free to work on it at once: a sub-tile queued behind unrelated work would
  // hold up every later sub-tile of its bundle (measured: taking tickets ahead doubled the kernel time).
  for (;;) {
    uint32_t t_cur = 0;
    if (lane == 0) t_cur = atomicAdd(v.tile_ticket, 1u);
    t_cur = __shfl_sync(FULL, t_cur, 0);
    if (t_cur >= n_main) break;
    uint32_t recw = 0;
    if (lane < 8) recw = __ldg(reinterpret_cast<const uint32_t*>(v.tile_plan + t_cur) + lane);

    const uint32_t gbase = __shfl_sync(FULL, recw, 0), i0 = __shfl_sync(FULL, recw, 1);
    const uint32_t e0 = __shfl_sync(FULL, recw, 2), nev_raw = __shfl_sync(FULL, recw, 3);
    const uint32_t stride = __shfl_sync(FULL, recw, 4), npos = __shfl_sync(FULL, recw, 5);
    const uint32_t tile_id = __shfl_sync(FULL, recw, 6), wait_tile = __shfl_sync(FULL, recw, 7);
    const uint32_t to_reset = (STG_BSIZE - i0 % STG_BSIZE) % STG_BSIZE;  // in-tile position where idx % BSIZE == 0
    float* out = v.bpcov + (size_t)gbase + i0;           // strand lane s of the batch starts at s * cov_pitch
    const uint32_t n4 = min((uint32_t)TW, stride - i0) / 4;              // float4 per lane incl. the zero padding

    const bool deep = (nev_raw & 0x80000000u) != 0u;     // difference array precomputed by deep_deposit_kernel
    const uint32_t nev = deep ? 0u : nev_raw;
    uint2 cur = no_event;
    if (lane < nev) cur = __ldcs(v.events + e0 + lane);

    // state at the end of the previous position (rlink.cpp:17037-17063): c = clamped coverage, bs = running block sum.
    // It is published by the nearest earlier sub-tile that had events; over the event-free positions in between c
    // stays put and bs is c added once per position (restarting from 0 at a block start): advance_bs.
    float c = 0.f, bs = 0.f;
    auto fetch_state = [&]() {
      if (wait_tile == STG_NO_TILE) return;            // nothing deposited before this sub-tile: c = bs = 0
      if (lane < 3) {
        const unsigned long long* p = v.tile_carry + (size_t)wait_tile * 4 + lane;
        unsigned long long x = ld_relaxed_u64(p);
        uint32_t ns = 32;
        while (x == CARRY_EMPTY) { __nanosleep(ns); if (ns < 256u) ns <<= 1; x = ld_relaxed_u64(p); }
        c = __uint_as_float((uint32_t)x); bs = __uint_as_float((uint32_t)(x >> 32));
        const uint32_t first = tile_id - i0 / TW;                        // first sub-tile of the bundle
        const uint32_t wlast = (wait_tile - first) * TW + TW - 1;        // last index of the sub-tile waited on
        const uint32_t gap = i0 - 1 - wlast;
        if (gap) {
          const uint32_t rst = (i0 - 1) / STG_BSIZE * STG_BSIZE;         // last block start at or before i0-1
          bs = rst > wlast ? advance_bs(0.f, c, i0 - rst) : advance_bs(bs, c, gap);
        }
      }
      __syncwarp();
    };
    auto write_out = [&](uint32_t fillmask) {
      // finished bpcov of the sub-tile: one streaming, coalesced write per strand lane (incl. the zero padding);
      // constant words are taken from `fills` instead of shared memory
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const uint32_t fm = __shfl_sync(FULL, fillmask, s);
        const float4* src = reinterpret_cast<const float4*>(d + s * TWS);
        float4* dst = reinterpret_cast<float4*>(out + (size_t)s * v.cov_pitch);
        for (uint32_t q = lane; q < n4; q += 32) {
          float4 x;
          if ((fm >> (q >> 3)) & 1u) { const float f = fills[s * TNW + (q >> 3)]; x = make_float4(f, f, f, f); }
          else x = src[q];
          __stcs(dst + q, x);
        }
      }
      __syncwarp();
    };

    {
      // ---- sub-tile with events ----
      if (deep) {
        const float4* src = reinterpret_cast<const float4*>(v.deep_d + (size_t)e0 * (3 * TW));
        float4* d4 = reinterpret_cast<float4*>(d);
        for (int q = lane; q < 3 * TW / 4; q += 32) d4[(q / (TW / 4)) * (TWS / 4) + (q % (TW / 4))] = __ldcs(src + q);
        __syncwarp();
        for (int q = 0; q < TNW; q++) {
#pragma unroll
          for (int s = 0; s < 3; s++) {
            const uint32_t m = __ballot_sync(FULL, d[s * TWS + q * 32 + lane] != 0.f);
            if (lane == 0) masks[s * TNW + q] = m;
          }
        }
      } else {
        float4* d4 = reinterpret_cast<float4*>(d);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = lane; q < 3 * TWS / 4; q += 32) d4[q] = z;
        for (int q = lane; q < 3 * TNW; q += 32) masks[q] = 0u;
      }
      __syncwarp();

      // ordered deposit of the event list.  Events of one 32-group that share a position are chained, in lane
      // (= reference) order, in the registers of the first of them; groups are applied one after the other.
      for (uint32_t eb = 0; eb < nev; eb += 32) {
        uint2 nxt = no_event;
        if (eb + 32 + lane < nev) nxt = __ldcs(v.events + e0 + eb + 32 + lane);
        const bool valid = eb + lane < nev;
        const uint32_t pos = cur.x & 511u, sl = cur.x >> 9;
        const float val = __uint_as_float(cur.y);
        const uint32_t grp = __match_any_sync(FULL, valid ? pos : (0x8000u | (uint32_t)lane));
        const uint32_t is0 = __ballot_sync(FULL, valid && sl == 0u), is2 = __ballot_sync(FULL, valid && sl == 2u);
        const bool leader = valid && (__ffs(grp) - 1 == lane);
        const int maxn = __reduce_max_sync(FULL, valid ? __popc(grp) : 0);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (leader) { a0 = d[pos]; a1 = d[TWS + pos]; a2 = d[2 * TWS + pos]; }
        uint32_t rem = leader ? grp : 0u;
        for (int r = 0; r < maxn; r++) {
          const int src = __ffs(rem) - 1;                 // -1 for idle lanes: the shuffle wraps, value unused
          const float vv = __shfl_sync(FULL, val, src);
          if (rem) {
            // cov_edge_add (rlink.cpp:133-142): the strand's own lane, and lane 1 as well when the strand is known
            if ((is0 >> src) & 1u) a0 = __fadd_rn(a0, vv);
            if ((is2 >> src) & 1u) a2 = __fadd_rn(a2, vv);
            a1 = __fadd_rn(a1, vv);
            rem &= rem - 1;
          }
        }
        if (leader) {
          const uint32_t bit = 1u << (pos & 31u), wq = pos >> 5;
          if (grp & is0) { d[pos] = a0; atomicOr(masks + wq, bit); }
          d[TWS + pos] = a1; atomicOr(masks + TNW + wq, bit);
          if (grp & is2) { d[2 * TWS + pos] = a2; atomicOr(masks + 2 * TNW + wq, bit); }
        }
        __syncwarp();
        cur = nxt;
      }

      fetch_state();

      // the two recurrences (rlink.cpp:17037-17063), one strand lane per thread:
      //   c = max(0, c + d[i]);  bs = (i % BSIZE == 0 ? 0 : bs) + c;  out[i] = bs
      const int nwords = (int)((npos + 31) / 32);
      uint32_t fillmask = 0;
      if (lane < 3) {
        float* dl = d + lane * TWS;
        const uint32_t* ml = masks + lane * TNW;
        float* fl = fills + lane * TNW;
        for (int q = 0; q < nwords; q++) {
          const uint32_t base = (uint32_t)q * 32;
          const bool full = base + 32 <= npos;
          const bool has_reset = to_reset >= base && to_reset < base + 32;
          if (full && !has_reset) {
            if (ml[q] == 0) {
              if (__fadd_rn(bs, c) == bs) {  // nothing deposited and c too small to move bs: constant word
                fl[q] = bs;
                fillmask |= 1u << q;
                continue;
              }
              // nothing deposited: c = max(0, c + 0) stays put (c >= 0, never -0), only the block sum runs
              float4 y[8];
#pragma unroll
              for (int u = 0; u < 8; u++) {
                bs = __fadd_rn(c, bs); y[u].x = bs;
                bs = __fadd_rn(c, bs); y[u].y = bs;
                bs = __fadd_rn(c, bs); y[u].z = bs;
                bs = __fadd_rn(c, bs); y[u].w = bs;
              }
#pragma unroll
              for (int u = 0; u < 8; u++) *reinterpret_cast<float4*>(dl + base + u * 4) = y[u];
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
        // publish (c, bs) after the last position: one 64-bit store, data and "ready" in the same word
        st_relaxed_u64(v.tile_carry + (size_t)tile_id * 4 + lane,
                       ((unsigned long long)__float_as_uint(bs) << 32) | (unsigned long long)__float_as_uint(c));
      }
      __syncwarp();
      write_out(fillmask);
    }

  }
}

// --------------------------------------------------------------------------------------------------------------
// K10: event-free sub-tiles.  Runs after cov_tile_kernel, so every carry it needs is final: no waiting, no shared
// memory, full occupancy — a streaming fill.  One warp per sub-tile.
// --------------------------------------------------------------------------------------------------------------
// one lane: [dst, dst + bytes) <- shared memory, as one asynchronous bulk copy (bytes and both addresses multiples of 16)
__device__ __forceinline__ void bulk_store(float* dst, const float* src_smem, uint32_t bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sa), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(256) cov_fill_kernel(DevBatchView v) {
  const int lane = threadIdx.x & 31;
  __shared__ __align__(128) float zero_row[TW];
  __shared__ __align__(128) float warp_rows[8][TW];
  float* warp_row = warp_rows[threadIdx.x >> 5];
  for (int q = threadIdx.x; q < TW; q += blockDim.x) zero_row[q] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t n_main = __ldg(v.round_off + v.n_rounds);
  const uint32_t n_tiles = (uint32_t)v.n_tiles;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  // software pipeline, two deep: while sub-tile t is written, the plan record of t + 2*warps and the carry of
  // t + warps (whose record arrived during the previous iteration) are in flight
  auto load_rec = [&](uint32_t tt) -> uint32_t {
    return (tt < n_tiles && lane < 8) ? __ldg(reinterpret_cast<const uint32_t*>(v.tile_plan + tt) + lane) : 0u;
  };
  auto load_carry = [&](uint32_t rec) -> unsigned long long {
    const uint32_t wt = __shfl_sync(FULL, rec, 7);
    return (wt != STG_NO_TILE && lane < 3) ? __ldcg(v.tile_carry + (size_t)wt * 4 + lane) : 0ull;
  };
  uint32_t t = n_main + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint32_t rec_a = load_rec(t), rec_b = load_rec(t + warps);
  unsigned long long cw_a = t < n_tiles ? load_carry(rec_a) : 0ull;
  for (; t < n_tiles; t += warps) {
    const uint32_t rec_c = load_rec(t + 2 * warps);
    const unsigned long long cw_b = t + warps < n_tiles ? load_carry(rec_b) : 0ull;
    const uint32_t recw = rec_a;
    const unsigned long long x = cw_a;
    rec_a = rec_b; rec_b = rec_c; cw_a = cw_b;
    const uint32_t gbase = __shfl_sync(FULL, recw, 0), i0 = __shfl_sync(FULL, recw, 1);
    const uint32_t stride = __shfl_sync(FULL, recw, 4), npos = __shfl_sync(FULL, recw, 5);
    const uint32_t tile_id = __shfl_sync(FULL, recw, 6), wait_tile = __shfl_sync(FULL, recw, 7);
    const uint32_t to_reset = (STG_BSIZE - i0 % STG_BSIZE) % STG_BSIZE;  // in-tile position where idx % BSIZE == 0
    float* out = v.bpcov + (size_t)gbase + i0;
    const uint32_t n4 = min((uint32_t)TW, stride - i0) / 4;              // float4 per lane incl. the zero padding
    // state before the first position: from the nearest earlier sub-tile with events, advanced over the gap
    float c = 0.f, bs = 0.f;
    if (wait_tile != STG_NO_TILE && lane < 3) {
      c = __uint_as_float((uint32_t)x); bs = __uint_as_float((uint32_t)(x >> 32));
      const uint32_t first = tile_id - i0 / TW;
      const uint32_t wlast = (wait_tile - first) * TW + TW - 1;
      const uint32_t gap = i0 - 1 - wlast;
      if (gap) {
        const uint32_t rst = (i0 - 1) / STG_BSIZE * STG_BSIZE;
        bs = rst > wlast ? advance_bs(0.f, c, i0 - rst) : advance_bs(bs, c, gap);
      }
    }
    const bool flat = lane >= 3 || (to_reset < npos ? c == 0.f : __fadd_rn(bs, c) == bs);
    if (__all_sync(FULL, flat)) {
      // every position repeats bs (0 from a block start on): each strand lane of the sub-tile leaves as ONE bulk copy
      // shared -> global (cp.async.bulk, issued by one lane) instead of 32-lane store loops.  Rows that are all zero —
      // most of them: the gaps between loci — come straight from the CTA's zero row; other rows are laid out in the warp's
      // own row first (the async proxy must have read it before it is written again: wait_group.read).
      const float b0 = __shfl_sync(FULL, bs, 0), b1 = __shfl_sync(FULL, bs, 1), b2 = __shfl_sync(FULL, bs, 2);
      const uint32_t lim = to_reset < npos ? to_reset : npos;          // positions [0, lim) hold bs, the rest 0
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const float bv = s == 0 ? b0 : (s == 1 ? b1 : b2);
        float* dst = out + (size_t)s * v.cov_pitch;
        if (bv == 0.f || lim == 0u) {
          if (lane == 0) bulk_store(dst, zero_row, n4 * 16u);
        } else {
          float4* row4 = reinterpret_cast<float4*>(warp_row);
          for (uint32_t q = lane; q < n4; q += 32) {
            const uint32_t p = q * 4;
            float4 x;
            x.x = p + 0 < lim ? bv : 0.f; x.y = p + 1 < lim ? bv : 0.f; x.z = p + 2 < lim ? bv : 0.f; x.w = p + 3 < lim ? bv : 0.f;
            row4[q] = x;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) { bulk_store(dst, warp_row, n4 * 16u); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
          __syncwarp();
        }
      }
    } else {
      // a coverage residue travels through: bs grows by the same c at every position -> closed-form runs
#pragma unroll
      for (int s = 0; s < 3; s++) {
        const float cs = __shfl_sync(FULL, c, s);
        float bss = __shfl_sync(FULL, bs, s);
        float* dl = out + (size_t)s * v.cov_pitch;
        if (to_reset < npos) {
          fill_run(dl, 0, to_reset, cs, bss, lane);
          bss = 0.f;
          fill_run(dl, to_reset, npos, cs, bss, lane);
        } else {
          fill_run(dl, 0, npos, cs, bss, lane);
        }
        for (uint32_t q = npos + lane; q < n4 * 4; q += 32) dl[q] = 0.f;
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

cudaError_t launch_cov_fill(const DevBatchView& v, int num_sms, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  int64_t want = (v.n_tiles + 7) / 8;
  int grid = (int)(want < (int64_t)num_sms * 8 ? want : (int64_t)num_sms * 8);
  cov_fill_kernel<<<grid, 256, 0, st>>>(v);
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
  const float* base = v.bpcov + v.b_gbase[b];
  const int s = lane[i];
  if (!sign) out[i] = dev_get_cov(base + (size_t)s * v.cov_pitch, start[i], end[i]);
  else out[i] = dev_get_cov_sign(base + v.cov_pitch, base + (size_t)(2 - s) * v.cov_pitch, start[i], end[i]);
}

cudaError_t launch_query_cov(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* lane,
                             const uint32_t* start, const uint32_t* end, int sign, double* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  query_cov_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, bundle, lane, start, end, sign, out);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------------------
// K11b: bpcov for the trip to the host.  The host needs all of bpcov (printResults, rlink.cpp:18402-20260, reads the three
// lanes), but most of it says nothing: four fifths of the sub-tiles saw no event and repeat ONE value per strand lane (the
// block sum at a standstill between loci and inside introns).  A strand lane of a sub-tile either travels as that value or
// as its 512 floats, gathered into one dense buffer: cov_pack_flag marks the lanes that change inside the sub-tile, a scan
// numbers them, cov_pack_gather copies them; stg_unpack_bpcov lays the bundle's arrays out again on the host (a fill or
// a copy per sub-tile lane: what the host adaptor's copy into BundleData::bpcov costs anyway).
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cov_pack_flag_kernel(DevBatchView v, uint32_t* __restrict__ flag, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per sub-tile (plan record)
  if (w >= v.n_tiles) return;
  uint32_t recw = 0;
  if (lane < 8) recw = __ldg(reinterpret_cast<const uint32_t*>(v.tile_plan + w) + lane);
  const uint32_t gbase = __shfl_sync(FULL, recw, 0), i0 = __shfl_sync(FULL, recw, 1);
  const uint32_t npos = __shfl_sync(FULL, recw, 5), tile_id = __shfl_sync(FULL, recw, 6);
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(v.bpcov + (size_t)s * v.cov_pitch + gbase + i0);
    const uint32_t x0 = __ldg(p);
    bool same = true;
    for (uint32_t q = lane; q * 4 < npos; q += 32) {
      const uint4 y = __ldg(reinterpret_cast<const uint4*>(p) + q);
      const uint32_t k = q * 4;
      same = same && y.x == x0 && (k + 1 >= npos || y.y == x0) && (k + 2 >= npos || y.z == x0) && (k + 3 >= npos || y.w == x0);
    }
    same = __all_sync(FULL, same);
    if (lane == 0) { flag[(size_t)s * v.n_tiles + tile_id] = same ? 0u : 1u; val[(size_t)s * v.n_tiles + tile_id] = __uint_as_float(x0); }
  }
}

// slot[i] = exclusive prefix of the flags (slot[3 * n_tiles] = lanes that travel in full); desc[i] = slot or 0xffffffff
__global__ void __launch_bounds__(256) cov_pack_gather_kernel(DevBatchView v, const uint32_t* __restrict__ slot, uint32_t* __restrict__ desc,
                                                              float* __restrict__ raw) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= v.n_tiles) return;
  uint32_t recw = 0;
  if (lane < 8) recw = __ldg(reinterpret_cast<const uint32_t*>(v.tile_plan + w) + lane);
  const uint32_t gbase = __shfl_sync(FULL, recw, 0), i0 = __shfl_sync(FULL, recw, 1);
  const uint32_t npos = __shfl_sync(FULL, recw, 5), tile_id = __shfl_sync(FULL, recw, 6);
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const size_t i = (size_t)s * v.n_tiles + tile_id;
    const uint32_t a = __ldg(slot + i), z = __ldg(slot + i + 1);
    if (lane == 0) desc[i] = z != a ? a : 0xffffffffu;
    if (z == a) continue;
    const uint4* p = reinterpret_cast<const uint4*>(v.bpcov + (size_t)s * v.cov_pitch + gbase + i0);
    uint4* o = reinterpret_cast<uint4*>(raw + (size_t)a * STG_TILE_W);
    for (uint32_t q = lane; q * 4 < npos; q += 32) __stcs(o + q, __ldcs(p + q));
  }
}

cudaError_t launch_cov_pack_flag(const DevBatchView& v, uint32_t* flag, float* val, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  cov_pack_flag_kernel<<<(unsigned)((v.n_tiles * 32 + 255) / 256), 256, 0, st>>>(v, flag, val);
  return cudaGetLastError();
}
cudaError_t launch_cov_pack_gather(const DevBatchView& v, const uint32_t* slot, uint32_t* desc, float* raw, cudaStream_t st) {
  if (v.n_tiles == 0) return cudaSuccess;
  cov_pack_gather_kernel<<<(unsigned)((v.n_tiles * 32 + 255) / 256), 256, 0, st>>>(v, slot, desc, raw);
  return cudaGetLastError();
}


// per (bundle, lane): get_cov(s, 0, len-2) — the summed clamped coverage of the whole bundle
__global__ void cov_totals_kernel(DevBatchView v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_bundles * 3) return;
  const int64_t b = i / 3;
  const int s = (int)(i % 3);
  const float* base = v.bpcov + (size_t)s * v.cov_pitch + v.b_gbase[b];
  v.b_cov_total[i] = dev_get_cov(base, 0u, v.b_len[b] - 2u);
}

cudaError_t launch_cov_totals(const DevBatchView& v, cudaStream_t st) {
  if (v.n_bundles == 0) return cudaSuccess;
  cov_totals_kernel<<<(unsigned)((v.n_bundles * 3 + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}
