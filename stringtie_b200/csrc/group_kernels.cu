// group_kernels.cu — sm_100a kernels of SURVEY.md §8 rows a8 + a9 (read loop of build_graphs, groups, colours,
// bundles / bundlenodes).  The per-bundle algorithm is in groups_device.cuh; this file maps it onto the GPU:
//   group_caps_kernel    one thread per read: how many readgroup / group / colour slots the read can need at most
//   groups_walk_kernel   one CTA per bundle, bundles in order of decreasing read count: 128 reads classified per step
//   groups_colour_kernel one warp per bundle: the lanes compact the bundle's groups and colours into the dense
//                        arrays, lane 0 runs the colour pass and builds bundles / bundlenodes
//   groups_rg_kernel     one thread per read: compacts readgroup
//   groups_junc_kernel   one warp per bundle: the junction table in the order the reference re-sorts it into
// The reference's loop is sequential (its state flows both ways between repair, grouping and un-pairing); the walk
// kernel runs the reads that change no structure in parallel and only the others one by one (groups_device.cuh).
#include "groups_device.cuh"
#include "ordered_sum.cuh"

namespace {

__device__ __forceinline__ uint32_t bundle_of_read(const DevBatchView& v, uint32_t n) {
  uint32_t b = v.cta_bundle[n / STG_READ_BLOCK];
  while (v.b_read_off[b + 1] <= n) b++;
  return b;
}

// Upper bounds (groups_device.cuh): a call of merge_read_to_group for read n visits the segments of n and, when the
// mate continues the fragment, those of the mate; each visit creates at most one group and one readgroup entry, and
// there is one call per pair link plus one for the unpaired share.  Colours: at most 3 more per read.
__global__ void __launch_bounds__(STG_READ_BLOCK) group_caps_kernel(DevBatchView v, uint32_t* rcap, unsigned long long* total,
                                                                  int32_t* rj) {
  const int64_t n = (int64_t)blockIdx.x * STG_READ_BLOCK + threadIdx.x;
  unsigned long long cap = 0;
  if (n < v.n_reads) {
    const uint32_t b = bundle_of_read(v, (uint32_t)n);
    const uint32_t r0 = v.b_read_off[b], r1 = v.b_read_off[b + 1];
    const uint32_t s0 = v.r_seg_off[n], ns = v.r_seg_off[n + 1] - s0;
    const uint32_t p0 = v.r_pair_off[n], np = v.r_pair_off[n + 1] - p0;
    cap = (unsigned long long)(2 * np + 1) * ns;
    for (uint32_t p = 0; p < np; p++) {
      const int32_t m = v.pair_idx[p0 + p];
      if (m >= 0 && r0 + (uint32_t)m < r1) cap += v.r_seg_off[r0 + m + 1] - v.r_seg_off[r0 + m];
    }
    rcap[n] = (uint32_t)cap;
    // junction rows of the read: "no row" (-1) reads as the sentinel row 0, as in the reference's juncs[] of a read
    // whose junction was not found (rlink.cpp keeps a pointer; the flattened batch keeps -1)
    for (uint32_t i = 0; i + 1 < ns; i++) {
      const uint32_t k = s0 - (uint32_t)n + i;
      const int32_t j = v.rjunc_idx[k];
      rj[k] = j < 0 ? 0 : j;
    }
  } else if (n == v.n_reads) rcap[n] = 0;
  // block sum -> one atomic
  __shared__ unsigned long long sh[STG_READ_BLOCK / 32];
  for (int o = 16; o; o >>= 1) cap += __shfl_down_sync(0xffffffffu, cap, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cap;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int i = 0; i < STG_READ_BLOCK / 32; i++) s += sh[i];
    if (s) atomicAdd(total, s);
  }
}

// The group arrays of a bundle live in SHARED memory while they fit (a bundle has ~10 groups on average, a few hundred
// when it is deep): every read touches a handful of group fields read-modify-write, and from global memory each of
// those is an L2 round trip on the critical path.  A bundle that outgrows GROUPS_SMEM slots is moved to its global
// arrays and carries on there.
#define GROUPS_SMEM 256
// One CTA walks one bundle; NW warps, of which NW - 1 classify reads (warp 0 adds the float sums of the previous run).
// Deep bundles (>= GROUPS_DEEP_READS reads) get 4 warps: their runs of simple reads fill the widest window.  The rest get
// 2 (one classifying warp): their runs are short and what counts is how many bundles an SM holds.  The two grids run
// side by side on two streams; both kernels ask for the same shared-memory carve-out, otherwise an SM would have to
// drain before it can take a CTA of the other kernel.
#define GROUPS_DEEP_READS 16384
struct GroupSmem {
  uint32_t start[GROUPS_SMEM], end[GROUPS_SMEM];
  int32_t color[GROUPS_SMEM], next[GROUPS_SMEM], merge[GROUPS_SMEM], grid[GROUPS_SMEM];
  float cov[GROUPS_SMEM], multi[GROUPS_SMEM];
  int8_t alive[GROUPS_SMEM];
};

struct GroupGlobal { uint32_t *start, *end; int32_t *color, *next, *merge, *grid; float *cov, *multi; int8_t* alive; int cap; };

__device__ __forceinline__ void groups_to_global(const Grp& S, const GroupGlobal& G, int ng, int from, int step) {
  for (int i = from; i < ng; i += step) {
    G.start[i] = S.g_start[i]; G.end[i] = S.g_end[i]; G.color[i] = S.g_color[i]; G.next[i] = S.g_next[i];
    G.merge[i] = S.g_merge[i]; G.grid[i] = S.g_grid[i]; G.cov[i] = S.g_cov[i]; G.multi[i] = S.g_multi[i]; G.alive[i] = S.g_alive[i];
  }
}

// what thread 0 (the owner of the walk's mutable scalars) shares with the CTA between steps
template <int GROUPS_WARPS>
struct WalkCtl {
  static constexpr int GROUPS_CLS = 32 * GROUPS_WARPS - 32;
  int32_t curr[3];
  int32_t ng, color, err;
  int32_t last[3];             // last read of the run that set the cursor of strand lane s (atomicMax of the thread index)
  uint32_t ok_mask[GROUPS_WARPS], mint_mask[GROUPS_WARPS], wait_mask[GROUPS_WARPS], grow_mask[GROUPS_WARPS];   // [0] unused
  int32_t npc[GROUPS_CLS];
  int32_t n_grow[2];            // window positions of the step's reads that grow a group (GROW kernels; double-buffered)
  int32_t grow[2][GROUPS_CLS];
};

// The CTA's schedule (groups_device.cuh, "the parallel form of the loop"): warps 1.. classify the next GROUPS_CLS reads
// against the state as it stands and commit the order-free part of the leading run of simple ones, each thread for
// its own read; MEANWHILE warp 0 adds the float sums of the PREVIOUS run (one lane per group residue, so that every
// group's sum is added in read order) and the bundle's fragment sums — classification never reads those sums, so
// the two overlap and only meet at the barrier that ends a step.  A read that is not simple is walked the reference's
// way by thread 0 once the pending sums are in.
// GROW: reads that grow a group stay inside the runs (they record the bounds they depend on; pays off where runs are
// short, i.e. in shallow bundles); otherwise such a read ends its run.
template <int GROUPS_WARPS, int MINB, bool GROW>
__global__ void __launch_bounds__(32 * GROUPS_WARPS, MINB) groups_walk_kernel(DevBatchView v, GroupsView g, JuncParams prm,
                                                                             uint32_t order_base) {
  constexpr int GROUPS_STEP = 32 * GROUPS_WARPS, GROUPS_CLS = GROUPS_STEP - 32;
  __shared__ GroupSmem M;
  extern __shared__ __align__(16) unsigned char groups_dyn_smem[];
  auto Rb = reinterpret_cast<FastRec (*)[GROUPS_CLS]>(groups_dyn_smem);   // [2][GROUPS_CLS] records of the step being classified / of the run whose sums are pending
  __shared__ FastDeps Db[GROW ? GROUPS_CLS : 1];
  __shared__ WalkCtl<GROUPS_WARPS> ctl;
  const uint32_t b = g.order[order_base + blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t r0 = v.b_read_off[b], nr = v.b_read_off[b + 1] - r0;
  const uint32_t j0 = v.b_junc_off[b], nj0 = v.b_junc_off[b + 1] - j0;
  if (nr == 0) return;
  Grp S;
  S.nr = (int)nr;
  S.seg0 = v.r_seg_off[r0]; S.pair0 = v.r_pair_off[r0];
  const uint32_t i0 = S.seg0 - r0;                                         // intron base of the bundle
  const uint32_t nintr = (v.r_seg_off[r0 + nr] - S.seg0) - nr;
  S.r_start = v.r_start + r0; S.r_end = v.r_end + r0; S.r_len = v.r_len + r0; S.seg_off = v.r_seg_off + r0;
  S.pair_off = v.r_pair_off + r0; S.r_count = v.r_count + r0; S.pair_cnt = v.pair_cnt + S.pair0; S.r_nh = v.r_nh + r0;
  S.r_flags = v.r_flags + r0;
  S.strand = g.strand + r0; S.seg_s = g.seg_start + S.seg0; S.seg_e = g.seg_end + S.seg0; S.pair = g.pair + S.pair0;
  S.rj = g.rj + i0;
  S.j_start = v.j_start + j0; S.j_end = v.j_end + j0; S.j_nreads = v.jp_nreads + j0; S.j_good = v.jp_nreads_good + j0;
  S.j_ej = v.jp_ej + j0; S.jp_good = v.jp_good + j0; S.jp_good_strand = v.jp_good_strand + j0;
  S.jt_strand = g.jt_strand + j0; S.jt_mm = g.jt_mm + j0;
  S.ap_start = g.ap_start + i0; S.ap_end = g.ap_end + i0; S.ap_strand = g.ap_strand + i0; S.ap_mm = g.ap_mm + i0;
  S.nj0 = (int)nj0; S.nJ = (int)nj0; S.apcap = (int)nintr;
  const size_t G0 = g.rcap_off[r0];
  const uint32_t gcap = g.rcap_off[r0 + nr] - g.rcap_off[r0];
  const GroupGlobal GG{g.g_start + G0, g.g_end + G0, g.g_color + G0, g.g_next + G0, g.g_merge + G0, g.g_grid + G0, g.g_cov + G0,
                       g.g_multi + G0, g.g_alive + G0, (int)gcap};
  S.g_start = M.start; S.g_end = M.end; S.g_color = M.color; S.g_next = M.next; S.g_merge = M.merge; S.g_grid = M.grid;
  S.g_cov = M.cov; S.g_multi = M.multi; S.g_alive = M.alive;
  S.ng = 0; S.gcap = GROUPS_SMEM;
  S.eqcol = g.eqcol + G0 + 3 * (size_t)r0; S.ncol = 0; S.ccap = (int)(gcap + 3 * nr);
  S.rg = g.rg; S.rg_cnt = g.rg_cnt + r0; S.rcap_off = g.rcap_off + r0;
  S.curr.fill(-1); S.startg.fill(-1);
  S.junctionsupport = prm.junctionsupport; S.bundledist = g.bundledist;
  S.refstart = v.b_start[b]; S.len = v.b_len[b];
  S.c1 = v.bpcov + v.cov_pitch + v.b_gbase[b];
  S.prm = prm; S.err = 0;
  if (gcap == 0 || gcap > 0x7fffffffu) { if (!tid) atomicOr(v.err_flag, STG_GERR_GROUPS); return; }
  bool in_smem = true;
  int color = 0;                 // uniform (every thread follows the colour counter)
  int ng_u = 0;                  // uniform copy of thread 0's S.ng (changes only when a read is walked)
  double nf = 0, fl = 0;         // thread 0
  int n = 0, slow_len = 0, slow_budget = 0;   // uniform
  int width = 32;                             // uniform: reads classified in the next step (follows the length of the runs)
  int buf = 0, pend_k = 0, pend_buf = 0;      // uniform: record buffer of this step; run whose float sums are pending
  const int inr = (int)nr;
  const unsigned FULL = 0xffffffffu;
  const int ct = tid - 32;                    // index among the classifying threads (warp 0: negative)
  if (tid == 0) { ctl.err = 0; ctl.last[0] = ctl.last[1] = ctl.last[2] = -1; ctl.n_grow[0] = ctl.n_grow[1] = 0; }
  __syncthreads();

  // float sums of a committed run, in read order: lane l owns the groups with slot % 32 == l; the sum of the group being
  // worked on is kept in registers.  Thread 0 also adds the bundle's fragment sums (a read that adds nothing holds 0).
  auto pending_sums = [&](const FastRec* R, int k) {
    int cur = -1;
    float acc_cov = 0.f, acc_multi = 0.f;
    for (int j = 0; j < k; j++) {
      const int nv = R[j].nv, hm = R[j].has_multi;
      for (int x = 0; x < nv; x++) {
        const int gr = R[j].grp[x];
        if ((gr & 31) == lane) {
          if (gr != cur) {
            if (cur >= 0) { S.g_cov[cur] = acc_cov; S.g_multi[cur] = acc_multi; }
            cur = gr; acc_cov = S.g_cov[gr]; acc_multi = S.g_multi[gr];
          }
          if (hm >> x & 1) acc_multi += R[j].multi[x];
          acc_cov += R[j].cov[x];
        }
      }
    }
    if (cur >= 0) { S.g_cov[cur] = acc_cov; S.g_multi[cur] = acc_multi; }
    if (lane == 0) {
      double a = fl, c = nf;
#pragma unroll 8
      for (int j = 0; j < k; j++) { a += R[j].frag_len; c += R[j].numfrag; }
      fl = a; nf = c;
    }
  };

  while (true) {
    const bool classify = n < inr && slow_budget == 0;
    FastRec* R = Rb[buf];
    const int hi = n + width < inr ? n + width : inr, me = n + ct;
    int st = 0;                                 // 1 simple, 3 simple + grows a group, 0 not simple, 2 waits
    if (warp == 0) { if (pend_k) pending_sums(Rb[pend_buf], pend_k); }
    else if (classify) {
      if (me < hi) st = GROW ? classify_read<true>(S, me, n, hi, R[ct], &Db[GROW ? ct : 0]) : classify_read<false>(S, me, n, hi, R[ct]);
      ctl.npc[ct] = (st == 1 || (GROW && st == 3)) ? R[ct].npc : -1;
      if (GROW && st == 3) ctl.grow[buf][atomicAdd(&ctl.n_grow[buf], 1)] = ct;
    }
    pend_k = 0;
    __syncthreads();   // (B) the pending sums are in; the records of this step are written
    if (n >= inr) break;
    if (tid == 32) {   // everybody read them before (B); the cursors are set again after (C), the list is filled next step
      ctl.last[0] = ctl.last[1] = ctl.last[2] = -1;
      ctl.n_grow[buf ^ 1] = 0;
    }
    int k = 0;
    bool no_walk = false;
    if (classify) {
      if (warp != 0) {
        // reads that continue into the SAME mate see each other's readgroup pushes: only the first one stays simple, the
        // others wait for the next step
        if ((st == 1 || st == 3) && R[ct].npc >= 0) {
          const int mine = R[ct].npc;
          for (int t = 0; t < ct; t++)
            if (ctl.npc[t] == mine) { st = 2; break; }
        }
        if (GROW && (st == 1 || st == 3) && Db[GROW ? ct : 0].n) {
          // a read that took for granted a bound which an earlier read of the step moves waits for the next step
          const int ngr = ctl.n_grow[buf];
          for (int q = 0; q < ngr; q++) {
            const int t = ctl.grow[buf][q];
            if (t < ct && grows_past(Db[GROW ? ct : 0], R[t])) { st = 2; break; }
          }
        }
        const unsigned okm = __ballot_sync(FULL, st == 1 || (GROW && st == 3)), mim = __ballot_sync(FULL, (st == 1 || st == 3) && R[ct].newcol);
        const unsigned wam = __ballot_sync(FULL, st == 2), grm = __ballot_sync(FULL, !GROW && st == 3);
        if (lane == 0) { ctl.ok_mask[warp] = okm; ctl.mint_mask[warp] = mim; ctl.wait_mask[warp] = wam; ctl.grow_mask[warp] = grm; }
      }
      __syncthreads();   // (C)
      // k = length of the leading run of simple reads, plus the read behind it if all it does beyond that is to grow a
      // group; fresh colours are numbered in read order
      {
        bool open = true;
        for (int w2 = 1; w2 < GROUPS_WARPS; w2++) {
          const unsigned m = ctl.ok_mask[w2];
          const int kw = !open ? 0 : (m == FULL ? 32 : __ffs(~m) - 1);
          k += kw;
          if (kw < 32) open = false;
        }
      }
      const bool grows = k < GROUPS_CLS && (ctl.grow_mask[1 + (k >> 5)] >> (k & 31) & 1u);
      const bool waits0 = k < GROUPS_CLS && (ctl.wait_mask[1 + (k >> 5)] >> (k & 31) & 1u);
      if (grows) k++;
      int before = 0;     // colours minted by the run's reads ahead of this thread
      int minted = 0;
      for (int w2 = 1; w2 < GROUPS_WARPS; w2++) {
        const int rem = k - 32 * (w2 - 1);
        const int kw = rem >= 32 ? 32 : (rem > 0 ? rem : 0);
        const unsigned run = kw == 32 ? FULL : (1u << kw) - 1u;
        const unsigned mi = ctl.mint_mask[w2] & run;
        if (w2 < warp) before += __popc(mi);
        else if (w2 == warp) before += __popc(mi & ((1u << lane) - 1u));
        minted += __popc(mi);
      }
      if (warp != 0 && ct < k) {
        commit_free(S, R[ct], me, color + before);
        if (S.err) atomicOr(&ctl.err, (int)S.err);
        if (R[ct].sno >= 0) atomicMax(&ctl.last[R[ct].sno], ct);
      }
      __syncthreads();   // (D) colours, readgroup entries and cursors of the run are in place
      if (ctl.err) break;
      // lane cursors: the last read of the run that moved the cursor of a strand lane decides it
      {
        const int l0 = ctl.last[0], l1 = ctl.last[1], l2 = ctl.last[2];
        if (l0 >= 0) S.curr.a = R[l0].curr;
        if (l1 >= 0) S.curr.b = R[l1].curr;
        if (l2 >= 0) S.curr.c = R[l2].curr;
      }
      color += minted;
      if (tid == 0) S.ncol += minted;
      pend_k = k; pend_buf = buf; buf ^= 1;
      // the run reached the end of the window, ended with a read that grows a group, or stopped at a read that waits for
      // its mate: no sequential read in this step
      no_walk = k >= hi - n || grows || waits0;
      // short runs: classify fewer reads next time (fewer divergent lanes make a cheaper step); full runs: widen again
      if (k >= width) width = width * 2 < GROUPS_CLS ? width * 2 : GROUPS_CLS;
      else { const int w2 = 2 * k + 8; width = w2 < 16 ? 16 : (w2 < GROUPS_CLS ? w2 : GROUPS_CLS); }
      n += k;
      if (k == 0) { slow_len = slow_len * 2 + 1 < 15 ? slow_len * 2 + 1 : 15; slow_budget = slow_len; }
      else if (k >= 8) slow_len = 0;
    } else slow_budget--;
    if (n < inr && !no_walk) {
      // a read for thread 0 to walk: it works on the float sums (they have to be in) and may create groups (leave
      // shared memory before it can overflow)
      if (warp == 0 && pend_k) pending_sums(Rb[pend_buf], pend_k);
      pend_k = 0;
      __syncthreads();   // (E1) sums in; everybody has read ctl.last
      if (in_smem && ng_u + (int)(S.rcap_off[n + 1] - S.rcap_off[n]) > GROUPS_SMEM) {
        groups_to_global(S, GG, ng_u, tid, GROUPS_STEP);
        S.g_start = GG.start; S.g_end = GG.end; S.g_color = GG.color; S.g_next = GG.next; S.g_merge = GG.merge; S.g_grid = GG.grid;
        S.g_cov = GG.cov; S.g_multi = GG.multi; S.g_alive = GG.alive; S.gcap = GG.cap;
        in_smem = false;
        __syncthreads();
      }
      if (tid == 0) {
        S.err |= (uint32_t)ctl.err;
        if (!S.err) color = walk_read(S, n, color, nf, fl);
        ctl.curr[0] = S.curr.a; ctl.curr[1] = S.curr.b; ctl.curr[2] = S.curr.c;
        ctl.ng = S.ng; ctl.color = color; ctl.err = (int)S.err;
      }
      __syncthreads();   // (E2)
      S.curr.a = ctl.curr[0]; S.curr.b = ctl.curr[1]; S.curr.c = ctl.curr[2];
      ng_u = ctl.ng; color = ctl.color;
      if (ctl.err) break;
      n++;
    }
  }
  __syncthreads();
  if (warp == 0 && pend_k) pending_sums(Rb[pend_buf], pend_k);
  __syncthreads();
  if (tid == 0) {
    if (ctl.err) S.err |= (uint32_t)ctl.err;
    if (!S.err) walk_finish(S, g.jt_perm + j0 + i0, g.jt_inv + j0 + i0);
    if (S.err) atomicOr(v.err_flag, S.err);
    else {
      g.b_ng[b] = (uint32_t)S.ng; g.b_ncol[b] = (uint32_t)S.ncol; g.b_napp[b] = (uint32_t)(S.nJ - S.nj0);
      g.b_numfrag[b] = nf; g.b_fraglen[b] = fl;
      g.b_startgroup[3 * b] = S.startg.a; g.b_startgroup[3 * b + 1] = S.startg.b; g.b_startgroup[3 * b + 2] = S.startg.c;
    }
    ctl.ng = S.err ? 0 : S.ng;
  }
  __syncthreads();
  if (in_smem) groups_to_global(S, GG, ctl.ng, tid, GROUPS_STEP);   // the CTA writes the groups out of shared memory
}

// ---- oversized bundles: one thread-block CLUSTER per bundle ------------------------------------
// A bundle of 10^6 reads walked by one CTA is bound by the latency of a step (classification = chains of dependent L2
// loads, ~40 us whatever the width of the window).  Nearly every read of such a bundle is simple, so the window can be as
// wide as the hardware allows: CL CTAs of 256 threads on CL SMs of one GPC classify CL * 256 reads per step against the
// same state; the schedule of a step is the one of groups_walk_kernel (groups_device.cuh, groups_first_half_steps): reads
// that share a continuation mate yield to the first, the leading run of simple reads commits its order-free part in
// parallel, its ordered part goes to a log, the first read that is not simple is walked by one thread
// after the logged sums are in.  Cross-CTA state lives in global memory (group arrays, records, masks); the CTAs meet at
// cluster barriers (barrier.cluster, which orders global memory at cluster scope).
__device__ int g_groups_debug = 0;
#define CLUSTER_CTAS 8
#define CLUSTER_THREADS 256
#define CLUSTER_WINDOW (CLUSTER_CTAS * CLUSTER_THREADS)
#define CLUSTER_HOPS 24           // links of the group chain a read of the window may follow (classify_read)
struct ClusterCtl {
  int32_t err, ng, color;
  int32_t curr[3];
  int32_t last[2][3];
  int32_t n_grow[3];              // reads of the step that grow a group (three buffers: see the reset in the kernel)
  uint32_t ok_mask[CLUSTER_WINDOW / 32], mint_mask[CLUSTER_WINDOW / 32], wait_mask[CLUSTER_WINDOW / 32];
  int32_t grow[3][CLUSTER_WINDOW];
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// one oversized bundle (order[order_base + cid]) walked by the calling cluster
__device__ void cluster_walk_bundle(const DevBatchView& v, const GroupsView& g, const JuncParams& prm, uint32_t order_base, uint32_t cid) {
  const uint32_t crank = cluster_cta_rank();
  const uint32_t b = g.order[order_base + cid];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gt = (int)crank * CLUSTER_THREADS + tid;              // slot in the window
  const int gw = gt >> 5;
  const uint32_t r0 = v.b_read_off[b], nr = v.b_read_off[b + 1] - r0;
  const uint32_t j0 = v.b_junc_off[b], nj0 = v.b_junc_off[b + 1] - j0;
  if (nr == 0) return;                                            // (uniform over the cluster)
  Grp S;
  S.nr = (int)nr;
  S.seg0 = v.r_seg_off[r0]; S.pair0 = v.r_pair_off[r0];
  const uint32_t i0 = S.seg0 - r0;
  const uint32_t nintr = (v.r_seg_off[r0 + nr] - S.seg0) - nr;
  S.r_start = v.r_start + r0; S.r_end = v.r_end + r0; S.r_len = v.r_len + r0; S.seg_off = v.r_seg_off + r0;
  S.pair_off = v.r_pair_off + r0; S.r_count = v.r_count + r0; S.pair_cnt = v.pair_cnt + S.pair0; S.r_nh = v.r_nh + r0;
  S.r_flags = v.r_flags + r0;
  S.strand = g.strand + r0; S.seg_s = g.seg_start + S.seg0; S.seg_e = g.seg_end + S.seg0; S.pair = g.pair + S.pair0;
  S.rj = g.rj + i0;
  S.j_start = v.j_start + j0; S.j_end = v.j_end + j0; S.j_nreads = v.jp_nreads + j0; S.j_good = v.jp_nreads_good + j0;
  S.j_ej = v.jp_ej + j0; S.jp_good = v.jp_good + j0; S.jp_good_strand = v.jp_good_strand + j0;
  S.jt_strand = g.jt_strand + j0; S.jt_mm = g.jt_mm + j0;
  S.ap_start = g.ap_start + i0; S.ap_end = g.ap_end + i0; S.ap_strand = g.ap_strand + i0; S.ap_mm = g.ap_mm + i0;
  S.nj0 = (int)nj0; S.nJ = (int)nj0; S.apcap = (int)nintr;
  const size_t G0 = g.rcap_off[r0];
  const uint32_t gcap = g.rcap_off[r0 + nr] - g.rcap_off[r0];
  // the group arrays stay in global memory: every CTA of the cluster reads them
  S.g_start = g.g_start + G0; S.g_end = g.g_end + G0; S.g_color = g.g_color + G0; S.g_next = g.g_next + G0; S.g_merge = g.g_merge + G0;
  S.g_grid = g.g_grid + G0; S.g_cov = g.g_cov + G0; S.g_multi = g.g_multi + G0; S.g_alive = g.g_alive + G0;
  S.ng = 0; S.gcap = (int)gcap;
  S.eqcol = g.eqcol + G0 + 3 * (size_t)r0; S.ncol = 0; S.ccap = (int)(gcap + 3 * nr);
  S.rg = g.rg; S.rg_cnt = g.rg_cnt + r0; S.rcap_off = g.rcap_off + r0;
  S.curr.fill(-1); S.startg.fill(-1);
  S.junctionsupport = prm.junctionsupport; S.bundledist = g.bundledist;
  S.refstart = v.b_start[b]; S.len = v.b_len[b];
  S.c1 = v.bpcov + v.cov_pitch + v.b_gbase[b];
  S.prm = prm; S.err = 0;
  ClusterCtl* ctl = (ClusterCtl*)g.cl_ctl + cid;
  FastRec* R = g.cl_rec + (size_t)cid * CLUSTER_WINDOW;
  const size_t vbase = (size_t)g.vlog_off[order_base + cid];
  int32_t* claim = g.cl_claim + vbase;                            // [nr] first simple read of the step that continues into read i
  const bool boss = gt == 0;
  if (gcap == 0 || gcap > 0x7fffffffu) { if (boss) atomicOr(v.err_flag, STG_GERR_GROUPS); return; }
  __shared__ OrderedSumSmem osm;   // scratch of the flush
  __shared__ double red_a[CLUSTER_THREADS / 32], red_c[CLUSTER_THREADS / 32];
  __shared__ int red_ea[CLUSTER_THREADS / 32], red_ec[CLUSTER_THREADS / 32];
  __shared__ int32_t gl_t[CLUSTER_THREADS], gl_g[CLUSTER_THREADS];   // a chunk of the step's growing reads: slot, group,
  __shared__ uint32_t gl_s[CLUSTER_THREADS], gl_e[CLUSTER_THREADS];  // ... new bounds
  int emin_a = 4096, emin_c = 4096;   // the boss: smallest unit of the fragment-sum terms so far
  int color = 0;
  double nf = 0, fl = 0;         // the boss
  // a step costs nearly the same whatever the width of the window (the latency of the classification's load chains and of
  // the cluster barriers), so the window is wide: four times the recent run length, at least 512 reads (reads classified
  // behind the stop only add L2 traffic and a longer wait for the slowest thread)
  int width = CLUSTER_WINDOW, run_avg = CLUSTER_WINDOW;
  int n = 0, slow_len = 0, slow_budget = 0, pb = 0, gb = 0, flushed = 0;
  const int inr = (int)nr;
  const unsigned FULL = 0xffffffffu;
  if (boss) { ctl->err = 0; for (int x = 0; x < 3; x++) { ctl->last[0][x] = -1; ctl->last[1][x] = -1; ctl->n_grow[x] = 0; } }
  cluster_sync_all();

  auto flush = [&](int upto) {    // the logged sums of the reads [flushed, upto): CTA 0 adds them (ordered_sum.cuh)
    if (upto > flushed && crank == 0) {
      const uint32_t n6 = (uint32_t)(upto - flushed) * STG_FP_MAXV;
      const int32_t* lg = g.vlog_grp + (vbase + (size_t)flushed) * STG_FP_MAXV;
      const float2* lv = g.vlog_val + (vbase + (size_t)flushed) * STG_FP_MAXV;
      const bool sorted_ok = cta_ordered_keyed_sum2(n6, [&](uint32_t i) { return lg[i]; }, [&](uint32_t i) { return lv[i]; }, S.g_cov, S.g_multi, osm,
                                                    g.vlog_sorted + (vbase + (size_t)flushed) * STG_FP_MAXV);
      if (!sorted_ok && warp == 0) {   // too many distinct groups for the shared-memory counts: lane l owns the groups with slot % 32 == l
        int cur = -1;
        float acc_cov = 0.f, acc_multi = 0.f;
        for (uint32_t i = 0; i < n6; i++) {
          const int gr = lg[i];
          if (gr >= 0 && (gr & 31) == lane) {
            if (gr != cur) {
              if (cur >= 0) { S.g_cov[cur] = acc_cov; S.g_multi[cur] = acc_multi; }
              cur = gr; acc_cov = S.g_cov[gr]; acc_multi = S.g_multi[gr];
            }
            acc_multi += lv[i].y; acc_cov += lv[i].x;
          }
        }
        if (cur >= 0) { S.g_cov[cur] = acc_cov; S.g_multi[cur] = acc_multi; }
      }
      // the bundle's fragment sums (f64 sums of values that are floats): order-free while every partial sum is exactly
      // representable — all terms are >= 0 and multiples of 2^emin, the total stays below 2^(emin + 53) — which is checked;
      // otherwise the boss adds this range one read after the other, as the reference does
      {
        const double* pf = g.vlog_fl + vbase;
        const double* pn = g.vlog_nf + vbase;
        double pa = 0, pc = 0;
        int ea = 4096, ec = 4096;
        for (int r = flushed + tid; r < upto; r += CLUSTER_THREADS) {
          const double a = pf[r], c = pn[r];
          pa += a; pc += c;
          if (a > 0) { const int e = ((__double2hiint(a) >> 20) & 0x7ff) - 1023 - 23; ea = e < ea ? e : ea; }
          if (c > 0) { const int e = ((__double2hiint(c) >> 20) & 0x7ff) - 1023 - 23; ec = e < ec ? e : ec; }
        }
        for (int o = 16; o; o >>= 1) {
          pa += __shfl_down_sync(FULL, pa, o); pc += __shfl_down_sync(FULL, pc, o);
          const int xa = __shfl_down_sync(FULL, ea, o), xc = __shfl_down_sync(FULL, ec, o);
          ea = xa < ea ? xa : ea; ec = xc < ec ? xc : ec;
        }
        if (lane == 0) { red_a[warp] = pa; red_c[warp] = pc; red_ea[warp] = ea; red_ec[warp] = ec; }
        __syncthreads();
        if (boss) {
          double ta = 0, tc = 0;
          for (int x = 0; x < CLUSTER_THREADS / 32; x++) {
            ta += red_a[x]; tc += red_c[x];
            emin_a = red_ea[x] < emin_a ? red_ea[x] : emin_a; emin_c = red_ec[x] < emin_c ? red_ec[x] : emin_c;
          }
          const double na = fl + ta, nc = nf + tc;
          const int xa = ((__double2hiint(na) >> 20) & 0x7ff) - 1023, xc = ((__double2hiint(nc) >> 20) & 0x7ff) - 1023;
          if ((na == 0 || xa + 1 - emin_a <= 52) && (nc == 0 || xc + 1 - emin_c <= 52)) { fl = na; nf = nc; }
          else {
            double a = fl, c = nf;
#pragma unroll 4
            for (int r = flushed; r < upto; r++) { a += pf[r]; c += pn[r]; }
            fl = a; nf = c;
          }
        }
        __syncthreads();
      }
    }
    if (upto > flushed) flushed = upto;
  };

  long long t_cls = 0, t_commit = 0, t_flush = 0, t_walk = 0, t_bar = 0, steps = 0, walks = 0, t0c = 0;
  while (true) {
    const bool classify = n < inr && slow_budget == 0;
    const int hi = n + width < inr ? n + width : inr, me = n + gt;
    int st = 0, npc = -1;
    if (boss) t0c = clock64();
    FastDeps D;
    D.n = 0;
    if (classify) {
      if (me < hi) st = classify_read<true>(S, me, n, hi, R[gt], &D, CLUSTER_HOPS);
      if (st == 1 || st == 3) npc = R[gt].npc;
      if (npc >= 0) atomicMin(&claim[npc], gt);
      if (st == 3) ctl->grow[gb][atomicAdd(&ctl->n_grow[gb], 1)] = gt;
    }
    if (boss) { const long long t = clock64(); t_cls += t - t0c; t0c = t; }
    cluster_sync_all();        // (1) records and claims of this step are written
    if (boss) { const long long t = clock64(); t_bar += t - t0c; t0c = t; steps++; }
    if (n >= inr) break;
    int k = 0;
    bool no_walk = false;
    if (classify) {
      // reads that continue into the SAME mate see each other's readgroup pushes: only the first simple one stays
      if (npc >= 0 && claim[npc] < gt) st = 2;
      // reads that grow a group stay inside the run (groups_device.cuh, FastDeps): a read that took for granted a bound
      // which an earlier read of the step moves waits for the next step.  The counter used two steps from now is cleared
      // here: a cluster barrier lies between this and both its last readers and its next writers.
      {
        const int ngr = ctl->n_grow[gb];
        if (boss) ctl->n_grow[(gb + 2) % 3] = 0;
        for (int base = 0; base < ngr; base += CLUSTER_THREADS) {
          if (base + tid < ngr) {
            const int t = ctl->grow[gb][base + tid];
            gl_t[tid] = t; gl_g[tid] = R[t].xg; gl_s[tid] = R[t].xs; gl_e[tid] = R[t].xe;
          }
          __syncthreads();
          if ((st == 1 || st == 3) && D.n) {
            const int m = ngr - base < CLUSTER_THREADS ? ngr - base : CLUSTER_THREADS;
            for (int q = 0; q < m && st != 2; q++) {
              if (gl_t[q] >= gt) continue;
              for (int i = 0; i < D.n; i++)
                if (D.g[i] == gl_g[q]) {
                  if (D.kind >> i & 1u) { if (gl_s[q] <= D.x[i]) st = 2; }
                  else if (gl_e[q] >= D.x[i]) st = 2;
                }
            }
          }
          __syncthreads();
        }
        gb = (gb + 1) % 3;
      }
      const unsigned okm = __ballot_sync(FULL, st == 1 || st == 3), mim = __ballot_sync(FULL, (st == 1 || st == 3) && R[gt].newcol);
      const unsigned wam = __ballot_sync(FULL, st == 2);
      if (lane == 0) { ctl->ok_mask[gw] = okm; ctl->mint_mask[gw] = mim; ctl->wait_mask[gw] = wam; }
      cluster_sync_all();      // (2) masks are in; every claim has been looked at
      if (npc >= 0) claim[npc] = 0x7f7f7f7f;
      // k = length of the leading run of simple reads: lane l looks at mask words l and 32 + l (one round trip to L2 for the
      // whole warp instead of a dependent load per word)
      const unsigned okA = ctl->ok_mask[lane], okB = ctl->ok_mask[32 + lane];
      const unsigned miA = ctl->mint_mask[lane], miB = ctl->mint_mask[32 + lane];
      static_assert(CLUSTER_WINDOW == 2048, "two mask words per lane");
      {
        const unsigned notfullA = __ballot_sync(FULL, okA != FULL), notfullB = __ballot_sync(FULL, okB != FULL);
        int wfirst;                                  // first word with a read that is not simple (64: none)
        unsigned mfirst = FULL;
        if (notfullA) { wfirst = __ffs(notfullA) - 1; mfirst = __shfl_sync(FULL, okA, wfirst); }
        else if (notfullB) { wfirst = __ffs(notfullB) - 1; mfirst = __shfl_sync(FULL, okB, wfirst); wfirst += 32; }
        else wfirst = 64;
        k = wfirst * 32 + (wfirst < 64 ? __ffs(~mfirst) - 1 : 0);
      }
      const bool waits0 = k < CLUSTER_WINDOW && (ctl->wait_mask[k >> 5] >> (k & 31) & 1u);
      int before = 0, minted = 0;     // colours minted by the run's reads ahead of this thread / by the whole run
      {
        auto run_of = [&](int w2) -> unsigned { const int rem = k - 32 * w2; return rem >= 32 ? FULL : (rem > 0 ? (1u << rem) - 1u : 0u); };
        const unsigned a = miA & run_of(lane), b2 = miB & run_of(32 + lane);
        int mine = 0;                                // colours of the run in the words this lane holds, below word gw
        if (lane < gw) mine += __popc(a);
        if (32 + lane < gw) mine += __popc(b2);
        int tot = __popc(a) + __popc(b2);
        for (int o = 16; o; o >>= 1) { mine += __shfl_xor_sync(FULL, mine, o); tot += __shfl_xor_sync(FULL, tot, o); }
        const unsigned own = __shfl_sync(FULL, gw < 32 ? a : b2, gw & 31);
        before = mine + __popc(own & ((1u << lane) - 1u));
        minted = tot;
      }
      if (gt < k) {
        commit_free(S, R[gt], me, color + before);
        if (S.err) atomicOr(&ctl->err, (int)S.err);
        if (R[gt].sno >= 0) atomicMax(&ctl->last[pb][R[gt].sno], gt);
        const size_t q = vbase + (size_t)me;
        const int nv = R[gt].nv, hm = R[gt].has_multi;
#pragma unroll
        for (int x = 0; x < STG_FP_MAXV; x++) {
          g.vlog_grp[q * STG_FP_MAXV + x] = x < nv ? R[gt].grp[x] : -1;
          g.vlog_val[q * STG_FP_MAXV + x] = x < nv ? make_float2(R[gt].cov[x], (hm >> x & 1) ? R[gt].multi[x] : 0.f) : make_float2(0.f, 0.f);
        }
        g.vlog_fl[q] = R[gt].frag_len; g.vlog_nf[q] = R[gt].numfrag;
      }
      cluster_sync_all();      // (3) colours, readgroup entries, cursors and the log of the run are in place
      if (boss) { const long long t = clock64(); t_commit += t - t0c; t0c = t; }
      if (ctl->err) break;
      {
        const int l0 = ctl->last[pb][0], l1 = ctl->last[pb][1], l2 = ctl->last[pb][2];
        if (l0 >= 0) S.curr.a = R[l0].curr;
        if (l1 >= 0) S.curr.b = R[l1].curr;
        if (l2 >= 0) S.curr.c = R[l2].curr;
      }
      if (boss) { ctl->last[pb ^ 1][0] = -1; ctl->last[pb ^ 1][1] = -1; ctl->last[pb ^ 1][2] = -1; S.ncol += minted; }
      pb ^= 1;
      color += minted;
      no_walk = k >= hi - n || waits0;
      run_avg = (3 * run_avg + (k >= hi - n ? CLUSTER_WINDOW : k)) / 4;
      width = 4 * run_avg < 512 ? 512 : (4 * run_avg < CLUSTER_WINDOW ? 4 * run_avg : CLUSTER_WINDOW);
      n += k;
      if (k == 0) { slow_len = slow_len * 2 + 1 < 15 ? slow_len * 2 + 1 : 15; slow_budget = slow_len; }
      else if (k >= 8) slow_len = 0;
    } else slow_budget--;
    if (n < inr && !no_walk) {
      flush(n);                // CTA 0 (it ends with __syncthreads: the boss, thread 0 of CTA 0, sees the sums)
      if (boss) { const long long t = clock64(); t_flush += t - t0c; t0c = t; walks++; }
      if (boss) {
        S.err |= (uint32_t)ctl->err;
        if (!S.err) color = walk_read(S, n, color, nf, fl);
        ctl->curr[0] = S.curr.a; ctl->curr[1] = S.curr.b; ctl->curr[2] = S.curr.c;
        ctl->ng = S.ng; ctl->color = color; ctl->err = (int)S.err;
      }
      cluster_sync_all();      // (5)
      if (boss) { const long long t = clock64(); t_walk += t - t0c; t0c = t; }
      S.curr.a = ctl->curr[0]; S.curr.b = ctl->curr[1]; S.curr.c = ctl->curr[2];
      color = ctl->color;
      if (ctl->err) break;
      n++;
      flushed = n;
    }
  }
  if (boss) t0c = clock64();
  if (!ctl->err) flush(inr);
  cluster_sync_all();
  if (boss && g_groups_debug)
    printf("cluster bundle %u: reads %u steps %lld walks %lld | Mcycles: classify %.1f barrier1 %.1f commit %.1f flush %.1f walk %.1f final flush %.1f\n", b, nr,
           steps, walks, t_cls * 1e-6, t_bar * 1e-6, t_commit * 1e-6, t_flush * 1e-6, t_walk * 1e-6, (clock64() - t0c) * 1e-6);
  if (boss) {
    if (ctl->err) S.err |= (uint32_t)ctl->err;
    if (!S.err) walk_finish(S, g.jt_perm + j0 + i0, g.jt_inv + j0 + i0);
    if (S.err) atomicOr(v.err_flag, S.err);
    else {
      g.b_ng[b] = (uint32_t)S.ng; g.b_ncol[b] = (uint32_t)S.ncol; g.b_napp[b] = (uint32_t)(S.nJ - S.nj0);
      g.b_numfrag[b] = nf; g.b_fraglen[b] = fl;
      g.b_startgroup[3 * b] = S.startg.a; g.b_startgroup[3 * b + 1] = S.startg.b; g.b_startgroup[3 * b + 2] = S.startg.c;
    }
  }
  __syncwarp();   // the boss's warp leaves together: the caller goes on to an aligned cluster barrier
}

// Persistent clusters: as many as the GPU holds at once (launch_groups_walk), each drawing the next oversized bundle (longest
// first) from a ticket.  A cluster that had to be scheduled in the middle of the stage would wait for 8 SMs of one GPC to be
// free at the same moment, which the small CTAs of the other walk kernels hardly ever allow; a resident one just goes on.
__global__ void __launch_bounds__(CLUSTER_THREADS, 1) groups_walk_cluster_kernel(DevBatchView v, GroupsView g, JuncParams prm,
                                                                                  uint32_t order_base, uint32_t n_cluster) {
  const uint32_t slot = blockIdx.x / CLUSTER_CTAS;                // this cluster
  volatile uint32_t* mine = g.cl_ticket + 1 + slot;
  for (;;) {
    if (cluster_cta_rank() == 0 && threadIdx.x == 0) *mine = atomicAdd(g.cl_ticket, 1u);
    __syncwarp();
    cluster_sync_all();
    const uint32_t cid = *mine;
    cluster_sync_all();          // everyone has it before the slot is written again
    if (cid >= n_cluster) break;
    cluster_walk_bundle(v, g, prm, order_base, cid);
  }
}

// compaction of the colour tables, one thread per colour of the batch: its bundle is found by bisection of the prefix b_ncol
__global__ void __launch_bounds__(256) groups_colcopy_kernel(DevBatchView v, GroupsView g, uint32_t ncol_total) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncol_total) return;
  uint32_t lo = 0, hi = (uint32_t)v.n_bundles;          // b_ncol[lo] <= j < b_ncol[hi]
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (g.b_ncol[mid] <= j) lo = mid; else hi = mid; }
  const uint32_t r0 = v.b_read_off[lo];
  const size_t C0 = (size_t)g.rcap_off[r0] + 3 * (size_t)r0;
  const int32_t e = g.eqcol[C0 + (j - g.b_ncol[lo])];
  g.d_eqcol[j] = e; g.d_eq2[j] = e; g.d_eqneg[j] = -1; g.d_eqpos[j] = -1; g.d_bundlecol[j] = -1;
}

// after the scans: b_ng / b_ncol / b_napp are exclusive prefixes
__global__ void __launch_bounds__(128) groups_colour_kernel(DevBatchView v, GroupsView g) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= v.n_bundles) return;
  const int lane = threadIdx.x & 31;
  const uint32_t b = g.order[w];
  const uint32_t r0 = v.b_read_off[b];
  const size_t G0 = g.rcap_off[r0], C0 = G0 + 3 * (size_t)r0;
  const size_t Gd = g.b_ng[b], Cd = g.b_ncol[b];
  const int ng = (int)(g.b_ng[b + 1] - g.b_ng[b]), nc = (int)(g.b_ncol[b + 1] - g.b_ncol[b]);
  for (int i = lane; i < ng; i += 32) {
    g.d_g_start[Gd + i] = g.g_start[G0 + i]; g.d_g_end[Gd + i] = g.g_end[G0 + i]; g.d_g_color[Gd + i] = g.g_color[G0 + i];
    g.d_g_next[Gd + i] = g.g_next[G0 + i]; g.d_g_merge[Gd + i] = g.g_merge[G0 + i]; g.d_g_cov[Gd + i] = g.g_cov[G0 + i];
    g.d_g_multi[Gd + i] = g.g_multi[G0 + i]; g.d_g_alive[Gd + i] = g.g_alive[G0 + i];
    g.d_color2[Gd + i] = g.g_color[G0 + i]; g.d_negp[Gd + i] = 0.f;
  }
  for (int i = lane; i < 3 * ng; i += 32) g.d_g2b[3 * Gd + i] = -1;
  // (the colour tables were compacted by groups_colcopy_kernel: a deep bundle has 10^6 colours)
  (void)C0; (void)Cd; (void)nc;
  __syncwarp();
  if (lane) return;
  for (int s = 0; s < 3; s++) { g.b_nbun[3 * b + s] = 0; g.b_nbn[3 * b + s] = 0; }
  if (ng == 0) return;
  Grp2 S;
  S.ng = ng; S.ncol = nc; S.bundledist = g.bundledist; S.startg = g.b_startgroup + 3 * (size_t)b;
  S.g_start = g.d_g_start + Gd; S.g_end = g.d_g_end + Gd; S.g_next = g.d_g_next + Gd; S.g_cov = g.d_g_cov + Gd;
  S.g_multi = g.d_g_multi + Gd;
  S.color = g.d_color2 + Gd; S.negp = g.d_negp + Gd;
  S.eq = g.d_eq2 + Cd; S.eqneg = g.d_eqneg + Cd; S.eqpos = g.d_eqpos + Cd; S.bundlecol = g.d_bundlecol + Cd;
  S.g2b = g.d_g2b + 3 * Gd;
  S.bun_len = g.bun_len + 3 * Gd; S.bun_start = g.bun_start + 3 * Gd; S.bun_last = g.bun_last + 3 * Gd;
  S.bun_cov = g.bun_cov + 3 * Gd; S.bun_multi = g.bun_multi + 3 * Gd;
  S.bn_s = g.bn_start + 3 * Gd; S.bn_e = g.bn_end + 3 * Gd; S.bn_cov = g.bn_cov + 3 * Gd; S.bn_next = g.bn_next + 3 * Gd;
  groups_second_half(S);
  for (int s = 0; s < 3; s++) { g.b_nbun[3 * b + s] = (uint32_t)S.nbun[s]; g.b_nbn[3 * b + s] = (uint32_t)S.nbn[s]; }
}

// rg_off = exclusive prefix of rg_cnt
__global__ void __launch_bounds__(256) groups_rg_kernel(DevBatchView v, GroupsView g) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= v.n_reads) return;
  const uint32_t cnt = g.rg_off[n + 1] - g.rg_off[n];
  const size_t src = g.rcap_off[n], dst = g.rg_off[n];
  for (uint32_t k = 0; k < cnt; k++) g.d_rg[dst + k] = g.rg[src + k];
}

__global__ void __launch_bounds__(128) groups_junc_kernel(DevBatchView v, GroupsView g) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= v.n_bundles) return;
  const int lane = threadIdx.x & 31;
  const uint32_t r0 = v.b_read_off[b], nr = v.b_read_off[b + 1] - r0;
  const uint32_t j0 = v.b_junc_off[b], nj0 = v.b_junc_off[b + 1] - j0;
  const uint32_t napp = g.b_napp[b + 1] - g.b_napp[b];
  const uint32_t i0 = nr ? v.r_seg_off[r0] - r0 : 0;
  const size_t o0 = (size_t)j0 + g.b_napp[b];
  const int32_t* perm = g.jt_perm + j0 + i0;
  for (uint32_t k = lane; k < nj0 + napp; k += 32) {
    const uint32_t r = nr ? (uint32_t)perm[k] : k;   // a bundle without reads keeps the order of the junction pass
    const size_t o = o0 + k;
    if (r < nj0) {
      const size_t j = (size_t)j0 + r;
      g.oj_start[o] = v.j_start[j]; g.oj_end[o] = v.j_end[j]; g.oj_strand[o] = g.jt_strand[j];
      g.oj_nreads[o] = v.jp_nreads[j]; g.oj_nreads_good[o] = v.jp_nreads_good[j]; g.oj_left[o] = v.jp_left[j];
      g.oj_right[o] = v.jp_right[j]; g.oj_nm[o] = v.j_nm[j]; g.oj_mm[o] = g.jt_mm[j];
    } else {
      const size_t a = (size_t)i0 + (r - nj0);
      g.oj_start[o] = g.ap_start[a]; g.oj_end[o] = g.ap_end[a]; g.oj_strand[o] = g.ap_strand[a];
      g.oj_nreads[o] = 0; g.oj_nreads_good[o] = 0; g.oj_left[o] = 0; g.oj_right[o] = 0; g.oj_nm[o] = 0;
      g.oj_mm[o] = g.ap_mm[a] ? -1.0 : 0.0;
    }
  }
}

}  // namespace

cudaError_t launch_group_caps(const DevBatchView& v, uint32_t* rcap, unsigned long long* total, int32_t* rj, cudaStream_t st) {
  group_caps_kernel<<<(unsigned)((v.n_reads + 1 + STG_READ_BLOCK - 1) / STG_READ_BLOCK), STG_READ_BLOCK, 0, st>>>(v, rcap, total, rj);
  return cudaGetLastError();
}

// three size classes side by side on three streams: the few oversized bundles (one cluster of CTAs each, below), the deep
// ones (4 warps), the shallow mass (2 warps)
template <int W, int MINB, bool GROW>
static cudaError_t walk_launch(const DevBatchView& v, const GroupsView& g, const JuncParams& p, unsigned grid, uint32_t base, cudaStream_t st) {
  const size_t dyn = 2 * (size_t)(32 * W - 32) * sizeof(FastRec);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(groups_walk_kernel<W, MINB, GROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)) != cudaSuccess) return e;
  // per device and cheap: the kernels must prefer the same carve-out to be able to share an SM.  72 % (164 KB of
  // shared memory, ~90 KB of L1) measured best on C2: 100 % 121 ms, 72-80 % 110 ms, 50 % 133 ms (too few shallow CTAs)
  if ((e = cudaFuncSetAttribute(groups_walk_kernel<W, MINB, GROW>, cudaFuncAttributePreferredSharedMemoryCarveout, 72)) != cudaSuccess) return e;
  groups_walk_kernel<W, MINB, GROW><<<grid, 32 * W, dyn, st>>>(v, g, p, base);
  return cudaGetLastError();
}

cudaError_t launch_groups_walk(const DevBatchView& v, const GroupsView& g, const JuncParams& p, int64_t n_cluster, int64_t n_deep,
                               cudaStream_t st, cudaStream_t aux, cudaStream_t aux3, cudaEvent_t fork, cudaEvent_t join, cudaEvent_t join3) {
  if (v.n_bundles == 0) return cudaSuccess;
  cudaError_t e = cudaSuccess, e2;
  bool forked_aux = false, forked_aux3 = false;
  if (n_deep > 0) e = cudaEventRecord(fork, st);
  if (e == cudaSuccess && n_cluster > 0) {   // the oversized bundles: one cluster of CTAs each
    if ((e = cudaStreamWaitEvent(aux3, fork, 0)) == cudaSuccess) {
      forked_aux3 = true;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)CLUSTER_CTAS); cfg.blockDim = dim3(CLUSTER_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = aux3;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLUSTER_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      static int resident = 0;   // clusters the device holds at once (they are persistent: see the kernel)
      if (!resident) {
        int nc = 0;
        cfg.gridDim = dim3((unsigned)(64 * CLUSTER_CTAS));
        if (cudaOccupancyMaxActiveClusters(&nc, groups_walk_cluster_kernel, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); nc = 8; }
        resident = nc;
      }
      const int64_t nclu = n_cluster < resident ? n_cluster : resident;
      cfg.gridDim = dim3((unsigned)(nclu * CLUSTER_CTAS));
      e = cudaLaunchKernelEx(&cfg, groups_walk_cluster_kernel, v, g, p, 0u, (uint32_t)n_cluster);
    }
  }
  if (e == cudaSuccess && n_deep > n_cluster) {   // the deep bundles next to the shallow mass
    if ((e = cudaStreamWaitEvent(aux, fork, 0)) == cudaSuccess) {
      forked_aux = true;
      e = walk_launch<4, 4, false>(v, g, p, (unsigned)(n_deep - n_cluster), (uint32_t)n_cluster, aux);
    }
  }
  if (e == cudaSuccess && v.n_bundles > n_deep) e = walk_launch<2, 8, true>(v, g, p, (unsigned)(v.n_bundles - n_deep), (uint32_t)n_deep, st);
  // join: also on the error path, so that nothing still runs on the side streams when the caller releases the scratch
  if (forked_aux) { if ((e2 = cudaEventRecord(join, aux)) == cudaSuccess) e2 = cudaStreamWaitEvent(st, join, 0); if (e == cudaSuccess) e = e2; }
  if (forked_aux3) { if ((e2 = cudaEventRecord(join3, aux3)) == cudaSuccess) e2 = cudaStreamWaitEvent(st, join3, 0); if (e == cudaSuccess) e = e2; }
  return e;
}

void groups_set_debug(int on) { cudaMemcpyToSymbol(g_groups_debug, &on, sizeof(int)); }
// measured on C2 (profiles/r3_groups_history.md), whole stage: 400 k 138 ms, 200 k 122 ms, 100 k 177 ms, 50 k 288 ms (more
// clusters than the GPU holds at once)
#define GROUPS_CLUSTER_READS 200000
int64_t groups_cluster_reads() { return GROUPS_CLUSTER_READS; }
size_t groups_cluster_ctl_bytes() { return sizeof(ClusterCtl); }
size_t groups_cluster_rec_bytes() { return sizeof(FastRec) * CLUSTER_WINDOW; }
int64_t groups_deep_reads() { return GROUPS_DEEP_READS; }

cudaError_t launch_groups_finish(const DevBatchView& v, const GroupsView& g, cudaStream_t st) {
  if (v.n_bundles == 0) return cudaSuccess;
  if (g.n_colors) groups_colcopy_kernel<<<(unsigned)((g.n_colors + 255) / 256), 256, 0, st>>>(v, g, (uint32_t)g.n_colors);
  groups_colour_kernel<<<(unsigned)((v.n_bundles + 3) / 4), 128, 0, st>>>(v, g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (v.n_reads) groups_rg_kernel<<<(unsigned)((v.n_reads + 255) / 256), 256, 0, st>>>(v, g);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  groups_junc_kernel<<<(unsigned)((v.n_bundles + 3) / 4), 128, 0, st>>>(v, g);
  return cudaGetLastError();
}
