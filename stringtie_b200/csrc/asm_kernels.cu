// asm_kernels.cu — sm_100a kernels of SURVEY.md §8 rows a10–a15: splice graphs, read -> transfrag mapping, transfrag
// processing, flow decomposition, predictions.  The per-graph / per-read logic lives in asm_device.cuh and frag_device.cuh
// (one WARP per graph or per bundle, one THREAD per read); this file maps it onto the batch:
//
//   enum / junc_lists / bounds          graph ids, strand-filtered junction lists of every bundle, capacity bounds
//   sweep -> sizes -> finish            create_graph: the event sweep into bound-sized temporaries, exact sizes, then
//                                       adjacency, coverage-drop links, DFS transfrags, reachability (rlink.cpp:3579-4351)
//   frag_count -> frag_write, tf0_write one thread per read: node-coverage contributions and transfrag keys of the fragments
//                                       it opens (rlink.cpp:15796-15820, 4870-5146), two passes around a scan
//   tf_commit, nodecov_commit           one warp per bundle: keys -> transfrags in order of first appearance, abundances and
//                                       node coverages as float sums in read order (update_abundance :4680, :4395)
//   tf_stats -> flow_sizes -> tf_gather per-graph transfrag lists in order
//   flow                                one warp per graph: process_transfrags :5768, find_transcripts / parse_trf :13633 /
//                                       :11992, push_max_flow :9075, store_transcript :9688
//   pred_*                              predictions into graph order, gene numbers, exon CSR
//
// Everything is HBM / L2 resident; graphs of a batch run side by side (a C2 batch has ~10^5 of them).
#include "asm_device.cuh"
#include "frag_device.cuh"

#define ASM_ERR_UNITIG 0x800u     // a unitig (super-read) in the batch: srabund / process_srfrag are not built

namespace {

__device__ __forceinline__ uint32_t bundle_of_read(const DevBatchView& v, uint32_t n) {
  uint32_t b = v.cta_bundle[n / STG_READ_BLOCK];
  while (v.b_read_off[b + 1] <= n) b++;
  return b;
}

struct LaneBase { uint32_t g0, ng; size_t base; };
// strand lane `lane` (0, 1, 2) of bundle b in the [3 * n_groups] arrays of GroupsView
__device__ __forceinline__ LaneBase lane_base(const GroupsView& g, uint32_t b, int lane) {
  LaneBase L;
  L.g0 = g.b_ng[b]; L.ng = g.b_ng[b + 1] - L.g0;
  L.base = 3 * (size_t)L.g0 + (size_t)lane * L.ng;
  return L;
}

__global__ void __launch_bounds__(256) asm_enum_kernel(DevBatchView v, AsmView A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * v.n_bundles) return;
  const uint32_t g0 = A.gr_first[i], g1 = A.gr_first[i + 1];
  for (uint32_t g = g0; g < g1; g++) { A.g_bundle[g] = (int32_t)(i >> 1); A.g_s[g] = (int8_t)(i & 1); A.g_k[g] = (int32_t)(g - g0); }
}

__global__ void __launch_bounds__(256) asm_junc_lists_kernel(DevBatchView v, GroupsView g, AsmView A) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= 2 * v.n_bundles) return;
  const uint32_t b = (uint32_t)(w >> 1);
  const int s = (int)(w & 1);
  const uint32_t j0 = A.jrow_off[b], nj = A.jrow_off[b + 1] - j0;
  const size_t o = 2 * (size_t)j0 + (size_t)s * nj;
  const int n = junc_lists_warp(g.oj_start + j0, g.oj_end + j0, g.oj_strand + j0, (int)nj, 2 * s - 1, A.js_start + o, A.js_end + o,
                                A.je_end + o, A.je_k + o);
  if ((threadIdx.x & 31) == 0) A.jl_cnt[w] = (uint32_t)n;
}

// is the CBundle worth a graph (rlink.cpp:15721-15724, no guides) and how much room can its sweep need
__global__ void __launch_bounds__(256) asm_bounds_kernel(DevBatchView v, GroupsView g, AsmView A, AsmParams P) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi > A.n_graphs) return;
  if (gi == A.n_graphs) { A.cap_n[gi] = 0; A.cap_e[gi] = 0; A.cap_f[gi] = 0; A.cap_t[gi] = 0; return; }
  const uint32_t b = (uint32_t)A.g_bundle[gi];
  const int s = A.g_s[gi];
  const LaneBase L = lane_base(g, b, 2 * s);
  const size_t k = L.base + (size_t)A.g_k[gi];
  const float cov = g.bun_cov[k];
  const bool process = cov != 0.f && (g.bun_multi[k] / cov) <= P.mcov && g.bun_len[k] >= P.mintranscriptlen;
  uint32_t ncap = 0, ecap = 0, fcap = 0, tcap = 0;
  if (process) {
    uint32_t maxlen = 0, lo = 0xffffffffu, hi = 0;
    ncap = 2;
    for (int x = g.bun_start[k]; x != -1; x = g.bn_next[L.base + x]) {
      const uint32_t a = g.bn_start[L.base + x], z = g.bn_end[L.base + x];
      const uint32_t len = z - a + 1;
      if (len > maxlen) maxlen = len;
      if (a < lo) lo = a;
      if (z > hi) hi = z;
      ncap += 2 + len / 50;
    }
    // junctions that can cut this graph's nodes: starts / ends inside its span
    const uint32_t j0 = A.jrow_off[b], nj = A.jrow_off[b + 1] - j0;
    const size_t o = 2 * (size_t)j0 + (size_t)s * nj;
    const int nJ = (int)A.jl_cnt[2 * b + s];
    auto lower = [&](const uint32_t* a, uint32_t key) { int l = 0, h = nJ; while (l < h) { const int m = (l + h) >> 1; if (a[m] < key) l = m + 1; else h = m; } return l; };
    const int js_in = lower(A.js_start + o, hi + 1) - lower(A.js_start + o, lo);
    const int je_in = lower(A.je_end + o, hi + 1) - lower(A.je_end + o, lo);
    ncap += 2u * (uint32_t)(js_in + je_in);
    ecap = (uint32_t)js_in + 6u * ncap;
    fcap = 4u * ncap;
    tcap = maxlen / 50 + 4;
  }
  A.cap_n[gi] = ncap; A.cap_e[gi] = ecap; A.cap_f[gi] = fcap; A.cap_t[gi] = tcap;
}

__device__ __forceinline__ void make_graph_src(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, int64_t gi,
                                               GraphSrc& G, LaneBase& L, size_t& k) {
  const uint32_t b = (uint32_t)A.g_bundle[gi];
  const int s = A.g_s[gi];
  L = lane_base(g, b, 2 * s);
  k = L.base + (size_t)A.g_k[gi];
  const float* base = v.bpcov + v.b_gbase[b];
  const size_t stride = v.cov_pitch;
  G.s = s; G.refstart = (uint32_t)v.b_start[b]; G.covlen = v.b_len[b];
  G.c_all = base + stride; G.c_opp = base + (size_t)(2 - 2 * s) * stride;
  G.bn_start = g.bn_start + L.base; G.bn_end = g.bn_end + L.base; G.bn_next = g.bn_next + L.base;
  G.startnode = g.bun_start[k]; G.lastnode = g.bun_last[k];
  const uint32_t j0 = A.jrow_off[b], nj = A.jrow_off[b + 1] - j0;
  const size_t o = 2 * (size_t)j0 + (size_t)s * nj;
  G.js_start = A.js_start + o; G.js_end = A.js_end + o; G.je_end = A.je_end + o; G.je_k = A.je_k + o;
  G.nJs = G.nJe = (int)A.jl_cnt[2 * b + s];
  G.reg_node = A.reg_node + o; G.reg_graph = A.reg_graph + o;
  G.gid = (int)gi; G.junctionsupport = P.junctionsupport; G.trim = P.trim;
}

__device__ __forceinline__ void make_sweep_out(const AsmView& A, int64_t gi, const LaneBase& L, SweepOut& o) {
  const uint32_t n0 = A.cap_n[gi], e0 = A.cap_e[gi], f0 = A.cap_f[gi], t0 = A.cap_t[gi];
  o.n_start = A.t_nstart + n0; o.n_end = A.t_nend + n0; o.ev_u = A.t_evu + e0; o.ev_v = A.t_evv + e0;
  o.ft_n1 = A.t_ft1 + f0; o.ft_n2 = A.t_ft2 + f0; o.ft_ab = A.t_ftab + f0;
  o.bn_first = A.bn_first + L.base; o.bn_cnt = A.bn_cnt + L.base;
  o.tr_pos = A.t_trpos + t0; o.tr_ab = A.t_trab + t0; o.tr_st = A.t_trst + t0; o.tr_cap = A.cap_t[gi + 1] - t0;
  o.ncap = (int)(A.cap_n[gi + 1] - n0); o.ecap = (int)(A.cap_e[gi + 1] - e0); o.fcap = (int)(A.cap_f[gi + 1] - f0);
}

__global__ void __launch_bounds__(128) asm_sweep_kernel(DevBatchView v, GroupsView g, AsmView A, AsmParams P) {
  const int64_t gi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gi >= A.n_graphs) return;
  const int lane = threadIdx.x & 31;
  if (A.cap_n[gi + 1] == A.cap_n[gi]) { if (lane == 0) { A.sw_nn[gi] = 0; A.sw_ne[gi] = 0; A.sw_nf[gi] = 0; A.sw_edgeno[gi] = 0; } return; }
  GraphSrc G; LaneBase L; size_t k;
  make_graph_src(v, g, A, P, gi, G, L, k);
  SweepOut o;
  make_sweep_out(A, gi, L, o);
  graph_sweep(G, o);
  if (lane == 0) {
    A.sw_nn[gi] = o.nn; A.sw_ne[gi] = o.ne; A.sw_nf[gi] = o.nf; A.sw_edgeno[gi] = o.edgeno;
    if (o.err) atomicOr(v.err_flag, o.err);
    for (int x = G.startnode; x != -1; x = G.bn_next[x]) A.bn_graph[L.base + x] = (int32_t)gi;
  }
}

__global__ void __launch_bounds__(256) asm_sizes_kernel(AsmView A) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi > A.n_graphs) return;
  uint32_t xn = 0, xe = 0, xt = 0, xr = 0;
  if (gi < A.n_graphs && A.sw_nn[gi] > 0) {
    const uint32_t nn = (uint32_t)A.sw_nn[gi], gno = nn + 1;
    xn = gno; xe = (uint32_t)A.sw_ne[gi] + 2 * nn + 4; xt = (uint32_t)A.sw_nf[gi] + 4 * nn + 4; xr = gno * ((gno + 63) >> 6);
  }
  A.x_node[gi] = xn; A.x_edge[gi] = xe; A.x_tf0[gi] = xt; A.x_reach[gi] = xr;
}

__device__ __forceinline__ void bind_graph(const AsmView& A, int64_t gi, Graph& gr) {
  const uint32_t n0 = A.x_node[gi], e0 = A.x_edge[gi], t0 = A.x_tf0[gi], r0 = A.x_reach[gi];
  gr.n_start = A.n_start + n0; gr.n_end = A.n_end + n0; gr.n_cov = A.n_cov + n0;
  gr.ch_off = A.ch_off + n0; gr.ch_cnt = A.ch_cnt + n0; gr.ch = A.ch + e0;
  gr.pa_off = A.pa_off + n0; gr.pa_cnt = A.pa_cnt + n0; gr.pa = A.pa + e0;
  gr.n_sinkpar = A.n_sinkpar + n0; gr.reach = A.reach + r0;
  gr.tf_n1 = A.tf_n1 + t0; gr.tf_n2 = A.tf_n2 + t0; gr.tf_ab = A.tf_ab + t0;
}

__global__ void __launch_bounds__(128) asm_finish_kernel(DevBatchView v, GroupsView g, AsmView A, AsmParams P) {
  const int64_t gi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gi >= A.n_graphs) return;
  const int lane = threadIdx.x & 31;
  Graph gr;
  gr.gno = 0; gr.nwn = 0; gr.ntf0 = 0; gr.edgeno = 0;
  bind_graph(A, gi, gr);
  if (A.sw_nn[gi] > 0) {
    GraphSrc G; LaneBase L; size_t k;
    make_graph_src(v, g, A, P, gi, G, L, k);
    SweepOut o;
    make_sweep_out(A, gi, L, o);
    o.nn = A.sw_nn[gi]; o.ne = A.sw_ne[gi]; o.nf = A.sw_nf[gi]; o.edgeno = A.sw_edgeno[gi]; o.err = 0;
    uint32_t err = 0;
    const size_t sb = (size_t)A.x_node[gi] + 2 * (size_t)gi;
    graph_finish(G, o, gr, A.fin_stack + 2 * sb, A.fin_ncs + sb, A.fin_visit + sb, P.allowed_nodes, &err);
    if (lane == 0 && err) atomicOr(v.err_flag, err);
    gr.ntf0 = __shfl_sync(0xffffffffu, gr.ntf0, 0); gr.edgeno = __shfl_sync(0xffffffffu, gr.edgeno, 0);
  }
  if (lane == 0) { A.graphs[gi] = gr; A.tf0_off[gi] = (uint32_t)gr.ntf0; A.g_node_off[gi] = (int32_t)A.x_node[gi]; }
  if (gi == 0 && lane == 0) A.tf0_off[A.n_graphs] = 0;
}

// ---- a11 ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void make_bundle_ctx(const DevBatchView& v, const GroupsView& g, const AsmView& A, uint32_t b, BundleCtx& B) {
  const uint32_t r0 = v.b_read_off[b];
  const uint32_t seg0 = v.r_seg_off[r0], pair0 = v.r_pair_off[r0];
  B.nr = (int)(v.b_read_off[b + 1] - r0);
  const uint32_t g0 = g.b_ng[b];
  B.ng = (int)(g.b_ng[b + 1] - g0);
  B.r_start = v.r_start + r0; B.r_end = v.r_end + r0; B.r_nh = v.r_nh + r0; B.r_flags = v.r_flags + r0; B.r_count = v.r_count + r0;
  B.strand = g.strand + r0;
  B.seg_off = v.r_seg_off + r0; B.seg0 = seg0; B.seg_s = g.seg_start + seg0; B.seg_e = g.seg_end + seg0;
  B.pair_off = v.r_pair_off + r0; B.pair0 = pair0; B.pair = g.pair + pair0; B.pair_cnt = v.pair_cnt + pair0;
  B.rj = g.rj + (seg0 - r0);
  const uint32_t j0 = A.jrow_off[b];
  B.j_start = g.oj_start + j0; B.j_end = g.oj_end + j0; B.j_strand = g.oj_strand + j0;
  B.rg_off = g.rg_off + r0; B.rg0 = g.rg_off[r0]; B.rg = g.d_rg + B.rg0;
  B.merge = g.d_g_merge + g0; B.negp = g.d_negp + g0; B.g2b = g.d_g2b + 3 * (size_t)g0;
  for (int s = 0; s < 2; s++) {
    const size_t o = 3 * (size_t)g0 + (size_t)(2 * s) * (size_t)B.ng;
    B.bn_graph[s] = A.bn_graph + o; B.bn_first[s] = A.bn_first + o; B.bn_cnt[s] = A.bn_cnt + o;
  }
  B.graphs = A.graphs;
}

__global__ void __launch_bounds__(STG_READ_BLOCK) asm_frag_count_kernel(DevBatchView v, GroupsView g, AsmView A) {
  const int64_t n = (int64_t)blockIdx.x * STG_READ_BLOCK + threadIdx.x;
  if (n > v.n_reads) return;
  if (n == v.n_reads) { A.r_ncov[n] = 0; A.r_ntf[n] = 0; A.r_nkey[n] = 0; return; }
  const uint32_t b = bundle_of_read(v, (uint32_t)n);
  BundleCtx B;
  make_bundle_ctx(v, g, A, b, B);
  FragCount fc{0, 0, 0};
  uint32_t err = 0;
  if (v.r_flags[n] & STG_RF_UNITIG) err |= ASM_ERR_UNITIG;
  else read_fragments(B, (int)((uint32_t)n - v.b_read_off[b]), fc, &err);
  if (err) atomicOr(v.err_flag, err);
  A.r_ncov[n] = fc.ncov; A.r_ntf[n] = fc.ntf; A.r_nkey[n] = fc.nkey;
}

__device__ __forceinline__ void make_frag_write(const AsmView& A, FragWrite& w) {
  w.graph_node_off = A.g_node_off; w.cov_node = A.cov_node; w.cov_val = A.cov_val; w.tf_graph = A.rec_graph; w.tf_ab = A.rec_ab;
  w.tf_key_off = A.rec_koff; w.tf_key_len = A.rec_klen; w.tf_hash = A.rec_hash; w.key = A.key;
  w.icov = 0; w.itf = 0; w.ikey = 0;
}

__global__ void __launch_bounds__(STG_READ_BLOCK) asm_frag_write_kernel(DevBatchView v, GroupsView g, AsmView A) {
  const int64_t n = (int64_t)blockIdx.x * STG_READ_BLOCK + threadIdx.x;
  if (n >= v.n_reads) return;
  if (A.r_ncov[n + 1] == A.r_ncov[n] && A.r_ntf[n + 1] == A.r_ntf[n]) return;
  const uint32_t b = bundle_of_read(v, (uint32_t)n);
  BundleCtx B;
  make_bundle_ctx(v, g, A, b, B);
  FragWrite w;
  make_frag_write(A, w);
  w.icov = A.r_ncov[n]; w.itf = (uint32_t)A.n_tf0 + A.r_ntf[n]; w.ikey = 2u * (uint32_t)A.n_tf0 + A.r_nkey[n];
  uint32_t err = 0;
  read_fragments(B, (int)((uint32_t)n - v.b_read_off[b]), w, &err);
}

// the transfrags create_graph left, as records in front of the reads' (record t0 of graph g at tf0_off[g] + t0)
__global__ void __launch_bounds__(256) asm_tf0_write_kernel(AsmView A) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= A.n_graphs) return;
  const Graph& gr = A.graphs[gi];
  if (!gr.gno) return;
  FragWrite w;
  make_frag_write(A, w);
  w.itf = A.tf0_off[gi]; w.ikey = 2u * A.tf0_off[gi];
  for (int t = 0; t < gr.ntf0; t++) {
    const int32_t nd[2] = {gr.tf_n1[t], gr.tf_n2[t]};
    const uint8_t fl[2] = {0, 1};
    w.tf((int)gi, nd, fl, 2, gr.tf_ab[t]);
  }
}

__device__ __forceinline__ uint32_t bundle_nrec(const DevBatchView& v, const AsmView& A, uint32_t b, uint32_t* nt0) {
  const uint32_t g0 = A.gr_first[2 * b], g1 = A.gr_first[2 * b + 2];
  *nt0 = A.tf0_off[g1] - A.tf0_off[g0];
  return A.r_ntf[v.b_read_off[b + 1]] - A.r_ntf[v.b_read_off[b]];
}

__global__ void __launch_bounds__(256) asm_table_caps_kernel(DevBatchView v, AsmView A) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > v.n_bundles) return;
  uint32_t cap = 0;
  if (b < v.n_bundles) {
    uint32_t nt0;
    const uint32_t nrec = bundle_nrec(v, A, (uint32_t)b, &nt0) + nt0;
    if (nrec) { cap = 16; while (cap < 2 * nrec + 2) cap <<= 1; }
  }
  A.tab_off[b] = cap;
}

// first transfrag slot of bundle b in tf_rep / tf_abund / tf_graph
__device__ __forceinline__ uint32_t tf_slot0(const DevBatchView& v, const AsmView& A, uint32_t b) {
  return A.tf0_off[A.gr_first[2 * b]] + A.r_ntf[v.b_read_off[b]];
}

// per graph: how many transfrags, how many of them have a pair of consecutive nodes without a pattern edge, and the
// sum of their node spans (bounds of the incidence lists of a12)
__global__ void __launch_bounds__(256) asm_tf_stats_kernel(DevBatchView v, AsmView A) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= v.n_bundles) return;
  const int lane = threadIdx.x & 31;
  const uint32_t s0 = tf_slot0(v, A, (uint32_t)w);
  const int ntf = A.b_ntf[w];
  for (int t = lane; t < ntf; t += 32) {
    const int gi = A.tf_graph[s0 + t];
    const uint32_t r = (uint32_t)A.tf_rep[s0 + t];
    const uint32_t* k = A.key + A.rec_koff[r];
    const uint32_t n = A.rec_klen[r];
    bool gap = false;
    for (uint32_t i = 1; i < n; i++) if (!(k[i] & FR_FLAG)) gap = true;
    atomicAdd(A.g_nt + gi, 1u);
    if (gap) atomicAdd(A.g_ngap + gi, 1u);
    atomicAdd(A.g_reachsum + gi, (k[n - 1] & ~FR_FLAG) - (k[0] & ~FR_FLAG) + 1u);
  }
}

__global__ void __launch_bounds__(256) asm_flow_sizes_kernel(AsmView A) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi > A.n_graphs) return;
  uint32_t ytf = 0, ypat = 0, ypath = 0, yntrf = 0, yistr = 0;
  if (gi < A.n_graphs) {
    const Graph& gr = A.graphs[gi];
    if (gr.gno) {
      const uint32_t nedges = (uint32_t)(gr.ch_off[gr.gno - 1] + gr.ch_cnt[gr.gno - 1]);
      uint32_t gmax = 1;
      for (int i = 1; i < gr.gno; i++) if ((uint32_t)gr.ch_cnt[i] > gmax) gmax = (uint32_t)gr.ch_cnt[i];
      ytf = A.g_nt[gi] + nedges + 1;
      ypat = ytf * (uint32_t)gr.nwn;
      ypath = A.g_ngap[gi] * gmax + 1;
      yntrf = A.g_reachsum[gi] + 2 * nedges + 2;
      yistr = (ytf + 63) / 64 + 1;
    }
  }
  A.y_tf[gi] = ytf; A.y_pat[gi] = ypat; A.y_path[gi] = ypath; A.y_ntrf[gi] = yntrf; A.y_istr[gi] = yistr;
}

// transfrags of a bundle, in id order, into the working slots of their graphs (order inside a graph = id order)
__global__ void __launch_bounds__(128) asm_tf_gather_kernel(DevBatchView v, AsmView A) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= v.n_bundles) return;
  const int lane = threadIdx.x & 31;
  const uint32_t s0 = tf_slot0(v, A, (uint32_t)b);
  const int ntf = A.b_ntf[b];
  for (int t0 = 0; t0 < ntf; t0 += 32) {
    const int t = t0 + lane;
    const int gi = t < ntf ? A.tf_graph[s0 + t] : -1 - lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, gi);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (t < ntf) base = A.g_fill[gi];
    __syncwarp();
    if (t < ntf) {
      const uint32_t slot = A.y_tf[gi] + base + (uint32_t)rank;
      const uint32_t r = (uint32_t)A.tf_rep[s0 + t];
      A.f_wkoff[slot] = A.rec_koff[r]; A.f_wklen[slot] = A.rec_klen[r]; A.f_wab[slot] = A.tf_abund[s0 + t];
      if (rank + 1 == __popc(peers)) A.g_fill[gi] = base + (uint32_t)__popc(peers);   // the last of the group moves the cursor
    }
    __syncwarp();
  }
}

// ---- a12–a15 ---------------------------------------------------------------------------------------------------------
// The graphs of a batch differ in size by four orders of magnitude (a handful of transfrags to tens of thousands), and one
// warp works on one graph: handed out in bundle order, the large graphs of the last bundles would run alone at the end.
// So the graphs are first binned by a quasi-logarithmic size key (heaviest bin first: the classic longest-processing-time
// order; the order inside a bin does not matter) and persistent warps pull them from that list through a ticket counter.
#define FLOW_BINS 256
__device__ __forceinline__ uint32_t flow_bin(uint32_t work) {   // 8 sub-bins per power of two, heaviest first
  if (work < 8u) return FLOW_BINS - 1u - work;
  const int e = 31 - __clz(work);
  const uint32_t b = (uint32_t)(e - 2) * 8u + ((work >> (e - 3)) & 7u);
  return b >= FLOW_BINS ? 0u : FLOW_BINS - 1u - b;
}
__device__ __forceinline__ uint32_t flow_work(const AsmView& A, int64_t gi) { return A.g_nt[gi] * 4u + (uint32_t)A.graphs[gi].gno; }

__global__ void __launch_bounds__(256) asm_flow_bin_kernel(AsmView A, uint32_t* __restrict__ bins) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi < A.n_graphs) atomicAdd(&bins[flow_bin(flow_work(A, gi))], 1u);
}
__global__ void __launch_bounds__(32) asm_flow_bin_scan_kernel(uint32_t* __restrict__ bins) {
  if (threadIdx.x == 0) { uint32_t run = 0; for (int b = 0; b < FLOW_BINS; b++) { const uint32_t c = bins[b]; bins[b] = run; run += c; } }
}
__global__ void __launch_bounds__(256) asm_flow_order_kernel(AsmView A, uint32_t* __restrict__ bins, uint32_t* __restrict__ order) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi < A.n_graphs) order[atomicAdd(&bins[flow_bin(flow_work(A, gi))], 1u)] = (uint32_t)gi;
}

__device__ int g_asm_debug = 0;   // STGPU_ASM_DEBUG: the graphs that keep a warp of asm_flow longest
__device__ void asm_flow_graph(const DevBatchView& v, const AsmView& A, const AsmParams& P, const int64_t gi) {
  const int lane = threadIdx.x & 31;
  const Graph gr = A.graphs[gi];
  if (!gr.gno) { if (lane == 0) A.g_npred[gi] = 0; return; }
  TfG f;
  const uint32_t t0 = A.y_tf[gi];
  f.nt = (int)A.g_nt[gi]; f.cap = (int)(A.y_tf[gi + 1] - t0);
  f.koff = A.f_koff + t0; f.klen = A.f_klen + t0; f.key = A.key; f.xkey0 = A.xkey_base + 2u * A.x_edge[gi] + 2u * (uint32_t)gi;
  f.ab = A.f_ab + t0; f.real = A.f_real + t0; f.usepath = A.f_usepath + t0; f.path_off = A.f_path_off + t0; f.path_cnt = A.f_path_cnt + t0;
  f.path_node = A.f_path_node + A.y_path[gi]; f.path_cont = A.f_path_cont + A.y_path[gi]; f.path_ab = A.f_path_ab + A.y_path[gi];
  f.path_cap = (int)(A.y_path[gi + 1] - A.y_path[gi]) - 1; f.path_used = 0;
  f.pat = A.f_pat + A.y_pat[gi];
  const size_t nb = (size_t)A.x_node[gi] + 2 * (size_t)gi;   // per-node scratch base (gno + 2 entries per graph)
  f.ntrf_off = A.f_ntrf_off + nb; f.ntrf_cnt = A.f_ntrf_cnt + nb; f.ntrf = A.f_ntrf + A.y_ntrf[gi]; f.ntrf_cap = (int)(A.y_ntrf[gi + 1] - A.y_ntrf[gi]);
  f.order = A.f_order + t0; f.w_koff = A.f_wkoff + t0; f.w_klen = A.f_wklen + t0; f.w_ab = A.f_wab + t0; f.skey = A.f_skey + t0;
  f.e_cov = A.f_ecov + A.x_edge[gi];
  uint32_t err = 0;
  const long long tp0 = g_asm_debug ? clock64() : 0;
  const int nt = process_transfrags_warp(gr, f, &err);
  f.nt = nt;
  const long long tp1 = g_asm_debug ? clock64() : 0;
  FlowScratch S;
  S.nodecov = A.s_nodecov + nb; S.path = A.s_path + nb; S.npath = 0; S.succ = A.s_succ + nb;
  S.pathpat = A.s_pathpat + nb;          // nwn <= gno words
  S.istr = A.s_istr + A.y_istr[gi];
  S.usednode = A.s_used + nb; S.prev_node = A.s_prevn + nb; S.prev_edge = A.s_preve + A.x_edge[gi]; S.node2path = A.s_n2p + nb;
  S.capl = A.s_capl + nb; S.capr = A.s_capr + nb; S.suml = A.s_suml + nb; S.sumr = A.s_sumr + nb; S.nodeflux = A.s_flux + nb;
  PredOut out;
  out.counters = A.p_counters; out.cap_pred = A.cap_pred; out.cap_exon = A.cap_exon;
  out.p_graph = A.pp_graph; out.p_seq = A.pp_seq; out.p_start = A.pp_start; out.p_end = A.pp_end; out.p_cov = A.pp_cov; out.p_tlen = A.pp_tlen;
  out.p_exon0 = A.pp_exon0; out.p_nex = A.pp_nex; out.e_start = A.pe_start; out.e_end = A.pe_end; out.e_cov = A.pe_cov;
  int ns = 0;
  if (!err) ns = find_transcripts_warp(gr, f, S, P, out, (int)gi, A.s_exs + nb, A.s_exe + nb, A.s_exc + nb, &err);
  if (g_asm_debug && lane == 0 && clock64() - tp0 > 8000000)
    printf("asm_flow graph %lld: %d nodes, %d transfrags, %d paths tried, %d stored | Mcycles: process_transfrags %.1f find_transcripts %.1f\n", (long long)gi,
           gr.gno, nt, S.dbg_iters, ns, (tp1 - tp0) * 1e-6, (clock64() - tp1) * 1e-6);
  if (lane == 0) { A.g_npred[gi] = (uint32_t)ns; if (err) atomicOr(v.err_flag, err); }
  if (gi == 0 && lane == 0) A.g_npred[A.n_graphs] = 0;
}

__global__ void __launch_bounds__(128) asm_flow_kernel(DevBatchView v, AsmView A, AsmParams P, const uint32_t* __restrict__ order,
                                                        uint32_t* __restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(ticket, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= (uint32_t)A.n_graphs) return;
    asm_flow_graph(v, A, P, (int64_t)order[i]);
    __syncwarp();
  }
}

// gene numbers: the counter runs on from the neutral bundles' (rlink.cpp:15491) and takes one step per graph that
// stored a transcript (find_transcripts' `first`, store_transcript :9799)
__global__ void __launch_bounds__(256) asm_geneno_kernel(DevBatchView v, AsmView A, const int32_t* b_geneno) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > v.n_bundles) return;
  if (b == v.n_bundles) { A.o_pred_off[b] = A.g_npred[A.n_graphs]; return; }
  int geneno = b_geneno ? b_geneno[b] : 0;
  const uint32_t g0 = A.gr_first[2 * b], g1 = A.gr_first[2 * b + 2];
  for (uint32_t gi = g0; gi < g1; gi++) {
    if (A.g_npred[gi + 1] > A.g_npred[gi]) geneno++;
    A.g_geneno[gi] = geneno - 1;
  }
  A.o_pred_off[b] = A.g_npred[g0];
}

__global__ void __launch_bounds__(256) asm_pred_scatter_kernel(AsmView A, uint32_t npool) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > npool) return;
  if (i == npool) { A.o_exon_off[npool] = 0; return; }
  const int gi = A.pp_graph[i];
  const uint32_t d = A.g_npred[gi] + (uint32_t)A.pp_seq[i];
  A.o_start[d] = A.pp_start[i]; A.o_end[d] = A.pp_end[i]; A.o_cov[d] = A.pp_cov[i]; A.o_tlen[d] = A.pp_tlen[i];
  A.o_geneno[d] = A.g_geneno[gi]; A.o_strand[d] = A.g_s[gi] ? '+' : '-'; A.o_src[d] = (int32_t)i;
  A.o_exon_off[d] = A.pp_nex[i];
}

__global__ void __launch_bounds__(256) asm_exon_scatter_kernel(AsmView A, uint32_t npool) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= npool) return;
  const int32_t i = A.o_src[d];
  const uint32_t e0 = A.pp_exon0[i], n = A.pp_nex[i], o = A.o_exon_off[d];
  for (uint32_t x = 0; x < n; x++) { A.oe_start[o + x] = A.pe_start[e0 + x]; A.oe_end[o + x] = A.pe_end[e0 + x]; A.oe_cov[o + x] = A.pe_cov[e0 + x]; }
}

// global sums the host accumulates per bundle under countMutex (stringtie.cpp:1490-1497): sum of the bundles'
// num_fragments and frag_len, plus the total clamped coverage of lane 1; one block, fixed order (deterministic)
__global__ void __launch_bounds__(1024) asm_totals_kernel(int64_t nb, const double* __restrict__ numfrag, const double* __restrict__ fraglen,
                                                          const double* __restrict__ covtot, double* out) {
  __shared__ double sh[3][1024];
  double a = 0, b = 0, c = 0;
  for (int64_t i = threadIdx.x; i < nb; i += 1024) { a += numfrag ? numfrag[i] : 0.0; b += fraglen ? fraglen[i] : 0.0; c += covtot[3 * i + 1]; }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[1][0]; out[2] = sh[2][0]; }
}

unsigned blocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

cudaError_t launch_asm_totals(int64_t nb, const double* numfrag, const double* fraglen, const double* covtot, double* out, cudaStream_t st) {
  asm_totals_kernel<<<1, 1024, 0, st>>>(nb, numfrag, fraglen, covtot, out);
  return cudaGetLastError();
}

cudaError_t launch_asm_enum(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  if (v.n_bundles) asm_enum_kernel<<<blocks(2 * v.n_bundles, 256), 256, 0, st>>>(v, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_junc_lists(const DevBatchView& v, const GroupsView& g, const AsmView& A, cudaStream_t st) {
  if (v.n_bundles) asm_junc_lists_kernel<<<blocks(2 * v.n_bundles * 32, 256), 256, 0, st>>>(v, g, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_bounds(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, cudaStream_t st) {
  asm_bounds_kernel<<<blocks(A.n_graphs + 1, 256), 256, 0, st>>>(v, g, A, P);
  return cudaGetLastError();
}
cudaError_t launch_asm_sweep(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, cudaStream_t st) {
  if (A.n_graphs) asm_sweep_kernel<<<blocks(A.n_graphs * 32, 128), 128, 0, st>>>(v, g, A, P);
  return cudaGetLastError();
}
cudaError_t launch_asm_sizes(const AsmView& A, cudaStream_t st) {
  asm_sizes_kernel<<<blocks(A.n_graphs + 1, 256), 256, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_asm_finish(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, cudaStream_t st) {
  if (A.n_graphs) asm_finish_kernel<<<blocks(A.n_graphs * 32, 128), 128, 0, st>>>(v, g, A, P);
  return cudaGetLastError();
}
cudaError_t launch_asm_frag_count(const DevBatchView& v, const GroupsView& g, const AsmView& A, cudaStream_t st) {
  asm_frag_count_kernel<<<blocks(v.n_reads + 1, STG_READ_BLOCK), STG_READ_BLOCK, 0, st>>>(v, g, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_frag_write(const DevBatchView& v, const GroupsView& g, const AsmView& A, cudaStream_t st) {
  if (A.n_graphs) asm_tf0_write_kernel<<<blocks(A.n_graphs, 256), 256, 0, st>>>(A);
  if (v.n_reads) asm_frag_write_kernel<<<blocks(v.n_reads, STG_READ_BLOCK), STG_READ_BLOCK, 0, st>>>(v, g, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_table_caps(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  asm_table_caps_kernel<<<blocks(v.n_bundles + 1, 256), 256, 0, st>>>(v, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_tf_stats(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  if (v.n_bundles) asm_tf_stats_kernel<<<blocks(v.n_bundles * 32, 256), 256, 0, st>>>(v, A);
  asm_flow_sizes_kernel<<<blocks(A.n_graphs + 1, 256), 256, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_asm_gather(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  if (v.n_bundles) asm_tf_gather_kernel<<<blocks(v.n_bundles * 32, 128), 128, 0, st>>>(v, A);
  return cudaGetLastError();
}
// scratch: FLOW_BINS + 1 words (zeroed here) and one word per graph
cudaError_t launch_asm_flow(const DevBatchView& v, const AsmView& A, const AsmParams& P, uint32_t* bins, uint32_t* order, int num_sms,
                            cudaStream_t st) {
  if (!A.n_graphs) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(bins, 0, sizeof(uint32_t) * (FLOW_BINS + 1), st);
  if (e != cudaSuccess) return e;
  asm_flow_bin_kernel<<<blocks(A.n_graphs, 256), 256, 0, st>>>(A, bins);
  asm_flow_bin_scan_kernel<<<1, 32, 0, st>>>(bins);
  asm_flow_order_kernel<<<blocks(A.n_graphs, 256), 256, 0, st>>>(A, bins, order);
  if ((e = cudaMemsetAsync(bins + FLOW_BINS, 0, sizeof(uint32_t), st)) != cudaSuccess) return e;
  const int64_t want = blocks(A.n_graphs * 32, 128), resident = (int64_t)num_sms * 5;   // 96 registers: 5 CTAs of 4 warps per SM
  asm_flow_kernel<<<(unsigned)(want < resident ? want : resident), 128, 0, st>>>(v, A, P, order, bins + FLOW_BINS);
  return cudaGetLastError();
}
size_t asm_flow_bins_words() { return FLOW_BINS + 1; }
void asm_set_debug(int on) { cudaMemcpyToSymbol(g_asm_debug, &on, sizeof(int)); }
cudaError_t launch_asm_geneno(const DevBatchView& v, const AsmView& A, const int32_t* b_geneno, cudaStream_t st) {
  asm_geneno_kernel<<<blocks(v.n_bundles + 1, 256), 256, 0, st>>>(v, A, b_geneno);
  return cudaGetLastError();
}
cudaError_t launch_asm_pred_scatter(const AsmView& A, uint32_t npool, cudaStream_t st) {
  asm_pred_scatter_kernel<<<blocks((int64_t)npool + 1, 256), 256, 0, st>>>(A, npool);
  return cudaGetLastError();
}
cudaError_t launch_asm_exon_scatter(const AsmView& A, uint32_t npool, cudaStream_t st) {
  if (npool) asm_exon_scatter_kernel<<<blocks(npool, 256), 256, 0, st>>>(A, npool);
  return cudaGetLastError();
}
size_t asm_graph_struct_bytes() { return sizeof(Graph); }
