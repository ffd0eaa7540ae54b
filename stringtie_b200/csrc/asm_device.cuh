// asm_device.cuh — SURVEY.md §8 rows a10–a15 for ONE splice graph, run by ONE warp: create_graph, the read -> transfrag
// bookkeeping a graph needs, process_transfrags, find_transcripts / parse_trf, push_max_flow and store_transcript.
//
// Reference (CPU, sequential per graph; short-read de novo path: no guides, -L, --mix, -N, --merge, point features):
//   create_graph            rlink.cpp:3579-4351   trimnode_all :2809   traverse_dfs :2874 (transfrags it leaves, reachability)
//   process_transfrags      rlink.cpp:5768-6910   eliminate_transfrags_under_thr :5224   trf_real :5351   trCmp :1129
//   find_transcripts        rlink.cpp:13633-13810 parse_trf :11992   back_to_source_fast :8651   fwd_to_sink_fast :8483
//   onpath :7863   push_max_flow :9075   store_transcript :9688
//
// What is NOT carried over from the reference's representation:
//   * GBitVec patterns over gno + edgeno bits with a hash (gpos) from node pairs to edge bits.  Here a transfrag is its
//     node list plus one flag per consecutive pair ("the edge between them is part of the pattern"), a node-bit pattern
//     for subset tests, and a path is its node list plus succ[] (edge (u, v) is on the path iff succ[u] == v).  Every
//     test the reference makes on edge bits is a test on consecutive nodes of a transfrag or of the path, so the two are
//     equivalent; pattern.count() (trCmp) is nodes + flagged pairs.
//   * childpat / parentpat per node: one reachability bit matrix (row u: nodes reachable from u).  parentpat's edge
//     bits are only ever used at rlink.cpp:5071, where they are implied by the node bits (an edge (u, v) of the graph is
//     in parentpat[p] iff v is p or an ancestor of p).
//   * CTreePat: a hash table keyed by (graph, node list, flags) (frag_device.cuh).
// Node ids are topological (edges go from lower to higher ids): nodes are created left to right.
//
// Execution model (warp_prims.cuh): warp-uniform control flow; loops over bases, transfrags of a node and pattern
// words are strided over the lanes; order-dependent float sums are taken in index order.  Sections that are inherently
// sequential and short (the event sweep, the DFS, the reference's quicksort, the greedy flow subtraction) run on
// warp-uniform scalars with lane 0 writing.  Compiled for the host by tools/asm_hostcheck.cu (one-lane warps).
#pragma once
#include "stgpu_internal.h"
#include "warp_prims.cuh"
#include "trims_device.cuh"

#define ASM_ERR_CAP 0x200u        // a capacity bound of the graph arrays was exceeded (a bug in the bound, not in the input)
#define ASM_ERR_PRUNE 0x400u      // a graph exceeded allowed_nodes: prune_graph_nodes (rlink.cpp:3270) is not built
#define ASM_EPSILON 0.000001      // rlink.h:25
#define ASM_TRTHR 1.0f            // rlink.h:26
#define ASM_ERROR_PERC 0.1        // rlink.h:14
#define ASM_DROP 0.5              // rlink.h:13
#define ASM_LIA 25u               // longintronanchor, rlink.h:32

// get_cov(1, a, b) on the bundle's lane 1 (rlink.cpp:1830); a, b bundle-relative, clamped to the array (the reference
// reads past it when a window sticks out of the bundle: undefined there, defined here)
STG_HD double asm_cov_all(const float* __restrict__ c, uint32_t covlen, uint32_t a, uint32_t b) {
  if (b + 1 >= covlen) b = covlen - 2;
  if (a > b + 1) a = b + 1;
  const int m = (int)((b + 1) / STG_BSIZE);
  const int k = (int)(a / STG_BSIZE);
  double cov = 0;
  for (int i = k + 1; i <= m; i++) cov += (double)c[(size_t)i * STG_BSIZE - 1];
  cov += (double)wp_fsub(c[b + 1], c[a]);
  if (cov < 0) cov = 0;
  return cov;
}

struct AsmParams {               // the options of stg_params these stages read (stringtie.cpp:115-192)
  uint32_t junctionsupport;      // -a
  int trim;                      // !-t
  float mcov;                    // -M
  int mintranscriptlen;          // -m
  int allowed_nodes;             // 1000
  float isofrac;                 // -f
  int includesource;             // !-z
};

// Strand-filtered views of a bundle's junction table (rows sorted by start, end): the rows of strand `strand` (+1 / -1)
// in table order, and the same rows ordered by end (ties keep table order).  create_graph walks both lists with two
// cursors and skips the rows of the other strand and the deleted ones (strand 0) wherever it meets them
// (rlink.cpp:3853-3875), so filtering up front is equivalent.  Returns the number of rows kept.
STG_HDN inline int junc_lists_warp(const uint32_t* __restrict__ j_start, const uint32_t* __restrict__ j_end, const int8_t* __restrict__ j_strand, int nj,
                                   int strand, uint32_t* js_start, uint32_t* js_end, uint32_t* je_end, int32_t* je_k) {
  const int lane = wp_lane();
  int n = 0;
  for (int j0 = 0; j0 < nj; j0 += WP_N) {
    const int j = j0 + lane;
    const bool keep = j < nj && (int)j_strand[j] == strand;
    const uint32_t m = wp_ballot(keep);
    if (keep) {
      const int k = n + wp_popc(m & ((1u << lane) - 1u));
      js_start[k] = j_start[j]; js_end[k] = j_end[j];
    }
    n += wp_popc(m);
  }
  wp_sync();
  // order by end: rank of row k = rows with a smaller (end, k)
  for (int k = lane; k < n; k += WP_N) {
    const uint32_t e = js_end[k];
    int rank = 0;
    for (int x = 0; x < n; x++) { const uint32_t ex = js_end[x]; rank += (ex < e || (ex == e && x < k)) ? 1 : 0; }
    je_end[rank] = e; je_k[rank] = k;
  }
  wp_sync();
  return n;
}

// ---- a10, first half: the event sweep of create_graph (rlink.cpp:3647-4128) --------------------------------------
struct GraphSrc {
  int s;                         // 0: '-' strand graph, 1: '+'
  uint32_t refstart, covlen;     // BundleData::start; entries of one bpcov lane
  const float *c_all, *c_opp;    // bpcov lane 1 and lane 2 - 2s
  const uint32_t *bn_start, *bn_end; const int32_t* bn_next;   // bundlenodes of strand lane 2s (lane-local ids)
  int startnode, lastnode;       // CBundle::startnode / lastnodeid
  const uint32_t *js_start, *js_end; int nJs;            // junctions of strand 2s-1 in table order (start, end)
  const uint32_t* je_end; const int32_t* je_k; int nJe;  // the same junctions by end (stable): end, index into js_*
  int32_t *reg_node, *reg_graph; // [nJs] `ends` (rlink.cpp:3625): node registered under junction k's end, and by which graph
  int gid;
  uint32_t junctionsupport; int trim;
};

struct SweepOut {                // temporary, bound-sized (exact sizes are known only after the sweep)
  uint32_t *n_start, *n_end;     // [ncap] node 0 = source
  int32_t *ev_u, *ev_v;          // [ecap] link events in order: (u, v) = child + parent link; v == -2: u becomes a parent of sink
  int32_t *ft_n1, *ft_n2; float* ft_ab;   // [fcap] futuretr triples (n2 == -1: link to sink)
  int32_t *bn_first, *bn_cnt;    // per bundlenode (lane-local id): its nodes are [first, first + cnt)
  uint32_t* tr_pos; float* tr_ab; uint8_t* tr_st; uint32_t tr_cap;   // scratch of find_trims_warp
  int ncap, ecap, fcap;
  int nn, ne, nf, edgeno;        // results
  uint32_t err;
};

STG_HD int sw_new_node(SweepOut& o, uint32_t start, uint32_t end, int bid) {
  const int id = o.nn;
  if (id >= o.ncap) { o.err |= ASM_ERR_CAP; return o.ncap - 1; }
  if (wp_lane() == 0) {
    o.n_start[id] = start; o.n_end[id] = end;
    if (o.bn_cnt[bid] == 0) o.bn_first[bid] = id;
    o.bn_cnt[bid]++;
  }
  o.nn++;
  return id;
}
STG_HD void sw_link(SweepOut& o, int u, int v) {
  if (o.ne >= o.ecap) { o.err |= ASM_ERR_CAP; return; }
  if (wp_lane() == 0) { o.ev_u[o.ne] = u; o.ev_v[o.ne] = v; }
  o.ne++;
}
STG_HD void sw_future(SweepOut& o, int n1, int n2, float ab) {
  if (o.nf >= o.fcap) { o.err |= ASM_ERR_CAP; return; }
  if (wp_lane() == 0) { o.ft_n1[o.nf] = n1; o.ft_n2[o.nf] = n2; o.ft_ab[o.nf] = ab; }
  o.nf++;
}

// trimnode_all, rlink.cpp:2809-2867: cut the open node [nstart, ...] at the trim points found up to newend
STG_HDN inline int sw_trim(const GraphSrc& G, SweepOut& o, int graphnode, uint32_t& nstart, uint32_t newend, int bid) {
  if (newend < nstart) return graphnode;
  uint32_t cap = (newend - nstart + 1) / 50u + 2u;
  if (cap > o.tr_cap) cap = o.tr_cap;
  wp_sync();
  const uint32_t nt = find_trims_warp(G.c_all, G.c_opp, G.refstart, nstart, newend, o.tr_pos, o.tr_ab, o.tr_st, cap);
  wp_sync();
  for (uint32_t i = 0; i < nt; i++) {
    const uint32_t pos = o.tr_pos[i];
    if (!pos) continue;
    const float ab = wp_fadd(o.tr_ab[i], ASM_TRTHR);
    const int prev = graphnode;
    if (o.tr_st[i]) {   // source trim
      if (wp_lane() == 0) o.n_end[prev] = pos - 1;
      graphnode = sw_new_node(o, pos, newend, bid);
      nstart = pos;
      sw_link(o, 0, graphnode);
      sw_link(o, prev, graphnode);
      sw_future(o, 0, graphnode, ab);
      sw_future(o, prev, graphnode, ASM_TRTHR);
    } else {            // sink trim
      if (wp_lane() == 0) o.n_end[prev] = pos;
      graphnode = sw_new_node(o, pos + 1, newend, bid);
      nstart = pos + 1;
      sw_link(o, prev, graphnode);
      sw_link(o, prev, -2);
      sw_future(o, prev, -1, ab);
      sw_future(o, prev, graphnode, ASM_TRTHR);
    }
    o.edgeno += 2;
  }
  wp_sync();
  return graphnode;
}

// link every node registered under `pos` by this graph to `graphnode` (the reference's ends[pos], rlink.cpp:3799, 4052);
// returns how many there were.  je_lo: some index into the end-ordered list at or before the first entry with end == pos
STG_HD int sw_link_ends(const GraphSrc& G, SweepOut& o, uint32_t pos, int graphnode, int je_hint) {
  int k = je_hint;
  while (k > 0 && G.je_end[k - 1] >= pos) k--;
  while (k < G.nJe && G.je_end[k] < pos) k++;
  int cnt = 0;
  for (; k < G.nJe && G.je_end[k] == pos; k++) {
    const int j = G.je_k[k];
    if (G.reg_graph[j] == G.gid && G.reg_node[j] >= 0) { sw_link(o, G.reg_node[j], graphnode); o.edgeno++; cnt++; }
  }
  return cnt;
}

// has this graph registered anything under `pos`?  (ends[pos] != NULL)
STG_HD bool sw_has_ends(const GraphSrc& G, uint32_t pos, int je_hint) {
  int k = je_hint;
  while (k > 0 && G.je_end[k - 1] >= pos) k--;
  while (k < G.nJe && G.je_end[k] < pos) k++;
  for (; k < G.nJe && G.je_end[k] == pos; k++) {
    const int j = G.je_k[k];
    if (G.reg_graph[j] == G.gid && G.reg_node[j] >= 0) return true;
  }
  return false;
}

STG_HDN inline void graph_sweep(const GraphSrc& G, SweepOut& o) {
  o.nn = 0; o.ne = 0; o.nf = 0; o.edgeno = 0; o.err = 0;
  if (wp_lane() == 0) { o.n_start[0] = 0; o.n_end[0] = 0; }
  o.nn = 1;   // the source
  const int lane = wp_lane();
  int njs = 0, nje = 0;
  const uint32_t bundle_end = G.bn_end[G.lastnode];
  {  // cursors start at the first junction that can matter (the reference scans from 0 and skips)
    const uint32_t first = G.bn_start[G.startnode];
    int lo = 0, hi = G.nJs;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (G.js_start[mid] < first) lo = mid + 1; else hi = mid; }
    njs = lo;
    lo = 0; hi = G.nJe;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (G.je_end[mid] < first) lo = mid + 1; else hi = mid; }
    nje = lo;
  }
  for (int bn = G.startnode; bn != -1 && !o.err; bn = G.bn_next[bn]) {
    uint32_t currentstart = G.bn_start[bn];
    const uint32_t endbundle = G.bn_end[bn];
    if (lane == 0) { o.bn_cnt[bn] = 0; o.bn_first[bn] = -1; }
    int end = 0;
    while (nje < G.nJe && G.je_end[nje] <= currentstart) { if (G.je_end[nje] == currentstart) end = 1; nje++; }
    if (!end) {   // a short hanging piece in front of a junction end whose coverage rises there: start the node at the junction end
      if (nje < G.nJe && G.je_end[nje] - currentstart < G.junctionsupport) {
        const uint32_t je = G.je_end[nje];
        const double covleft = asm_cov_all(G.c_all, G.covlen, currentstart - G.refstart, je - 1 - G.refstart);
        const double covright = asm_cov_all(G.c_all, G.covlen, je - G.refstart, 2 * je - currentstart - 1 - G.refstart);
        if (covleft < covright * (1 - ASM_ERROR_PERC)) {
          currentstart = je;
          while (nje < G.nJe && G.je_end[nje] <= currentstart) { if (G.je_end[nje] == currentstart) end = 1; nje++; }
        }
      }
    }
    int graphnode = sw_new_node(o, currentstart, endbundle, bn);
    uint32_t nstart = currentstart;      // start of the open node
    wp_sync();
    if (end && sw_has_ends(G, currentstart, nje)) sw_link_ends(G, o, currentstart, graphnode, nje);
    else { sw_link(o, 0, graphnode); o.edgeno++; }
    bool completed = false;
    do {
      wp_sync();
      while (njs < G.nJs && G.js_start[njs] < currentstart) njs++;
      int minjunction = -1;
      if ((nje < G.nJe && G.je_end[nje] <= endbundle) || (njs < G.nJs && G.js_start[njs] <= endbundle)) {
        if (njs < G.nJs && G.js_start[njs] <= endbundle && G.js_end[njs] > bundle_end) njs++;
        else if (nje < G.nJe) {
          if (njs < G.nJs) minjunction = G.js_start[njs] >= G.je_end[nje] ? 1 : 0;
          else minjunction = 1;
        } else minjunction = 0;
      }
      if (minjunction == 0) {          // a junction starts here: the open node ends at its first base
        const uint32_t pos = G.js_start[njs];
        if (G.trim) graphnode = sw_trim(G, o, graphnode, nstart, pos, bn);
        if (lane == 0) o.n_end[graphnode] = pos;
        while (njs < G.nJs && G.js_start[njs] == pos) {
          if (lane == 0) { G.reg_node[njs] = graphnode; G.reg_graph[njs] = G.gid; }
          njs++;
        }
        if (pos < endbundle) {
          if (endbundle - pos < G.junctionsupport) {
            if ((njs >= G.nJs || G.js_start[njs] > endbundle) && (nje >= G.nJe || G.je_end[nje] > endbundle)) {
              const double covleft = asm_cov_all(G.c_all, G.covlen, 2 * pos - endbundle + 1 - G.refstart, pos - G.refstart);
              const double covright = asm_cov_all(G.c_all, G.covlen, pos + 1 - G.refstart, endbundle - G.refstart);
              if (covright < covleft * (1 - ASM_ERROR_PERC)) completed = true;
            }
          }
          if (!completed) {
            const int nextnode = sw_new_node(o, pos + 1, endbundle, bn);
            sw_link(o, graphnode, nextnode); o.edgeno++;
            graphnode = nextnode; nstart = pos + 1;
          }
        } else completed = true;
      } else if (minjunction == 1) {   // a junction ends here: a node starts at this base
        const uint32_t pos = G.je_end[nje];
        while (nje < G.nJe && G.je_end[nje] == pos) nje++;
        if (nstart < pos) {
          if (G.trim) graphnode = sw_trim(G, o, graphnode, nstart, pos - 1, bn);
          if (lane == 0) o.n_end[graphnode] = pos - 1;
          const int nextnode = sw_new_node(o, pos, endbundle, bn);
          sw_link(o, graphnode, nextnode); o.edgeno++;
          graphnode = nextnode; nstart = pos;
        }
        wp_sync();
        sw_link_ends(G, o, pos, graphnode, nje);
      }
    } while (!o.err && ((nje < G.nJe && G.je_end[nje] <= endbundle) || (njs < G.nJs && G.js_start[njs] <= endbundle)));
    if (!completed) {
      if (G.trim) graphnode = sw_trim(G, o, graphnode, nstart, endbundle, bn);
      if (lane == 0) o.n_end[graphnode] = endbundle;
      o.edgeno++;
    }
  }
  wp_sync();
}

// ---- a10, second half: adjacency, coverage-drop links, futuretr, DFS, reachability --------------------------------
struct Graph {                   // one built graph, exact-size arrays (read-only after graph_finish)
  int gno;                       // nodes incl. source (0) and sink (gno - 1); 0: no graph
  int nwn;                       // 64-bit words of a node-bit pattern: ceil(gno / 64)
  uint32_t *n_start, *n_end;     // [gno]
  float* n_cov;                  // [gno] CGraphnode::cov (a11)
  int32_t *ch_off, *ch_cnt, *ch; // child lists: node u owns ch[ch_off[u] .. ch_off[u] + ch_cnt[u])
  int32_t *pa_off, *pa_cnt, *pa; // parent lists (pa_off already points at the first parent)
  uint8_t* n_sinkpar;            // [gno] node is a parent of sink (CGraphnode::parent of the sink, as a set)
  unsigned long long* reach;     // [gno * nwn] row u: nodes reachable from u (childpat's node bits; column v: parentpat)
  int32_t *tf_n1, *tf_n2; float* tf_ab; int ntf0;   // transfrags create_graph leaves (two nodes, edge flagged)
  int edgeno;                    // the reference's edge count (an overestimate; only reported)
};

STG_HD bool g_reach(const Graph& g, int u, int v) { return (g.reach[(size_t)u * g.nwn + (v >> 6)] >> (v & 63)) & 1ull; }
STG_HD bool g_is_child(const Graph& g, int u, int v) {   // is (u, v) an edge of the graph (gpos[edge(u, v)] != NULL)
  const int32_t* c = g.ch + g.ch_off[u];
  const int n = g.ch_cnt[u];
  for (int i = 0; i < n; i++) if (c[i] == v) return true;
  return false;
}
STG_HD uint32_t g_len(const Graph& g, int i) { return g.n_end[i] - g.n_start[i] + 1u; }

// Capacities the caller must provide in g: ch: ne + 2*nn + 2 entries, pa: ne + 2*nn + 2, tf: nf + 2*nn + 2, reach:
// (nn+1) * ceil((nn+1)/64).  stack: [2*(nn+1)] ints of scratch, ncs: [nn+1] doubles of scratch.
STG_HDN inline void graph_finish(const GraphSrc& G, const SweepOut& o, Graph& g, int32_t* stack, double* ncs, uint8_t* visit,
                                 int allowed_nodes, uint32_t* err) {
  const int lane = wp_lane();
  const int nn = o.nn;            // nodes so far (source + real nodes); sink becomes node nn
  const int gno = nn + 1;
  g.gno = gno; g.nwn = (gno + 63) >> 6;
  // mean signed coverage of every node (the drop pass below reads them many times): lanes over nodes
  for (int i = 1 + lane; i < nn; i += WP_N) {
    const uint32_t a = o.n_start[i], z = o.n_end[i];
    ncs[i] = trims_cov_sign(G.c_all, G.c_opp, a - G.refstart, z - G.refstart) / (double)(z - a + 1u);
  }
  wp_sync();
  if (lane == 0) {
    for (int i = 0; i < nn; i++) { g.n_start[i] = o.n_start[i]; g.n_end[i] = o.n_end[i]; g.n_cov[i] = 0.f; g.n_sinkpar[i] = 0; }
    g.n_start[nn] = 0; g.n_end[nn] = 0; g.n_cov[nn] = 0.f; g.n_sinkpar[nn] = 0;   // CGraphnode() default: start = end = 0
    // adjacency from the link events, in event order; slack: one child slot per node (sink), nn for the source (drop
    // pass), one parent slot in FRONT of every list (source inserted at position 0, rlink.cpp:4154)
    for (int i = 0; i <= nn; i++) { g.ch_cnt[i] = 0; g.pa_cnt[i] = 0; }
    for (int e = 0; e < o.ne; e++) if (o.ev_v[e] >= 0) { g.ch_cnt[o.ev_u[e]]++; g.pa_cnt[o.ev_v[e]]++; }
    int co = 0, po = 0;
    for (int i = 0; i <= nn; i++) {
      g.ch_off[i] = co; co += g.ch_cnt[i] + 1 + (i == 0 ? nn : 0);
      g.pa_off[i] = po + 1; po += g.pa_cnt[i] + 1;
      g.ch_cnt[i] = 0; g.pa_cnt[i] = 0;
    }
    for (int e = 0; e < o.ne; e++) {
      const int u = o.ev_u[e], v = o.ev_v[e];
      if (v >= 0) { g.ch[g.ch_off[u] + g.ch_cnt[u]++] = v; g.pa[g.pa_off[v] + g.pa_cnt[v]++] = u; }
      else g.n_sinkpar[u] = 1;
    }
    int edgeno = o.edgeno;
    int nf = o.nf;
    int32_t *ft_n1 = o.ft_n1, *ft_n2 = o.ft_n2; float* ft_ab = o.ft_ab;
    // source / sink links for very high coverage drops, rlink.cpp:4133-4207
    for (int i = 1; i < nn; i++) {
      double icov = 0;
      if (i > 1 && g.pa[g.pa_off[i]]) {
        icov = ncs[i];
        double parcov = 0;
        bool addsource = true;
        for (int j = 0; j < g.pa_cnt[i]; j++) {
          const int p = g.pa[g.pa_off[i] + j];
          if (!p) { addsource = false; break; }
          parcov += ncs[p];
        }
        if (addsource && parcov < icov * ASM_ERROR_PERC * ASM_DROP) {
          g.pa_off[i]--; g.pa[g.pa_off[i]] = 0; g.pa_cnt[i]++;
          g.ch[g.ch_off[0] + g.ch_cnt[0]++] = i;
          if (nf < o.fcap) { ft_n1[nf] = 0; ft_n2[nf] = i; ft_ab[nf] = (float)((icov - parcov) / ASM_DROP); nf++; } else *err |= ASM_ERR_CAP;
          edgeno++;
        }
      }
      if (i >= nn - 1) continue;   // (the reference tests i < graphno - 1 before the sink has been added)
      if (g.ch_cnt[i]) {
        if (!g.n_sinkpar[i]) {
          if (!icov) icov = ncs[i];
          double chcov = 0;
          for (int j = 0; j < g.ch_cnt[i]; j++) chcov += ncs[g.ch[g.ch_off[i] + j]];
          if (chcov < icov * ASM_ERROR_PERC * ASM_DROP) {
            g.n_sinkpar[i] = 1;
            if (nf < o.fcap) { ft_n1[nf] = i; ft_n2[nf] = -1; ft_ab[nf] = (float)((icov - chcov) / ASM_DROP); nf++; } else *err |= ASM_ERR_CAP;
            edgeno++;
          }
        }
      } else edgeno++;
    }
    if (nn > allowed_nodes) *err |= ASM_ERR_PRUNE;
    const int sink = nn;
    // futuretr -> transfrags, rlink.cpp:4226-4316
    int ntf = 0;
    for (int f = 0; f < nf; f++) {
      const int n1 = ft_n1[f];
      if (n1 < 0) continue;
      int n2 = ft_n2[f];
      if (n2 < 0) {
        int n3 = n1 + 1;
        if (n3 < sink) {   // a sink link is dropped when a node right behind (within longintronanchor bases) already has one
          int prevnode = n1, nextnode = n3;
          uint32_t dist = g_len(g, nextnode) - 1;
          bool keeptr = true;
          while (g.n_start[nextnode] == g.n_end[prevnode] + 1 && dist < ASM_LIA) {
            for (int c = 0; c < g.ch_cnt[nextnode]; c++) if (g.ch[g.ch_off[nextnode] + c] == sink) { keeptr = false; break; }
            n3++;
            if (n3 == sink) break;
            prevnode = nextnode; nextnode = n3;
            dist += g_len(g, nextnode);
          }
          if (!keeptr) continue;
        }
        g.ch[g.ch_off[n1] + g.ch_cnt[n1]++] = sink;
        n2 = sink;
      }
      g.tf_n1[ntf] = n1; g.tf_n2[ntf] = n2; g.tf_ab[ntf] = ft_ab[f]; ntf++;
    }
    // traverse_dfs, rlink.cpp:2874-3027: the order in which it meets the nodes decides the order of the transfrags it adds
    for (int i = 0; i <= nn; i++) visit[i] = 0;
    int sp = 0;
    stack[0] = 0; stack[1] = 0; sp = 1;
    // first visit of the source
    visit[0] = 1;
    if (g.ch_cnt[0] == 0) {   // (an empty graph: source straight to sink)
      g.ch[g.ch_off[0] + g.ch_cnt[0]++] = sink; g.n_sinkpar[0] = 1;
      g.tf_n1[ntf] = 0; g.tf_n2[ntf] = sink; g.tf_ab[ntf] = ASM_TRTHR; ntf++;
    }
    while (sp > 0) {
      const int u = stack[2 * (sp - 1)];
      const int ci = stack[2 * (sp - 1) + 1];
      if (ci >= g.ch_cnt[u]) { sp--; continue; }
      stack[2 * (sp - 1) + 1] = ci + 1;
      const int v = g.ch[g.ch_off[u] + ci];
      if (visit[v]) continue;
      visit[v] = 1;
      if (g.pa_cnt[v] == 1 && g.pa[g.pa_off[v]] == 0) { g.tf_n1[ntf] = 0; g.tf_n2[ntf] = v; g.tf_ab[ntf] = ASM_TRTHR; ntf++; }
      if (v != sink && g.ch_cnt[v] == 0) {
        g.ch[g.ch_off[v] + g.ch_cnt[v]++] = sink; g.n_sinkpar[v] = 1;
        g.tf_n1[ntf] = v; g.tf_n2[ntf] = sink; g.tf_ab[ntf] = ASM_TRTHR; ntf++;
      }
      stack[2 * sp] = v; stack[2 * sp + 1] = 0; sp++;
    }
    g.ntf0 = ntf;
    g.edgeno = edgeno;
  }
  wp_sync();
  // reachability: row u = OR over the children c of (row c | bit c); children have higher ids, so go down from the sink
  const int nwn = g.nwn;
  for (int u = gno - 1; u >= 0; u--) {
    const int nc = g.ch_cnt[u];
    for (int w = lane; w < nwn; w += WP_N) {
      unsigned long long acc = 0;
      for (int c = 0; c < nc; c++) {
        const int v = g.ch[g.ch_off[u] + c];
        acc |= g.reach[(size_t)v * nwn + w];
        if ((v >> 6) == w) acc |= 1ull << (v & 63);
      }
      g.reach[(size_t)u * nwn + w] = acc;
    }
    wp_sync();
  }
}

// ---- a12: process_transfrags (rlink.cpp:5768-6910), short-read de novo path --------------------------------------
// a key node: the node id, bit 31 = "the edge from the previous node of the key is part of the pattern"
#define TF_FLAG 0x80000000u
#define TF_MAXTRF 40000          // max_trf_number, rlink.h:41

struct TfG {                     // the transfrags of one graph
  int nt, cap;                   // in use / capacity (transfrags after a11 + one per edge of the graph)
  uint32_t *koff, *klen;         // [cap] key of transfrag t: key[koff[t] .. koff[t] + klen[t])
  uint32_t* key;                 // key pool of the batch (a11's record keys; the edge transfrags added here get new storage)
  uint32_t xkey0;                // first free entry of this graph's extra key storage (2 per edge of the graph)
  float* ab;                     // [cap] CTransfrag::abundance
  uint8_t* real;                 // [cap]
  int32_t* usepath;              // [cap]
  int32_t *path_off, *path_cnt;  // [cap] CTransfrag::path: entries path_*[path_off[t] .. + path_cnt[t])
  int32_t *path_node, *path_cont; float* path_ab; int path_cap, path_used;
  unsigned long long* pat;       // [cap * nwn] node bits of the pattern
  int32_t *ntrf_off, *ntrf_cnt, *ntrf; int ntrf_cap;   // CGraphnode::trf of node i: ntrf[ntrf_off[i] .. + ntrf_cnt[i])
  // working arrays of the stage
  int32_t* order;                // [cap] the list being edited: positions -> transfrag slots
  uint32_t *w_koff, *w_klen; float* w_ab;   // [cap] slots before the sort
  unsigned long long* skey;      // [cap] (reach, nodes, pattern bits) packed for trCmp
  uint8_t* e_cov;                // [edges] the edge is part of some transfrag's pattern (allpat)
};

STG_HD int tf_node(const TfG& f, int t, int i) { return (int)(f.key[f.koff[t] + (uint32_t)i] & ~TF_FLAG); }
STG_HD bool tf_flag(const TfG& f, int t, int i) { return (f.key[f.koff[t] + (uint32_t)i] & TF_FLAG) != 0; }
STG_HD int tf_first(const TfG& f, int t) { return (int)(f.key[f.koff[t]] & ~TF_FLAG); }
STG_HD int tf_last(const TfG& f, int t) { return (int)(f.key[f.koff[t] + f.klen[t] - 1] & ~TF_FLAG); }
STG_HD bool tf_has(const TfG& f, int nwn, int t, int node) { return (f.pat[(size_t)t * nwn + (node >> 6)] >> (node & 63)) & 1ull; }
// does transfrag t carry the edge u -> v in its pattern (pattern[*gpos[edge(u, v)]])
STG_HD bool tf_has_edge(const TfG& f, int nwn, int t, int u, int v) {
  if (!tf_has(f, nwn, t, u) || !tf_has(f, nwn, t, v)) return false;
  const uint32_t* k = f.key + f.koff[t];
  const int n = (int)f.klen[t];
  for (int i = 1; i < n; i++) if ((int)(k[i] & ~TF_FLAG) == v) return (k[i] & TF_FLAG) && (int)(k[i - 1] & ~TF_FLAG) == u;
  return false;
}
STG_HD int g_edge_index(const Graph& g, int u, int v) {   // position of (u, v) in the child array, -1: not an edge
  const int n = g.ch_cnt[u];
  for (int i = 0; i < n; i++) if (g.ch[g.ch_off[u] + i] == v) return g.ch_off[u] + i;
  return -1;
}

// trCmp, rlink.cpp:1129-1141, on slots a, b of the working arrays
STG_HD int tf_cmp(const TfG& f, int a, int b) {
  if (f.skey[a] < f.skey[b]) return 1;
  if (f.skey[a] > f.skey[b]) return -1;
  if (f.w_ab[a] < f.w_ab[b]) return 1;
  if (f.w_ab[a] > f.w_ab[b]) return -1;
  return 0;
}

// GPVec::qSort (gclib/GVec.hh:896-920): Hoare partition around the middle element, not stable; comparators return 0 on
// ties, so the permutation it leaves is part of the result (SURVEY.md §7 hard part 3).  The two halves are independent,
// so an explicit stack (smaller half first) gives the same permutation as the reference's recursion.
// The sort keys travel WITH the list (klo / khi / kab [position], swapped together with it) instead of being looked up
// through the slot: the two scans of a partition then read consecutive words (one cache line per 32 steps) where the
// indirect form paid two dependent loads from L2 per comparison — the sort of the largest graph was the tail of asm_flow.
// The three key arrays are koff / klen / ab of the final list, which is only written after the sort.
STG_HDN inline void tf_sort_ref(const TfG& f, int32_t* list, int n) {
  uint32_t* klo = f.koff; uint32_t* khi = f.klen; float* kab = f.ab;
  for (int t = 0; t < n; t++) {
    const int sl = list[t];
    klo[t] = (uint32_t)f.skey[sl]; khi[t] = (uint32_t)(f.skey[sl] >> 32); kab[t] = f.w_ab[sl];
  }
  int stack[2 * 48];
  int sp = 0;
  if (n < 2) return;
  stack[sp++] = 0; stack[sp++] = n - 1;
  while (sp) {
    const int R = stack[--sp], L = stack[--sp];
    int I = L, J = R;
    const int M = L + ((R - L) >> 1);
    const unsigned long long pk = ((unsigned long long)khi[M] << 32) | klo[M];   // the pivot ELEMENT's keys (trCmp: larger keys first)
    const float pa = kab[M];
    do {
      for (;;) {    // while (trCmp(list[I], P) < 0) I++
        const unsigned long long k = ((unsigned long long)khi[I] << 32) | klo[I];
        if (k > pk || (k == pk && kab[I] > pa)) I++; else break;
      }
      for (;;) {    // while (trCmp(list[J], P) > 0) J--
        const unsigned long long k = ((unsigned long long)khi[J] << 32) | klo[J];
        if (k < pk || (k == pk && kab[J] < pa)) J--; else break;
      }
      if (I <= J) {
        const int t = list[I]; list[I] = list[J]; list[J] = t;
        const uint32_t a = klo[I]; klo[I] = klo[J]; klo[J] = a;
        const uint32_t b = khi[I]; khi[I] = khi[J]; khi[J] = b;
        const float c = kab[I]; kab[I] = kab[J]; kab[J] = c;
        I++; J--;
      }
    } while (I <= J);
    const bool left = L < J, right = I < R;
    if (left && right) {
      if (J - L > R - I) { stack[sp++] = L; stack[sp++] = J; stack[sp++] = I; stack[sp++] = R; }
      else { stack[sp++] = I; stack[sp++] = R; stack[sp++] = L; stack[sp++] = J; }
    } else if (left) { stack[sp++] = L; stack[sp++] = J; }
    else if (right) { stack[sp++] = I; stack[sp++] = R; }
  }
}

// trf_real, rlink.cpp:5351-5410: does the incomplete transfrag t have only one way through the graph?  Otherwise its
// abundance is shared out among the branches leaving the node where it first breaks (CTransfrag::path, as proportions).
STG_HDN inline bool tf_trf_real(const Graph& g, TfG& f, int t) {
  const int nwn = g.nwn;
  int i = 0;
  const int nt = (int)f.klen[t] - 1;
  int n = tf_node(f, t, 0);
  double totalcov = 0;
  const int po = f.path_off[t];
  int np_stored = 0;
  while (i < nt) {
    const int nx = tf_node(f, t, i + 1);
    if (n + 1 == nx) { n = nx; i++; continue; }
    if (g_is_child(g, n, nx) && tf_flag(f, t, i + 1) ) {
      // (an edge of the pattern between consecutive transfrag nodes: flag set means the graph edge n -> nx is in it.  n can
      // differ from node i of the transfrag after a detour below; the reference then asks for pattern[edge(n, nx)], which
      // is only set when n is the transfrag's own previous node)
      if (n == tf_node(f, t, i)) { n = nx; i++; continue; }
    }
    // a potential discontinuity: which children of n lead to nx, and how much abundance leaves through each
    const int nextnode = nx;
    const int ntrf = f.ntrf_cnt[n];
    int np = 0;
    totalcov = 0;
    np_stored = 0;
    for (int j = 0; j < g.ch_cnt[n]; j++) {
      const int c = g.ch[g.ch_off[n] + j];
      if (c > nextnode) break;
      if (c == nextnode || g_reach(g, c, nextnode)) {
        float pab = 0.f;
        for (int x = 0; x < ntrf; x++) {
          const int tt = f.ntrf[f.ntrf_off[n] + x];
          if (tf_has_edge(f, nwn, tt, n, c)) { pab = wp_fadd(pab, f.ab[tt]); totalcov += (double)f.ab[tt]; }
        }
        if (pab != 0.f) {
          f.path_node[po + np_stored] = n; f.path_cont[po + np_stored] = c; f.path_ab[po + np_stored] = pab;
          np_stored++; np++;
        }
      }
    }
    if (!(totalcov != 0) || !np) { f.path_cnt[t] = np_stored; return true; }
    if (np == 1) {
      n = f.path_cont[po];
      if (n == nextnode) i++;
      np_stored = 0;
    } else break;
  }
  f.path_cnt[t] = np_stored;
  if (i == nt) return true;
  for (int j = 0; j < np_stored; j++) f.path_ab[po + j] = (float)((double)f.path_ab[po + j] / totalcov);
  return false;
}

// Input: f.nt transfrags in slots 0..nt-1 of w_koff / w_klen / w_ab, in order of creation (create_graph's, then the
// reads' in order of first appearance).  Output: the final list in koff / klen / ab / real / pat, the node incidence
// lists, the CPath proportions of the incomplete transfrags; returns the number of transfrags.
STG_HDN inline int process_transfrags_warp(const Graph& g, TfG& f, uint32_t* err) {
  const int lane = wp_lane();
  const int gno = g.gno, nwn = g.nwn, sink = gno - 1;
  int nt = f.nt;
  // eliminate_transfrags_under_thr, rlink.cpp:5224-5262: the list is edited from the back (swap with the last, pop)
  if (lane == 0) {
    for (int t = 0; t < nt; t++) f.order[t] = t;
    float threshold = ASM_TRTHR;
    for (;;) {
      for (int t = nt - 1; t >= 0; t--) {
        const int sl = f.order[t];
        const uint32_t* k = f.key + f.w_koff[sl];
        const int first = (int)(k[0] & ~TF_FLAG), last = (int)(k[f.w_klen[sl] - 1] & ~TF_FLAG);
        if (f.w_ab[sl] < threshold && first && last < sink) { f.order[t] = f.order[nt - 1]; f.order[nt - 1] = sl; nt--; }
      }
      if (nt <= TF_MAXTRF) break;
      threshold = threshold + 1.f;
    }
  }
  nt = wp_shfl(nt, 0);
  wp_sync();
  // which edges of the graph are part of some pattern (allpat, rlink.cpp:6721)
  int nedges = g.ch_off[sink] + g.ch_cnt[sink];   // child array extent (slack entries are never marked)
  for (int e = lane; e < nedges; e += WP_N) f.e_cov[e] = 0;
  wp_sync();
  for (int t = lane; t < nt; t += WP_N) {
    const int sl = f.order[t];
    const uint32_t* k = f.key + f.w_koff[sl];
    const int n = (int)f.w_klen[sl];
    for (int i = 1; i < n; i++) if (k[i] & TF_FLAG) {
      const int e = g_edge_index(g, (int)(k[i - 1] & ~TF_FLAG), (int)(k[i] & ~TF_FLAG));
      if (e >= 0) f.e_cov[e] = 1;
    }
  }
  wp_sync();
  // a transfrag of abundance trthr for every edge no pattern covers, rlink.cpp:6745-6761
  if (lane == 0) {
    uint32_t xk = f.xkey0;
    for (int i = 1; i < gno - 1; i++) {
      for (int c = 0; c < g.ch_cnt[i]; c++) {
        const int e = g.ch_off[i] + c;
        if (f.e_cov[e]) continue;
        if (nt >= f.cap) { *err |= ASM_ERR_CAP; break; }
        const int sl = f.nt++;         // a fresh slot behind the ones a11 filled
        f.key[xk] = (uint32_t)i; f.key[xk + 1] = (uint32_t)g.ch[e] | TF_FLAG;
        f.w_koff[sl] = xk; f.w_klen[sl] = 2; f.w_ab[sl] = ASM_TRTHR;
        xk += 2;
        f.order[nt++] = sl;
      }
    }
    // sort keys of trCmp
    for (int t = 0; t < nt; t++) {
      const int sl = f.order[t];
      const uint32_t* k = f.key + f.w_koff[sl];
      const int n = (int)f.w_klen[sl];
      int pc = n;
      for (int i = 1; i < n; i++) pc += (k[i] & TF_FLAG) ? 1 : 0;
      const unsigned long long reach = (unsigned long long)((k[n - 1] & ~TF_FLAG) - (k[0] & ~TF_FLAG));
      f.skey[sl] = (reach << 42) | ((unsigned long long)n << 21) | (unsigned long long)pc;
    }
    tf_sort_ref(f, f.order, nt);
  }
  nt = wp_shfl(nt, 0);
  wp_sync();
  // the final list: lanes over transfrags
  for (int t = lane; t < nt; t += WP_N) {
    const int sl = f.order[t];
    f.koff[t] = f.w_koff[sl]; f.klen[t] = f.w_klen[sl]; f.ab[t] = f.w_ab[sl]; f.real[t] = 0; f.usepath[t] = -1;
    f.path_off[t] = 0; f.path_cnt[t] = 0;
    unsigned long long* p = f.pat + (size_t)t * nwn;
    for (int w = 0; w < nwn; w++) p[w] = 0ull;
    const uint32_t* k = f.key + f.koff[t];
    for (uint32_t i = 0; i < f.klen[t]; i++) { const uint32_t nd = k[i] & ~TF_FLAG; p[nd >> 6] |= 1ull << (nd & 63); }
  }
  wp_sync();
  // node incidence lists (rlink.cpp:6790-6862) — a transfrag is listed at each of its nodes and, where two consecutive
  // nodes are not joined by an edge of its pattern, at every node in between that lies on a way from the one to the other
  // (assign_incomplete_trf_to_nodes :5323).  Lists are ascending in t.  Pass 1 counts, pass 2 fills.
  int gmax = 1;
  for (int i = 1; i < gno; i++) if (g.ch_cnt[i] > gmax) gmax = g.ch_cnt[i];
  if (lane == 0) {
    for (int i = 0; i < gno; i++) f.ntrf_cnt[i] = 0;
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1) {
        int o = 0;
        for (int i = 0; i < gno; i++) { f.ntrf_off[i] = o; o += f.ntrf_cnt[i]; f.ntrf_cnt[i] = 0; }
        if (o > f.ntrf_cap) { *err |= ASM_ERR_CAP; break; }
      }
      int pused = 0;
      for (int t = 0; t < nt; t++) {
        const int n1 = (int)f.klen[t];
        if (n1 < 2) continue;
        bool incomplete = false;
        for (int n = 0; n < n1; n++) {
          const int nd = tf_node(f, t, n);
          if (pass) f.ntrf[f.ntrf_off[nd] + f.ntrf_cnt[nd]] = t;
          f.ntrf_cnt[nd]++;
          if (n) {
            const int pv = tf_node(f, t, n - 1);
            if (pv && nd < sink && !(tf_flag(f, t, n) && g_is_child(g, pv, nd))) {
              for (int x = pv + 1; x < nd; x++) if (g_reach(g, pv, x) && g_reach(g, x, nd)) {
                if (pass) f.ntrf[f.ntrf_off[x] + f.ntrf_cnt[x]] = t;
                f.ntrf_cnt[x]++;
                incomplete = true;
              }
            }
          }
        }
        if (pass) {
          if (incomplete) {   // CPath storage: at most one entry per child of the node where it breaks
            if (pused + gmax > f.path_cap) { *err |= ASM_ERR_CAP; break; }
            f.path_off[t] = pused; pused += gmax;
          } else f.real[t] = 1;
        }
      }
      if (pass) f.path_used = pused;
    }
  }
  wp_sync();
  // trf_real for the incomplete ones: independent of each other (they read abundances, they write their own path)
  for (int t = lane; t < nt; t += WP_N) if (!f.real[t] && f.klen[t] > 1) f.real[t] = tf_trf_real(g, f, t) ? 1 : 0;
  wp_sync();
  // the transfrag from the source to a node that has no other parent takes the abundance of everything that leaves the
  // node (rlink.cpp:6876-6890); the transfrag from a node to the sink takes what the node's coverage leaves unexplained
  // (:6893-6908).  One node per lane: each touches only its own source / sink transfrag.
  for (int ci = lane; ci < g.ch_cnt[0]; ci += WP_N) {
    const int c = g.ch[g.ch_off[0] + ci];
    if (!(g.pa_cnt[c] == 1 && g.pa[g.pa_off[c]] == 0)) continue;
    double abundance = 0;
    int t0 = -1;
    for (int j = 0; j < f.ntrf_cnt[c]; j++) {
      const int t = f.ntrf[f.ntrf_off[c] + j];
      if (tf_last(f, t) == c) t0 = t; else abundance += (double)f.ab[t];
    }
    if (t0 > -1 && f.ab[t0] != 0.f) f.ab[t0] = (float)abundance;
  }
  wp_sync();
  for (int u = 1 + lane; u < sink; u += WP_N) {
    if (!g.n_sinkpar[u]) continue;
    double abundance = 0;
    int t0 = -1;
    for (int j = 0; j < f.ntrf_cnt[u]; j++) {
      const int t = f.ntrf[f.ntrf_off[u] + j];
      if (tf_last(f, t) == sink) t0 = t; else abundance += (double)f.ab[t];
    }
    if (t0 > -1 && f.ab[t0] != 0.f && f.ab[t0] <= 1.f) {
      float a = (float)((double)(g.n_cov[u] / (float)g_len(g, u)) - abundance);
      if (a < 1.f) a = 1.f;
      f.ab[t0] = a;
    }
  }
  wp_sync();
  return nt;
}

// ---- a13–a15: find_transcripts / parse_trf (rlink.cpp:13633, 11992), back_to_source_fast :8651, fwd_to_sink_fast :8483,
// onpath :7863, push_max_flow :9075, store_transcript :9688 — short-read de novo path ----------------------------------
struct FlowScratch {             // per graph
  float* nodecov;                // [gno]
  int32_t* path; int npath;      // [gno + 1]
  int32_t* succ;                 // [gno] next node of the path (the path's edge bits): -1 none
  unsigned long long* pathpat;   // [nwn] node bits of the path
  unsigned long long* istr;      // [ceil(cap / 64)] istranscript
  uint8_t *usednode, *prev_node; // [gno]
  uint8_t* prev_edge;            // [edges] prevpath's edge bits
  int32_t* node2path;            // [gno]
  float *capl, *capr, *suml, *sumr, *nodeflux;   // [gno + 1], indexed by position on the path
  int dbg_iters;                 // paths tried (STGPU_ASM_DEBUG)
};

struct PredOut {                 // predictions of the whole batch, appended as they are stored (any order across graphs)
  uint32_t* counters;            // [0] predictions, [1] exons
  uint32_t cap_pred, cap_exon;
  int32_t *p_graph, *p_seq;      // graph and position within the graph's predictions
  uint32_t *p_start, *p_end; double* p_cov; int32_t* p_tlen; uint32_t *p_exon0, *p_nex;
  uint32_t *e_start, *e_end; float* e_cov;
};

STG_HD bool fl_onpath_bit(const FlowScratch& S, int node) { return (S.pathpat[node >> 6] >> (node & 63)) & 1ull; }
STG_HD void fl_set_bit(FlowScratch& S, int node, bool on) {
  if (on) S.pathpat[node >> 6] |= 1ull << (node & 63); else S.pathpat[node >> 6] &= ~(1ull << (node & 63));
}

// onpath, rlink.cpp:7863-7898.  The pattern's edge bit between consecutive transfrag nodes is the key flag; the path's
// edge bit for (u, v) is succ[u] == v.
STG_HD bool fl_onpath(const Graph& g, const TfG& f, const FlowScratch& S, int t, int mini, int maxi) {
  const uint32_t* k = f.key + f.koff[t];
  const int n = (int)f.klen[t];
  const int tr0 = (int)(k[0] & ~TF_FLAG), trl = (int)(k[n - 1] & ~TF_FLAG);
  if (tr0 < mini && !g_reach(g, tr0, mini)) return false;
  if (trl > maxi && !g_reach(g, maxi, trl)) return false;
  bool first = true;
  for (int i = 0; i < n; i++) {
    const int nd = (int)(k[i] & ~TF_FLAG);
    const bool fl = i && (k[i] & TF_FLAG);
    if (nd >= mini && nd <= maxi) {
      if (!fl_onpath_bit(S, nd)) return false;
      if (first) {
        first = false;
        if (i && nd > mini && fl) return false;
        if (i && !g_reach(g, (int)(k[i - 1] & ~TF_FLAG), mini)) return false;
      } else if (fl && S.succ[(int)(k[i - 1] & ~TF_FLAG)] != nd) return false;
    }
    if (nd > maxi) {
      if (i && (int)(k[i - 1] & ~TF_FLAG) < maxi && fl) return false;
      if (!g_reach(g, maxi, nd)) return false;
      break;
    }
  }
  return true;
}

// drop the used-up transfrags (abundance < epsilon) from the list of node v, keeping the order (the reference deletes
// them while it scans, rlink.cpp:8529-8532)
STG_HDN inline void fl_compact_trf(const TfG& f, int v) {
  const int lane = wp_lane();
  const int n = f.ntrf_cnt[v];
  int32_t* L = f.ntrf + f.ntrf_off[v];
  int out = 0;
  for (int j0 = 0; j0 < n; j0 += WP_N) {
    const int j = j0 + lane;
    const int t = j < n ? L[j] : -1;
    const bool keep = j < n && !((double)f.ab[t] < ASM_EPSILON);
    const uint32_t m = wp_ballot(keep);
    wp_sync();
    if (keep) L[out + wp_popc(m & ((1u << lane) - 1u))] = t;
    out += wp_popc(m);
    wp_sync();
  }
  if (lane == 0) f.ntrf_cnt[v] = out;
  wp_sync();
}

// sum, in list order, of the abundances of the transfrags of node v that span [lo, hi] and are compatible with the path
STG_HDN inline double fl_support(const Graph& g, const TfG& f, const FlowScratch& S, int v, int lo, int hi, int mini, int maxi) {
  const int lane = wp_lane();
  fl_compact_trf(f, v);
  const int n = f.ntrf_cnt[v];
  const int32_t* L = f.ntrf + f.ntrf_off[v];
  double cov = 0;
  for (int j0 = 0; j0 < n; j0 += WP_N) {
    const int j = j0 + lane;
    bool ok = false;
    float ab = 0.f;
    if (j < n) {
      const int t = L[j];
      ab = f.ab[t];
      ok = tf_first(f, t) <= lo && tf_last(f, t) >= hi && fl_onpath(g, f, S, t, mini, maxi);
    }
    uint32_t m = wp_ballot(ok);
    while (m) { const int l = wp_ffs(m) - 1; m &= m - 1; cov += (double)wp_shfl(ab, l); }
  }
  return cov;
}

// back_to_source_fast, rlink.cpp:8651-8807 (the recursion is a loop: one step per node)
STG_HDN inline bool fl_back_to_source(const Graph& g, const TfG& f, FlowScratch& S, int i) {
  const int lane = wp_lane();
  for (;;) {
    const int nparents = g.pa_cnt[i];
    if (!nparents) return true;
    int maxp = -1;
    double maxcov = 0;
    bool exclude = false;
    for (int p = 0; p < nparents; p++) {
      const int pp = g.pa[g.pa_off[i] + p];
      if (pp == i - 1 && i > 1 && g.n_start[i] == g.n_end[pp] + 1 && S.nodecov[i - 1] / (float)g_len(g, pp) < 1000.f &&
          (double)S.nodecov[i] * (ASM_DROP + ASM_ERROR_PERC) > (double)S.nodecov[i - 1]) {
        exclude = true;
      } else {
        if (lane == 0) { fl_set_bit(S, pp, true); S.succ[pp] = i; }
        wp_sync();
        const double parentcov = fl_support(g, f, S, pp, pp, i, pp, S.path[0]);
        if (parentcov > maxcov) { maxcov = parentcov; maxp = pp; }
        else if (parentcov == maxcov && maxp != -1) { if (S.nodecov[maxp] < S.nodecov[pp]) maxp = pp; }
        wp_sync();
        if (lane == 0) { fl_set_bit(S, pp, false); S.succ[pp] = -1; }
        wp_sync();
      }
    }
    if (maxp == -1) {
      if (exclude && S.nodecov[i - 1] != 0.f) {
        const int pp = i - 1;
        if (lane == 0) { fl_set_bit(S, pp, true); S.succ[pp] = i; }
        wp_sync();
        const double parentcov = fl_support(g, f, S, pp, pp, i, pp, S.path[0]);
        wp_sync();
        if (lane == 0) { fl_set_bit(S, pp, false); S.succ[pp] = -1; }
        wp_sync();
        if (parentcov != 0) maxp = pp; else return false;
      } else return false;
    }
    if (lane == 0) {
      if (maxp) S.path[S.npath] = maxp;
      fl_set_bit(S, maxp, true); S.succ[maxp] = i;
    }
    if (maxp) S.npath++;
    wp_sync();
    i = maxp;
  }
}

// fwd_to_sink_fast, rlink.cpp:8483-8649
STG_HDN inline bool fl_fwd_to_sink(const Graph& g, const TfG& f, FlowScratch& S, int i) {
  const int lane = wp_lane();
  const int gno = g.gno;
  for (;;) {
    const int nchildren = g.ch_cnt[i];
    if (!nchildren) return true;
    int maxc = -1;
    double maxcov = 0;
    bool exclude = false;
    for (int c = 0; c < nchildren; c++) {
      const int cc = g.ch[g.ch_off[i] + c];
      if (cc == i + 1 && i < gno - 2 && g.n_end[i] + 1 == g.n_start[cc] && S.nodecov[i + 1] / (float)g_len(g, cc) < 1000.f &&
          (double)S.nodecov[i] * (ASM_DROP + ASM_ERROR_PERC) > (double)S.nodecov[i + 1]) {
        exclude = true;
      } else {
        if (lane == 0) { fl_set_bit(S, cc, true); S.succ[i] = cc; }
        wp_sync();
        const double childcov = fl_support(g, f, S, cc, i, cc, S.path[0], cc);
        if (childcov > maxcov) { maxcov = childcov; maxc = cc; }
        else if (childcov == maxcov && maxc != -1) { if (S.nodecov[maxc] < S.nodecov[cc]) maxc = cc; }
        wp_sync();
        if (lane == 0) { fl_set_bit(S, cc, false); S.succ[i] = -1; }
        wp_sync();
      }
    }
    if (maxc == -1) {
      if (exclude && S.nodecov[i + 1] != 0.f) {
        const int cc = i + 1;
        if (lane == 0) { fl_set_bit(S, cc, true); S.succ[i] = cc; }
        wp_sync();
        const double childcov = fl_support(g, f, S, cc, i, cc, S.path[0], cc);
        wp_sync();
        if (lane == 0) { fl_set_bit(S, cc, false); S.succ[i] = -1; }
        wp_sync();
        if (childcov != 0) maxc = cc; else return false;
      } else return false;
    }
    if (lane == 0) { S.path[S.npath] = maxc; fl_set_bit(S, maxc, true); S.succ[i] = maxc; }
    S.npath++;
    wp_sync();
    i = maxc;
  }
}

STG_HD bool fl_istr(const FlowScratch& S, int t) { return (S.istr[t >> 6] >> (t & 63)) & 1ull; }

// is the pattern of transfrag t a subset of the path's (nodes and edges)?  (pathpat & pattern) == pattern, rlink.cpp:9125
STG_HD bool fl_subset(const TfG& f, const FlowScratch& S, int nwn, int t) {
  const unsigned long long* p = f.pat + (size_t)t * nwn;
  for (int w = 0; w < nwn; w++) if (p[w] & ~S.pathpat[w]) return false;
  const uint32_t* k = f.key + f.koff[t];
  const int n = (int)f.klen[t];
  for (int i = 1; i < n; i++) if ((k[i] & TF_FLAG) && S.succ[(int)(k[i - 1] & ~TF_FLAG)] != (int)(k[i] & ~TF_FLAG)) return false;
  return true;
}

// push_max_flow, rlink.cpp:9075-9374 (full == true on this path: the `full` clause never runs)
STG_HDN inline float fl_push_max_flow(const Graph& g, TfG& f, FlowScratch& S) {
  const int lane = wp_lane();
  const int gno = g.gno, nwn = g.nwn;
  const int n = S.npath;
  const int32_t* path = S.path;
  for (int i = lane; i < gno; i += WP_N) S.node2path[i] = -1;
  wp_sync();
  for (int i = lane; i < n; i += WP_N) {
    S.node2path[path[i]] = i;
    S.nodeflux[i] = 0.f; S.capl[i] = 0.f; S.capr[i] = 0.f; S.suml[i] = 0.f; S.sumr[i] = 0.f;
  }
  wp_sync();
  const int plast = path[n - 1];
  for (int i = 1; i < n - 1; i++) {
    const int pi = path[i];
    const int nt = f.ntrf_cnt[pi];
    const int32_t* L = f.ntrf + f.ntrf_off[pi];
    float capl = 0.f, capr = 0.f, suml = 0.f, sumr = 0.f;
    for (int j0 = 0; j0 < nt; j0 += WP_N) {
      const int j = j0 + lane;
      float v_sl = 0.f, v_sr = 0.f, v_cl = 0.f, v_cr = 0.f;
      int has = 0;   // bit 0: sumleft, 1: sumright, 2: capleft, 3: capright
      bool mark = false;
      int t = -1;
      if (j < nt) {
        t = L[j];
        const float ab = f.ab[t];
        if (ab != 0.f) {
          const int t0 = tf_first(f, t), tl = tf_last(f, t);
          bool keeptr = false;
          if (fl_istr(S, t)) keeptr = true;
          else if (!t0) { if (tl == path[1]) keeptr = true; }
          else if (tl == plast) { if (t0 == path[n - 2]) keeptr = true; }
          else if (t0 == pi && fl_subset(f, S, nwn, t)) keeptr = true;
          if (keeptr) {
            mark = true;
            if (i == 1 || t0 == pi) {
              if (!f.real[t]) {
                int up = -1;
                for (int p = 0; p < f.path_cnt[t]; p++) {
                  const int a = f.path_node[f.path_off[t] + p], b = f.path_cont[f.path_off[t] + p];
                  if (g_is_child(g, a, b) && S.succ[a] == b) { up = p; break; }
                }
                f.usepath[t] = up;
              }
            }
            const int up = f.usepath[t];
            const bool useful = f.real[t] || (up > -1 && up < f.path_cnt[t]);
            const float part = f.real[t] ? ab : (useful ? wp_fmul(ab, f.path_ab[f.path_off[t] + up]) : 0.f);
            if (t0 < pi) { has |= 1; v_sl = ab; if (useful) { has |= 4; v_cl = part; } }
            if (tl > pi) { has |= 2; v_sr = ab; if (useful) { has |= 8; v_cr = part; } }
          } else {
            if (pi > t0) { has |= 1; v_sl = ab; }
            if (pi < tl) { has |= 2; v_sr = ab; }
          }
        }
      }
      // istranscript bits of this step (distinct transfrags: one word may be shared by several lanes)
      uint32_t mm = wp_ballot(mark);
      while (mm) {
        const int l = wp_ffs(mm) - 1; mm &= mm - 1;
        const int tt = wp_shfl(t, l);
        if (lane == 0) S.istr[tt >> 6] |= 1ull << (tt & 63);
      }
      // the four float sums, in list order
      uint32_t hm = wp_ballot(has != 0);
      while (hm) {
        const int l = wp_ffs(hm) - 1; hm &= hm - 1;
        const int h = wp_shfl(has, l);
        const float a_sl = wp_shfl(v_sl, l), a_sr = wp_shfl(v_sr, l), a_cl = wp_shfl(v_cl, l), a_cr = wp_shfl(v_cr, l);
        if (h & 1) suml = wp_fadd(suml, a_sl);
        if (h & 4) capl = wp_fadd(capl, a_cl);
        if (h & 2) sumr = wp_fadd(sumr, a_sr);
        if (h & 8) capr = wp_fadd(capr, a_cr);
      }
      wp_sync();
    }
    if (lane == 0) { S.capl[i] = capl; S.capr[i] = capr; S.suml[i] = suml; S.sumr[i] = sumr; }
    wp_sync();
    if (!(capl != 0.f)) return 0.f;
    if (!(capr != 0.f)) return 0.f;
  }
  // the flow the path can carry (rlink.cpp:9260-9281), then what every transfrag gives up for it
  float result = 0.f;
  if (lane == 0) {
    float prevflow = S.capl[1];
    for (int i = 1; i < n - 1; i++) {
      const float percleft = prevflow / S.suml[i];
      float percright = S.capr[i] / S.sumr[i];
      if (percright > percleft) percright = percleft;
      prevflow = wp_fmul(percright, S.sumr[i]);
    }
    if (prevflow != 0.f) {
      for (int i = n - 2; i > 0; i--) {
        S.nodeflux[i] = prevflow / S.sumr[i];
        S.capr[i] = prevflow;
        prevflow = wp_fmul(S.nodeflux[i], S.suml[i]);
      }
      for (int i = 1; i < n - 1; i++) if (S.capr[i] != 0.f) {
        const int pi = path[i];
        const int nt = f.ntrf_cnt[pi];
        for (int j = 0; j < nt; j++) {
          const int t = f.ntrf[f.ntrf_off[pi] + j];
          if (!(fl_istr(S, t) && f.ab[t] != 0.f)) continue;
          double trabundance = (double)f.ab[t];
          if (!f.real[t]) {
            if (f.usepath[t] > -1 && f.usepath[t] < f.path_cnt[t]) trabundance = (double)wp_fmul(f.ab[t], f.path_ab[f.path_off[t] + f.usepath[t]]);
            else trabundance = 0;
          }
          if (trabundance != 0 && tf_first(f, t) == pi) {
            if ((double)S.capr[i] > trabundance) {
              S.capr[i] = (float)((double)S.capr[i] - trabundance);
              const int n2 = S.node2path[tf_last(f, t)];
              for (int k = i + 1; k < n2; k++) S.capr[k] = (float)((double)S.capr[k] - trabundance);
              f.ab[t] = (float)((double)f.ab[t] - trabundance);
              if ((double)f.ab[t] < ASM_EPSILON) f.ab[t] = 0.f;
              else if (!f.real[t]) {
                float* pa = f.path_ab + f.path_off[t];
                pa[f.usepath[t]] = 0.f;
                if (f.path_cnt[t] - 1 < 2) f.real[t] = 1;
                else {
                  int np = 0;
                  for (int p = 0; p < f.path_cnt[t]; p++) if (pa[f.usepath[t]] != 0.f) np++;
                  if (np < 2) f.real[t] = 1;
                }
              }
            } else {
              f.ab[t] = wp_fsub(f.ab[t], S.capr[i]);
              if ((double)f.ab[t] < ASM_EPSILON) f.ab[t] = 0.f;
              else if (!f.real[t]) {
                float* pa = f.path_ab + f.path_off[t];
                if ((double)wp_fsub(wp_fmul(pa[f.usepath[t]], f.ab[t]), S.capr[i]) < ASM_EPSILON) {
                  pa[f.usepath[t]] = 0.f;
                  if (f.path_cnt[t] - 1 < 2) f.real[t] = 1;
                  else {
                    int np = 0;
                    for (int p = 0; p < f.path_cnt[t]; p++) if (pa[f.usepath[t]] != 0.f) np++;
                    if (np < 2) f.real[t] = 1;
                  }
                }
              }
              const int n2 = S.node2path[tf_last(f, t)];
              for (int k = i + 1; k < n2; k++) S.capr[k] = wp_fsub(S.capr[k], S.capr[i]);
              S.capr[i] = 0.f;
              break;
            }
          }
        }
      }
      // the transfrag from the source
      const int p0 = path[0];
      for (int j = 0; j < f.ntrf_cnt[p0]; j++) {
        const int t = f.ntrf[f.ntrf_off[p0] + j];
        if (fl_istr(S, t) && f.ab[t] != 0.f) {
          f.ab[t] = wp_fsub(f.ab[t], prevflow);
          if ((double)f.ab[t] < ASM_EPSILON) f.ab[t] = 0.f;
          break;
        }
      }
      result = S.nodeflux[1];
    }
  }
  wp_sync();
  return wp_shfl(result, 0);
}

STG_HD uint32_t out_alloc(uint32_t* ctr, uint32_t n) {
#ifdef __CUDA_ARCH__
  return atomicAdd(ctr, n);
#else
  const uint32_t v = *ctr; *ctr += n; return v;
#endif
}

// store_transcript, rlink.cpp:9688-9866, without a guide: exons are runs of contiguous path nodes; every node gives up
// the share nodeflux of its coverage.  Lane 0 only.  Returns the transcript's coverage (0: nothing stored).
// ex_s / ex_e / ex_c: [gno] scratch for the exons.
STG_HDN inline double fl_store_transcript(const Graph& g, FlowScratch& S, const AsmParams& P, PredOut& out, int graph_id, int& nstored,
                                          bool& included, uint32_t* ex_s, uint32_t* ex_e, float* ex_c, uint32_t* err) {
  const int n = S.npath;
  const int32_t* path = S.path;
  double cov = 0, excov = 0;
  int len = 0;
  int prevnode = -1;
  int nex = 0;
  const int s = path[0] ? 0 : 1;
  bool firstex = true;
  for (int i = s; i < n - 1; i++) {
    const int nd = path[i];
    if (!S.prev_node[nd]) { included = false; S.prev_node[nd] = 1; }
    if (i) {
      const int e = g_edge_index(g, path[i - 1], nd);
      if (e >= 0 && !S.prev_edge[e]) { included = false; S.prev_edge[e] = 1; }
    }
    const uint32_t nstart = g.n_start[nd], nend = g.n_end[nd];
    const double usedcov = (double)wp_fmul(wp_fmul(S.nodecov[nd], S.nodeflux[i]), (float)(nend - nstart + 1u));
    S.nodecov[nd] = wp_fmul(S.nodecov[nd], wp_fsub(1.f, S.nodeflux[i]));
    if (S.nodecov[nd] < 0.f) S.nodecov[nd] = 0.f;
    if (prevnode < 0 || firstex || nstart > g.n_end[prevnode] + 1) {
      if (prevnode >= 0 && !firstex) {
        excov /= (double)(ex_e[nex - 1] - ex_s[nex - 1] + 1u);
        ex_c[nex - 1] = (float)excov;
        excov = 0;
      }
      ex_s[nex] = nstart; ex_e[nex] = nend; nex++;
      firstex = false;
    } else ex_e[nex - 1] = nend;
    len += (int)(nend - nstart + 1u);
    cov += usedcov;
    excov += usedcov;
    prevnode = nd;
  }
  if (prevnode >= 0) {
    excov /= (double)(ex_e[nex - 1] - ex_s[nex - 1] + 1u);
    ex_c[nex - 1] = (float)excov;
  }
  if (len) cov /= (double)len;
  if (cov != 0 && len >= P.mintranscriptlen) {
    const uint32_t p = out_alloc(out.counters, 1u);
    const uint32_t e0 = out_alloc(out.counters + 1, (uint32_t)nex);
    if (p >= out.cap_pred || e0 + (uint32_t)nex > out.cap_exon) { *err |= ASM_ERR_CAP; return 0; }
    out.p_graph[p] = graph_id; out.p_seq[p] = nstored++;
    out.p_start[p] = ex_s[0]; out.p_end[p] = ex_e[nex - 1]; out.p_cov[p] = cov; out.p_tlen[p] = len;
    out.p_exon0[p] = e0; out.p_nex[p] = (uint32_t)nex;
    for (int x = 0; x < nex; x++) { out.e_start[e0 + x] = ex_s[x]; out.e_end[e0 + x] = ex_e[x]; out.e_cov[e0 + x] = ex_c[x]; }
  } else cov = 0;
  return cov;
}

// find_transcripts (rlink.cpp:13633-13810) + the parse_trf recursion (:11992-12123) as a loop: seed = the node of highest
// coverage, extend to the source and to the sink along the best-supported parents / children, push the flow the path's
// transfrags allow, store the transcript, take its share out of the node coverages, repeat while a node has coverage >= 1.
// ex_*: [gno] scratch.  Returns the number of predictions stored.
STG_HDN inline int find_transcripts_warp(const Graph& g, TfG& f, FlowScratch& S, const AsmParams& P, PredOut& out, int graph_id,
                                         uint32_t* ex_s, uint32_t* ex_e, float* ex_c, uint32_t* err) {
  const int lane = wp_lane();
  const int gno = g.gno, nwn = g.nwn;
  const int nedges = g.ch_off[gno - 1] + g.ch_cnt[gno - 1];
  for (int i = lane; i < gno; i += WP_N) {
    float nc = 0.f;
    if (i && i < gno - 1 && g_len(g, i)) nc = g.n_cov[i] / (float)g_len(g, i);
    S.nodecov[i] = nc; S.usednode[i] = 0; S.prev_node[i] = 0; S.succ[i] = -1;
  }
  for (int e = lane; e < nedges; e += WP_N) S.prev_edge[e] = 0;
  wp_sync();
  int maxi = 0;
  for (int i = 1; i < gno; i++) if (S.nodecov[i] > S.nodecov[maxi]) maxi = i;
  int nstored = 0;
  float maxcov = 0.f;
  const int istr_words = (f.nt + 63) >> 6;
  int guard = 0;
  S.dbg_iters = 0;
  while (S.nodecov[maxi] >= 1.f && !*err) {
    S.dbg_iters++;
    // a fresh path
    for (int w = lane; w < nwn; w += WP_N) S.pathpat[w] = 0ull;
    for (int w = lane; w < istr_words; w += WP_N) S.istr[w] = 0ull;
    for (int i = lane; i < gno; i += WP_N) S.succ[i] = -1;
    wp_sync();
    if (lane == 0) { S.path[0] = maxi; fl_set_bit(S, maxi, true); }
    S.npath = 1;
    wp_sync();
    float flux = 0.f;
    if (fl_back_to_source(g, f, S, maxi)) {
      if (lane == 0) {
        if (P.includesource) S.path[S.npath] = 0;
        const int np = S.npath + (P.includesource ? 1 : 0);
        for (int a = 0, b = np - 1; a < b; a++, b--) { const int t = S.path[a]; S.path[a] = S.path[b]; S.path[b] = t; }
      }
      if (P.includesource) S.npath++;
      wp_sync();
      if (fl_fwd_to_sink(g, f, S, maxi)) flux = fl_push_max_flow(g, f, S);
    }
    if ((double)flux > ASM_EPSILON) {
      bool included = true;
      double cov = 0;
      if (lane == 0) cov = fl_store_transcript(g, S, P, out, graph_id, nstored, included, ex_s, ex_e, ex_c, err);
      wp_sync();
      cov = wp_shfl(cov, 0); included = wp_shfl(included ? 1 : 0, 0) != 0; nstored = wp_shfl(nstored, 0);
      if (included || cov < (double)wp_fmul(P.isofrac, maxcov)) {
        if (lane == 0) S.usednode[maxi] = 1;
        maxi = 0; maxcov = 0.f;
      } else if (cov > (double)maxcov) maxcov = (float)cov;
    } else {
      if (lane == 0) S.usednode[maxi] = 1;
      maxi = 0; maxcov = 0.f;
    }
    wp_sync();
    for (int i = 1; i < gno; i++) if (!S.usednode[i] && S.nodecov[i] > S.nodecov[maxi]) maxi = i;
    if (++guard > 1000000) { *err |= ASM_ERR_CAP; break; }
  }
  return nstored;
}
