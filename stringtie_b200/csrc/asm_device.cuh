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
