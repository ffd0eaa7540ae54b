// frag_device.cuh — SURVEY.md §8 row a11: every read / fragment of a bundle is mapped to the splice-graph nodes it
// covers and to its transfrag (the node list + the edges it vouches for).
//
// Reference: the loop at rlink.cpp:15796-15820 -> get_fragment_pattern :4870-5146 -> get_read_pattern :4354-4454 (node
// coverage, node lists) and update_abundance :4680-4867 (look the pattern up in the CTreePat trie, create the transfrag
// on first sight, add the abundance).  Sequential over the reads of a bundle in the reference; what is order-dependent:
//   * CGraphnode::cov and CTransfrag::abundance are float sums in read order;
//   * transfrags are numbered in order of first appearance (the order later decides ties of the reference's sort).
// Here the work is split in three:
//   1. fragment_pattern(): one THREAD per read works out, for the fragments the read opens (rlink.cpp:15810-15817), the
//      node-coverage contributions and the transfrag keys — a pure function of the read, its mate and the finished
//      graphs, so all reads of a batch run side by side (two passes: count, then write at the scanned offsets).
//   2. tf_commit_warp(): one WARP per bundle walks the bundle's keys in read order, 32 at a time: hash look-up of every
//      key in parallel, new keys numbered in order by ballot, abundances added per transfrag in key order.
//   3. nodecov_commit_warp(): one WARP per bundle: lane l owns the nodes with id = l (mod 32) and adds their
//      contributions in read order; no two lanes ever touch the same sum.
// Unitigs (super-reads: srabund, process_srfrag) are not supported: stg_assemble rejects batches that carry them.
#pragma once
#include "asm_device.cuh"

#define FR_MAXN 96       // nodes of one read (or of read + mate) inside one graph
#define FR_MAXG 4        // graphs one read can touch on one strand
#define FR_MAXB 32       // bundlenodes one read can touch
#if !defined(__CUDA_ARCH__) && defined(ASM_HOST_DEBUG)
#define FR_CAPFAIL(why) fprintf(stderr, "frag cap: %s\n", why)
#else
#define FR_CAPFAIL(why) ((void)0)
#endif

struct BundleCtx {               // one input bundle after rows a8 + a9 and the graphs of its two strands
  int nr, ng;
  const uint32_t *r_start, *r_end; const int16_t* r_nh; const uint8_t* r_flags; const float* r_count;
  const int8_t* strand;          // CReadAln::strand after the read loop of build_graphs
  const uint32_t* seg_off; uint32_t seg0; const uint32_t *seg_s, *seg_e;
  const uint32_t* pair_off; uint32_t pair0; const int32_t* pair; const float* pair_cnt;
  const int32_t* rj;             // junction row of intron i of read n: rj[seg_off[n] - seg0 - n + i]
  const uint32_t *j_start, *j_end; const int8_t* j_strand;   // the bundle's junction table after the loop
  const uint32_t* rg_off; uint32_t rg0; const int32_t* rg;   // readgroup[n] = rg[rg_off[n] - rg0 ...]
  const int32_t* merge; const float* negp; const int32_t* g2b;   // groups: merge[], neg_prop, group2bundle[lane * ng + gr]
  const int32_t *bn_graph[2], *bn_first[2], *bn_cnt[2];      // bundle2graph of strand s: graph, first node, node count
  const Graph* graphs;           // indexed by the ids in bn_graph
};

STG_HD int fr_nseg(const BundleCtx& B, int n) { return (int)(B.seg_off[n + 1] - B.seg_off[n]); }
STG_HD uint32_t fr_seg(const BundleCtx& B, int n, int i) { return B.seg_off[n] - B.seg0 + (uint32_t)i; }
STG_HD int fr_junc(const BundleCtx& B, int n, int i) { return B.rj[B.seg_off[n] - B.seg0 - (uint32_t)n + (uint32_t)i]; }
STG_HD uint32_t fr_jstart(const BundleCtx& B, int jr) { return jr >= 0 ? B.j_start[jr] : 0u; }
STG_HD uint32_t fr_jend(const BundleCtx& B, int jr) { return jr >= 0 ? B.j_end[jr] : 0u; }
STG_HD int fr_jstrand(const BundleCtx& B, int jr) { return jr >= 0 ? (int)B.j_strand[jr] : 0; }
STG_HD int fr_root(const BundleCtx& B, int gr) { while (B.merge[gr] != gr) gr = B.merge[gr]; return gr; }

struct ReadPat {                 // rgno / rnode of get_read_pattern
  int ng;
  int graph[FR_MAXG], cnt[FR_MAXG];
  int32_t node[FR_MAXG][FR_MAXN];
};

// get_read_pattern(s, ...), rlink.cpp:4354-4454, without the unitig clauses.  emit.cov(graph, node, value) once per
// (exon, node) overlap, in the reference's order.
template <class E>
STG_HDN inline void read_pattern(const BundleCtx& B, int s, int n, float rprop, float readcov, ReadPat& R, E& emit, uint32_t* err) {
  R.ng = 0;
  int lastgnode = -1, lastngraph = -1;
  const int ncoord = fr_nseg(B, n);
  int k = 0;
  int seen[FR_MAXB]; int nseen = 0;
  const uint32_t g0 = B.rg_off[n] - B.rg0, g1 = B.rg_off[n + 1] - B.rg0;
  for (uint32_t i = g0; i < g1; i++) {
    const int gr = fr_root(B, B.rg[i]);
    const int bnode = B.g2b[(size_t)(2 * s) * B.ng + gr];
    if (bnode < 0 || B.bn_cnt[s][bnode] <= 0) continue;
    bool dup = false;
    for (int x = 0; x < nseen; x++) if (seen[x] == bnode) { dup = true; break; }
    if (dup) continue;
    if (nseen >= FR_MAXB) { FR_CAPFAIL("bundlenodes"); *err |= ASM_ERR_CAP; return; }
    seen[nseen++] = bnode;
    const int nbnode = B.bn_cnt[s][bnode];
    const int ngraph = B.bn_graph[s][bnode];
    const Graph& G = B.graphs[ngraph];
    for (int j = 0; j < nbnode && k < ncoord; j++) {
      const int gnode = B.bn_first[s][bnode] + j;
      const uint32_t ns = G.n_start[gnode], ne = G.n_end[gnode];
      bool intersect = false;
      while (k < ncoord) {
        const uint32_t ss = B.seg_s[fr_seg(B, n, k)], se = B.seg_e[fr_seg(B, n, k)];
        if (se < ns) { k++; continue; }
        if (ss > ne) break;                     // overlapLen == 0
        const uint32_t lo = ss > ns ? ss : ns, hi = se < ne ? se : ne;
        const int bp = (int)(hi - lo + 1u);
        intersect = true;
        emit.cov(ngraph, gnode, wp_fmul(wp_fmul(rprop, (float)bp), readcov));
        if (se <= ne) k++; else break;
      }
      if (intersect) {
        int g = -1;
        for (int r = 0; r < R.ng; r++) if (R.graph[r] == ngraph) { g = r; break; }
        if (g < 0) {
          if (R.ng >= FR_MAXG) { FR_CAPFAIL("graphs"); *err |= ASM_ERR_CAP; return; }
          g = R.ng++; R.graph[g] = ngraph; R.cnt[g] = 0;
        }
        if (lastngraph != ngraph || lastgnode != gnode) {
          if (R.cnt[g] >= FR_MAXN) { FR_CAPFAIL("nodes"); *err |= ASM_ERR_CAP; return; }
          R.node[g][R.cnt[g]++] = gnode;
        }
        lastgnode = gnode; lastngraph = ngraph;
      }
    }
  }
}

// strand proportions of a fragment, rlink.cpp:4882-4947
STG_HD void fr_group_prop(const BundleCtx& B, int n, float& sum, int& ngroup) {
  const uint32_t g0 = B.rg_off[n] - B.rg0, g1 = B.rg_off[n + 1] - B.rg0;
  for (uint32_t i = g0; i < g1; i++) { ngroup++; sum = wp_fadd(sum, B.negp[fr_root(B, B.rg[i])]); }
}

// One fragment (read n, mate np or -1, weight readcov): rlink.cpp:4870-5146.
//   emit.cov(graph, node, value)                      CGraphnode::cov += value
//   emit.tf(graph, nodes, flags, count, abundance)    update_abundance(): flags[i] = the edge nodes[i-1] -> nodes[i] is
//                                                     part of the pattern (flags[0] unused)
template <class E>
STG_HDN inline void fragment_pattern(const BundleCtx& B, int n, int np, float readcov, E& emit, uint32_t* err) {
  float rprop[2] = {0.f, 0.f};
  const bool n_ok = B.r_nh[n] != 0;
  const bool p_ok = np > -1 && B.r_nh[np] != 0;
  if (n_ok && !B.strand[n] && p_ok && !B.strand[np]) {
    int ngroup = 0;
    fr_group_prop(B, n, rprop[0], ngroup);
    fr_group_prop(B, np, rprop[0], ngroup);
    if (ngroup) { rprop[0] = rprop[0] / (float)ngroup; rprop[1] = wp_fsub(1.f, rprop[0]); }
  } else if (n_ok) {
    if (!B.strand[n]) {
      int ngroup = 0;
      fr_group_prop(B, n, rprop[0], ngroup);
      if (ngroup) { rprop[0] = rprop[0] / (float)ngroup; rprop[1] = wp_fsub(1.f, rprop[0]); }
    } else if (B.strand[n] == -1) rprop[0] = 1.f; else rprop[1] = 1.f;
  } else if (p_ok) {
    if (!B.strand[np]) {
      int ngroup = 0;
      fr_group_prop(B, np, rprop[0], ngroup);
      if (ngroup) { rprop[0] = rprop[0] / (float)ngroup; rprop[1] = wp_fsub(1.f, rprop[0]); }
    } else if (B.strand[np] == -1) rprop[0] = 1.f; else rprop[1] = 1.f;
  }
  for (int s = 0; s < 2; s++) {
    if (!(rprop[s] != 0.f)) continue;
    ReadPat R, Q;
    R.ng = 0; Q.ng = 0;
    if (n_ok) read_pattern(B, s, n, rprop[s], readcov, R, emit, err);
    if (p_ok) read_pattern(B, s, np, rprop[s], readcov, Q, emit, err);
    if (*err) return;
    const float abundance = wp_fmul(rprop[s], readcov);
    int usedp = 0;
    bool pused[FR_MAXG] = {false, false, false, false};
    int32_t nodes[FR_MAXN]; uint8_t flags[FR_MAXN];
    const int nj_n = fr_nseg(B, n) - 1;
    const int nj_p = np > -1 ? fr_nseg(B, np) - 1 : 0;
    for (int r = 0; r < R.ng; r++) {
      const int gph = R.graph[r];
      const Graph& G = B.graphs[gph];
      const int cnt = R.cnt[r];
      const int32_t* rn = R.node[r];
      int i = 0, clear_rnode = 0;
      for (int j = 0; j < cnt; j++) {
        flags[j] = 0;
        if (!j) continue;
        const int prev = rn[j - 1], cur = rn[j];
        if (G.n_start[cur] - 1 == G.n_end[prev]) {            // continuous nodes
          flags[j] = g_is_child(G, prev, cur) ? 1 : 0;
        } else {
          while (i < nj_n && fr_jend(B, fr_junc(B, n, i)) < G.n_start[cur]) i++;
          if (i < nj_n && fr_jstart(B, fr_junc(B, n, i)) <= G.n_end[prev]) {
            const int jr = fr_junc(B, n, i);
            if (fr_jstrand(B, jr) && fr_jstart(B, jr) == G.n_end[prev] && fr_jend(B, jr) == G.n_start[cur]) {
              flags[j] = g_is_child(G, prev, cur) ? 1 : 0;
            } else {   // the read skips from prev to cur without a junction of the graph: what came so far is one transfrag
              if (j - clear_rnode > 1) emit.tf(gph, rn + clear_rnode, flags + clear_rnode, j - clear_rnode, abundance);
              clear_rnode = j;
            }
          }
        }
      }
      const int rc = cnt - clear_rnode;          // what is left of rnode[r]
      const int32_t* rr = rn + clear_rnode;
      const uint8_t* rf = flags + clear_rnode;
      int p = 0;
      while (p < Q.ng && (pused[p] || Q.graph[p] != gph)) p++;
      if (p < Q.ng) {
        usedp++;
        const int pc = Q.cnt[p];
        const int32_t* pn = Q.node[p];
        uint8_t pflags[FR_MAXN];
        int ip = 0;
        for (int j = 0; j < pc; j++) {
          pflags[j] = 0;
          if (!j) continue;
          const int prev = pn[j - 1], cur = pn[j];
          if (G.n_start[cur] - 1 == G.n_end[prev]) pflags[j] = g_is_child(G, prev, cur) ? 1 : 0;
          else {
            while (ip < nj_p && fr_jstart(B, fr_junc(B, np, ip)) < G.n_end[prev]) ip++;
            if (ip < nj_p) {
              const int jr = fr_junc(B, np, ip);
              if (fr_jstrand(B, jr) && fr_jstart(B, jr) == G.n_end[prev] && fr_jend(B, jr) == G.n_start[cur]) pflags[j] = g_is_child(G, prev, cur) ? 1 : 0;
            }
          }
        }
        // the mate continues the read's pattern unless one of the read's nodes cannot lead to the mate's first node
        const int p0 = pn[0];
        bool conflict = false;
        for (int j = 0; j < rc; j++) if (rr[j] != p0 && !g_reach(G, rr[j], p0)) { conflict = true; break; }
        if (!conflict) {
          int32_t mnodes[2 * FR_MAXN]; uint8_t mflags[2 * FR_MAXN];
          int mc = 0;
          for (int j = 0; j < rc; j++) { mnodes[mc] = rr[j]; mflags[mc] = j ? rf[j] : 0; mc++; }
          int jp = (p0 == rr[rc - 1]) ? 1 : 0;
          bool first = true;
          for (; jp < pc; jp++) {
            mnodes[mc] = pn[jp];
            mflags[mc] = (first && !(p0 == rr[rc - 1])) ? 0 : pflags[jp];
            first = false;
            mc++;
          }
          if (mc > 1) emit.tf(gph, mnodes, mflags, mc, abundance);
        } else {
          if (rc > 1) emit.tf(gph, rr, rf, rc, abundance);
          if (pc > 1) emit.tf(gph, pn, pflags, pc, abundance);
        }
        pused[p] = true;
      } else if (rc > 1) emit.tf(gph, rr, rf, rc, abundance);
    }
    if (usedp < Q.ng) {
      for (int p = 0; p < Q.ng; p++) {
        if (pused[p]) continue;
        const int gph = Q.graph[p];
        const Graph& G = B.graphs[gph];
        const int pc = Q.cnt[p];
        const int32_t* pn = Q.node[p];
        int ip = 0;
        for (int j = 0; j < pc; j++) {
          flags[j] = 0;
          if (!j) continue;
          const int prev = pn[j - 1], cur = pn[j];
          if (G.n_start[cur] - 1 == G.n_end[prev]) flags[j] = g_is_child(G, prev, cur) ? 1 : 0;
          else {
            while (ip < nj_p && fr_jstart(B, fr_junc(B, np, ip)) < G.n_end[prev]) ip++;
            if (ip < nj_p) {
              const int jr = fr_junc(B, np, ip);
              if (fr_jstrand(B, jr) && fr_jstart(B, jr) == G.n_end[prev] && fr_jend(B, jr) == G.n_start[cur]) flags[j] = g_is_child(G, prev, cur) ? 1 : 0;
            }
          }
        }
        if (pc > 1) emit.tf(gph, pn, flags, pc, abundance);
      }
    }
    (void)nodes;
  }
}

// every fragment read n opens, in the reference's order (rlink.cpp:15796-15820)
template <class E>
STG_HDN inline void read_fragments(const BundleCtx& B, int n, E& emit, uint32_t* err) {
  float single_count = B.r_count[n];
  const uint32_t p0 = B.pair_off[n] - B.pair0, p1 = B.pair_off[n + 1] - B.pair0;
  for (uint32_t j = p0; j < p1; j++) {
    const int np = B.pair[j];
    if (np > -1) {
      single_count = wp_fsub(single_count, B.pair_cnt[j]);
      if (n < np) fragment_pattern(B, n, np, B.pair_cnt[j], emit, err);
    }
  }
  if ((double)single_count > ASM_EPSILON) fragment_pattern(B, n, -1, single_count, emit, err);
}

// ---- emitters of the two passes ---------------------------------------------------------------------------------
struct FragCount {
  uint32_t ncov, ntf, nkey;
  STG_HD void cov(int, int, float) { ncov++; }
  STG_HD void tf(int, const int32_t*, const uint8_t*, int cnt, float) { ntf++; nkey += (uint32_t)cnt; }
};

// a key node: the node id, bit 31 = "the edge from the previous node of the key is part of the pattern"
#define FR_FLAG 0x80000000u

STG_HD unsigned long long fr_mix(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  return h;
}

struct FragWrite {
  const int32_t* graph_node_off;   // global id of node 0 of every graph
  uint32_t icov, itf, ikey;        // running output positions (start at the read's scanned offsets)
  uint32_t* cov_node; float* cov_val;                       // [ncov]  global node id, value
  int32_t* tf_graph; float* tf_ab; uint32_t* tf_key_off; uint32_t* tf_key_len; unsigned long long* tf_hash;   // [ntf]
  uint32_t* key;                                            // [nkey]
  STG_HD void cov(int graph, int node, float v) { cov_node[icov] = (uint32_t)(graph_node_off[graph] + node); cov_val[icov] = v; icov++; }
  STG_HD void tf(int graph, const int32_t* nodes, const uint8_t* flags, int cnt, float ab) {
    unsigned long long h = fr_mix(0x1234567ull, (unsigned long long)(uint32_t)graph);
    for (int i = 0; i < cnt; i++) {
      const uint32_t kv = (uint32_t)nodes[i] | ((i && flags[i]) ? FR_FLAG : 0u);
      key[ikey + i] = kv;
      h = fr_mix(h, kv);
    }
    tf_graph[itf] = graph; tf_ab[itf] = ab; tf_key_off[itf] = ikey; tf_key_len[itf] = (uint32_t)cnt; tf_hash[itf] = h ? h : 1ull;
    itf++; ikey += (uint32_t)cnt;
  }
};

// ---- 2. transfrag table of one bundle ---------------------------------------------------------------------------
// Records [rec0, rec1) of the batch-wide record arrays belong to this bundle, in read order.  The table (open
// addressing, `cap` slots, a power of two, zero-initialised) maps hash -> transfrag id; the ids count the bundle's
// distinct keys in order of first appearance.  rec_tf[i] receives the id of record i; tf_rep[id] = first record with
// that key (its key storage is the transfrag's node list), tf_abund[id] = the abundance after all records.
// The transfrags create_graph left (two nodes) are entered by the caller as records in front of the reads'.
struct TfTable {
  unsigned long long* slot_hash; int32_t* slot_id; uint32_t cap;
  const int32_t* rec_graph; const float* rec_ab; const uint32_t *rec_key_off, *rec_key_len; const unsigned long long* rec_hash;
  const uint32_t* key;
  int32_t* rec_tf;          // [records]  out
  int32_t* tf_rep; float* tf_abund; int32_t* tf_graph;   // [<= records] out, indexed by the bundle-local transfrag id
};

STG_HD bool tf_same_key(const TfTable& T, uint32_t a, uint32_t b) {
  if (T.rec_graph[a] != T.rec_graph[b] || T.rec_key_len[a] != T.rec_key_len[b]) return false;
  const uint32_t* ka = T.key + T.rec_key_off[a];
  const uint32_t* kb = T.key + T.rec_key_off[b];
  for (uint32_t i = 0; i < T.rec_key_len[a]; i++) if (ka[i] != kb[i]) return false;
  return true;
}

// ntf: distinct transfrags so far (0 for the first range of a bundle); returns the new count
STG_HDN inline int tf_commit_warp(const TfTable& T, uint32_t rec0, uint32_t rec1, float* stage /* [WP_N] shared scratch */, int ntf) {
  const int lane = wp_lane();
  for (uint32_t base = rec0; base < rec1; base += WP_N) {
    const uint32_t r = base + (uint32_t)lane;
    const bool have = r < rec1;
    int id = -1;
    uint32_t slot = 0;
    unsigned long long h = 0;
    if (have) {
      h = T.rec_hash[r];
      slot = (uint32_t)(h >> 17) & (T.cap - 1);
      for (;;) {   // probe: stop at the key or at an empty slot
        const unsigned long long sh = T.slot_hash[slot];
        if (sh == 0ull) break;
        if (sh == h && tf_same_key(T, (uint32_t)T.tf_rep[T.slot_id[slot]], r)) { id = T.slot_id[slot]; break; }
        slot = (slot + 1) & (T.cap - 1);
      }
    }
    // new keys, in record order: one lane at a time enters its key (another lane of this step may hold the same key,
    // or a different key that wants the same slot)
    uint32_t pending = wp_ballot(have && id < 0);
    while (pending) {
      const int src = wp_ffs(pending) - 1;
      pending &= pending - 1;
      if (lane == src) {
        for (;;) {   // re-probe from where it stopped: earlier lanes of this step may have filled slots
          const unsigned long long sh = T.slot_hash[slot];
          if (sh == 0ull) break;
          if (sh == h && tf_same_key(T, (uint32_t)T.tf_rep[T.slot_id[slot]], r)) { id = T.slot_id[slot]; break; }
          slot = (slot + 1) & (T.cap - 1);
        }
        if (id < 0) {
          id = ntf;
          T.slot_hash[slot] = h; T.slot_id[slot] = id;
          T.tf_rep[id] = (int32_t)r; T.tf_abund[id] = 0.f; T.tf_graph[id] = T.rec_graph[r];
        }
      }
      wp_sync();
      const int got = wp_shfl(id, src);
      if (got >= ntf) ntf = got + 1;
    }
    // abundances: per transfrag in record order.  Lane l stages its value; the first lane of every group of equal ids
    // adds the group's values in lane order.
    if (have) { T.rec_tf[r] = id; }
    stage[lane] = have ? T.rec_ab[r] : 0.f;
    wp_sync();
#ifdef __CUDA_ARCH__
    const uint32_t peers = __match_any_sync(WP_FULL, have ? id : -1 - lane);
#else
    const uint32_t peers = 1u;
#endif
    if (have && (wp_ffs(peers) - 1) == lane) {
      float a = T.tf_abund[id];
      uint32_t mm = peers;
      while (mm) { const int l = wp_ffs(mm) - 1; mm &= mm - 1; a = wp_fadd(a, stage[l]); }
      T.tf_abund[id] = a;
    }
    wp_sync();
  }
  return ntf;
}

// ---- 3. node coverage of one bundle ------------------------------------------------------------------------------
// contributions [c0, c1) of the batch-wide arrays, in read order; n_cov is the batch-wide node coverage array
STG_HDN inline void nodecov_commit_warp(const uint32_t* __restrict__ cov_node, const float* __restrict__ cov_val, uint32_t c0, uint32_t c1,
                                        float* n_cov) {
  const int lane = wp_lane();
  for (uint32_t base = c0; base < c1; base += WP_N) {
    const uint32_t i = base + (uint32_t)lane;
    const uint32_t node = i < c1 ? cov_node[i] : 0xffffffffu;
    const float val = i < c1 ? cov_val[i] : 0.f;
    const int cnt = (int)((c1 - base) < (uint32_t)WP_N ? (c1 - base) : (uint32_t)WP_N);
    for (int l = 0; l < cnt; l++) {
      const uint32_t nd = wp_shfl(node, l);
      const float v = wp_shfl(val, l);
      if ((int)(nd % (uint32_t)WP_N) == lane) n_cov[nd] = wp_fadd(n_cov[nd], v);
    }
  }
}
