// tools/asm_hostcheck.cu — DEVELOPMENT AID, not part of libstgpu: compiles the per-graph device code
// (stringtie_b200/csrc/asm_device.cuh, frag_device.cuh) for the HOST (one-lane warps, warp_prims.cuh) and steps through
// it on a dump of the reference (`oracle/_ref/ref_hot --full --dump`, or a golden fixture converted by
// tools/asm_hostcheck.py), starting from the reference's OWN state after the bundles were formed (a9 / a9b arrays), so
// that the graph / transfrag / flow code can be debugged on a CPU-only box against the reference's stage dumps.
// Writes what it computed as a10_* / a11_* / a12_* / hp_* arrays for tools/asm_hostcheck.py to compare.
//   nvcc -std=c++17 -O1 -gencode arch=compute_100a,code=sm_100a -Xcompiler -ffp-contract=off -o tools/_build/asm_hostcheck tools/asm_hostcheck.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../stringtie_b200/csrc/asm_device.cuh"
#include "../stringtie_b200/csrc/frag_device.cuh"

namespace {
struct FileArr { uint32_t dtype; uint64_t count; std::vector<char> bytes; };
typedef std::map<std::string, FileArr> FileMap;
const size_t ES[] = {0, 1, 1, 2, 4, 4, 4, 8, 8};

bool read_stgb(const char* path, FileMap& m) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  uint32_t hdr[3];
  if (fread(hdr, 4, 3, f) != 3 || hdr[0] != STG_FILE_MAGIC) { fclose(f); return false; }
  for (uint32_t i = 0; i < hdr[2]; i++) {
    char name[25]; name[24] = 0; uint32_t dt[2]; uint64_t cnt;
    if (fread(name, 1, 24, f) != 24 || fread(dt, 4, 2, f) != 2 || fread(&cnt, 8, 1, f) != 1) { fclose(f); return false; }
    FileArr a; a.dtype = dt[0]; a.count = cnt;
    size_t nb = cnt * ES[dt[0]];
    a.bytes.resize(nb);
    if (nb && fread(a.bytes.data(), 1, nb, f) != nb) { fclose(f); return false; }
    size_t pad = (8 - nb % 8) % 8; char z[8];
    if (pad && fread(z, 1, pad, f) != pad) { fclose(f); return false; }
    m[name] = std::move(a);
  }
  fclose(f);
  return true;
}
template <class T> const T* arr(const FileMap& m, const char* name) {
  FileMap::const_iterator it = m.find(name);
  if (it == m.end()) { fprintf(stderr, "asm_hostcheck: array %s missing\n", name); exit(2); }
  return (const T*)it->second.bytes.data();
}
size_t arrn(const FileMap& m, const char* name) { FileMap::const_iterator it = m.find(name); return it == m.end() ? 0 : (size_t)it->second.count; }

struct Out {
  std::vector<std::pair<std::string, std::pair<uint32_t, std::vector<char>>>> arrs;
  template <class T> std::vector<char>& add(const char* name, uint32_t dt, const std::vector<T>& v) {
    std::vector<char> b((const char*)v.data(), (const char*)v.data() + v.size() * sizeof(T));
    arrs.push_back({name, {dt, b}});
    return arrs.back().second.second;
  }
  void write(const char* path) {
    FILE* f = fopen(path, "wb");
    uint32_t hdr[3] = {STG_FILE_MAGIC, 1u, (uint32_t)arrs.size()};
    fwrite(hdr, 4, 3, f);
    for (auto& a : arrs) {
      char name[24]; memset(name, 0, 24); strncpy(name, a.first.c_str(), 23);
      fwrite(name, 1, 24, f);
      uint32_t dt[2] = {a.second.first, 0}; fwrite(dt, 4, 2, f);
      uint64_t cnt = a.second.second.size() / ES[a.second.first]; fwrite(&cnt, 8, 1, f);
      if (!a.second.second.empty()) fwrite(a.second.second.data(), 1, a.second.second.size(), f);
      size_t pad = (8 - a.second.second.size() % 8) % 8; char z[8] = {0};
      if (pad) fwrite(z, 1, pad, f);
    }
    fclose(f);
  }
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: asm_hostcheck <dump.stgb> <out.stgb> [stage: 10|11|12|15]\n"); return 2; }
  const int stage = argc > 3 ? atoi(argv[3]) : 15;
  FileMap m;
  if (!read_stgb(argv[1], m)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const size_t nb = arrn(m, "b_start");
  const int32_t* b_start = arr<int32_t>(m, "b_start"); const int32_t* b_end = arr<int32_t>(m, "b_end");
  const uint32_t* b_read_off = arr<uint32_t>(m, "b_read_off");
  const int64_t* cov_off = arr<int64_t>(m, "s1_cov_off"); const float* cov = arr<float>(m, "s1_cov");
  const uint32_t* a9_junc_off = arr<uint32_t>(m, "a9_junc_off");
  const uint32_t* j_start = arr<uint32_t>(m, "a9_j_start"); const uint32_t* j_end = arr<uint32_t>(m, "a9_j_end");
  const int8_t* j_strand = arr<int8_t>(m, "a9_j_strand");
  const uint32_t* group_off = arr<uint32_t>(m, "a9_group_off");
  const uint32_t* bun_off = arr<uint32_t>(m, "a9b_bun_off"); const uint32_t* bn_off = arr<uint32_t>(m, "a9b_bn_off");
  const int32_t* bun_len = arr<int32_t>(m, "a9b_bun_len"); const int32_t* bun_start = arr<int32_t>(m, "a9b_bun_start");
  const int32_t* bun_last = arr<int32_t>(m, "a9b_bun_last"); const float* bun_cov = arr<float>(m, "a9b_bun_cov");
  const float* bun_multi = arr<float>(m, "a9b_bun_multi");
  const uint32_t* bn_start = arr<uint32_t>(m, "a9b_bn_start"); const uint32_t* bn_end = arr<uint32_t>(m, "a9b_bn_end");
  const int32_t* bn_next = arr<int32_t>(m, "a9b_bn_next");
  const int32_t* g2b_all = arr<int32_t>(m, "a9b_g2b"); const float* negp_all = arr<float>(m, "a9b_g_neg_prop");
  const int32_t* merge_all = arr<int32_t>(m, "a9_merge");
  const uint32_t* rg_off = arr<uint32_t>(m, "a9_rg_off"); const int32_t* rg_all = arr<int32_t>(m, "a9_rg");
  const uint32_t* r_start = arr<uint32_t>(m, "r_start"); const uint32_t* r_end = arr<uint32_t>(m, "r_end");
  const float* r_count = arr<float>(m, "r_count"); const int16_t* r_nh = arr<int16_t>(m, "r_nh");
  const uint8_t* r_flags = arr<uint8_t>(m, "r_flags");
  const uint32_t* r_seg_off = arr<uint32_t>(m, "r_seg_off"); const uint32_t* r_pair_off = arr<uint32_t>(m, "r_pair_off");
  const float* pair_cnt = arr<float>(m, "pair_cnt");
  const int8_t* a9_strand = arr<int8_t>(m, "a9_r_strand"); const int32_t* a9_pair = arr<int32_t>(m, "a9_pair_idx");
  const uint32_t* a9_seg_s = arr<uint32_t>(m, "a9_seg_start"); const uint32_t* a9_seg_e = arr<uint32_t>(m, "a9_seg_end");
  const int32_t* a9_rjunc = arr<int32_t>(m, "a9_rjunc");
  (void)bun_len; (void)r_end;

  AsmParams P;
  P.junctionsupport = 10; P.trim = 1; P.mcov = 1.f; P.mintranscriptlen = 200; P.allowed_nodes = 1000; P.isofrac = 0.01f;
  P.includesource = 1;

  Out out;
  std::vector<uint32_t> o_graph_off{0}, o_node_off{0}, o_par_off{0}, o_child_off{0}, o_tf_off{0}, o_tf_node_off{0}, o_nstart, o_nend;
  std::vector<int8_t> o_gs; std::vector<int32_t> o_gb, o_gno, o_edgeno, o_par, o_child, o_tf_nodes;
  std::vector<float> o_tf_ab;
  // a11
  std::vector<uint32_t> q_graph_off{0}, q_node_off{0}, q_tf_off{0}, q_tf_node_off{0}, q_tf_cnt_off{0};
  std::vector<float> q_ncov, q_tf_ab; std::vector<int32_t> q_tf_nodes, q_tf_pcount;
  // a12
  std::vector<uint32_t> w_graph_off{0}, w_tf_off{0}, w_tf_node_off{0}, w_path_off{0}, w_node_off{0}, w_ntrf_off{0};
  std::vector<int8_t> w_gs; std::vector<int32_t> w_gb, w_tf_nodes, w_path_node, w_path_cont, w_ntrf; std::vector<float> w_tf_ab, w_path_ab;
  std::vector<uint8_t> w_tf_real;
  // predictions
  std::vector<uint32_t> h_pred_off{0}, h_exon_off{0}, h_es, h_ee; std::vector<int32_t> h_start, h_end, h_tlen, h_geneno;
  std::vector<double> h_cov; std::vector<int8_t> h_strand; std::vector<float> h_ecov;

  uint32_t err = 0;
  for (size_t b = 0; b < nb; b++) {
    const uint32_t refstart = (uint32_t)b_start[b];
    const uint32_t covlen = (uint32_t)(b_end[b] - b_start[b] + 3);
    const float* c0 = cov + cov_off[b];
    const int ng = (int)(group_off[b + 1] - group_off[b]);
    const uint32_t jo = a9_junc_off[b]; const int nj = (int)(a9_junc_off[b + 1] - jo);
    const int r0 = (int)b_read_off[b], nr = (int)(b_read_off[b + 1] - b_read_off[b]);
    // per bundle: the graphs of both strands
    std::vector<std::vector<uint32_t>> js_s(2), js_e(2), je_e(2); std::vector<std::vector<int32_t>> je_k(2), regn(2), regg(2);
    struct GStore {
      std::vector<uint32_t> n_start, n_end; std::vector<float> n_cov; std::vector<int32_t> ch_off, ch_cnt, ch, pa_off, pa_cnt, pa;
      std::vector<uint8_t> sinkpar; std::vector<unsigned long long> reach; std::vector<int32_t> tf_n1, tf_n2; std::vector<float> tf_ab;
    };
    std::vector<GStore> store;
    std::vector<Graph> graphs;
    std::vector<int> g_s, g_b;
    std::vector<std::vector<int32_t>> bnfirst(2), bncnt(2), bngraph(2);
    int first_graph[3] = {0, 0, 0};
    for (int s = 0; s < 2; s++) {
      const int lane = 2 * s;
      // strand-filtered junction lists (junc_lists_warp is the device function the kernel uses)
      js_s[s].resize(nj + 1); js_e[s].resize(nj + 1); je_e[s].resize(nj + 1); je_k[s].resize(nj + 1);
      const int nJ = junc_lists_warp(j_start + jo, j_end + jo, j_strand + jo, nj, 2 * s - 1, js_s[s].data(), js_e[s].data(), je_e[s].data(), je_k[s].data());
      regn[s].assign(nJ + 1, -1); regg[s].assign(nJ + 1, -1);
      const uint32_t bu0 = bun_off[3 * b + lane], bu1 = bun_off[3 * b + lane + 1];
      const uint32_t bn0 = bn_off[3 * b + lane], bn1 = bn_off[3 * b + lane + 1];
      const int nbn = (int)(bn1 - bn0);
      bnfirst[s].assign(nbn + 1, -1); bncnt[s].assign(nbn + 1, 0); bngraph[s].assign(nbn + 1, -1);
      first_graph[s] = (int)graphs.size();
      for (uint32_t k = bu0; k < bu1; k++) {
        Graph g; memset(&g, 0, sizeof(g));
        store.emplace_back();
        const int gid = (int)graphs.size();
        g_s.push_back(s); g_b.push_back((int)(k - bu0));
        bool process = bun_cov[k] != 0.f && (bun_multi[k] / bun_cov[k]) <= P.mcov && bun_len[k] >= P.mintranscriptlen;
        if (process) {
          GraphSrc G;
          G.s = s; G.refstart = refstart; G.covlen = covlen; G.c_all = c0 + covlen; G.c_opp = c0 + (size_t)(2 - 2 * s) * covlen;
          G.bn_start = bn_start + bn0; G.bn_end = bn_end + bn0; G.bn_next = bn_next + bn0;
          G.startnode = bun_start[k]; G.lastnode = bun_last[k];
          G.js_start = js_s[s].data(); G.js_end = js_e[s].data(); G.nJs = nJ; G.je_end = je_e[s].data(); G.je_k = je_k[s].data(); G.nJe = nJ;
          G.reg_node = regn[s].data(); G.reg_graph = regg[s].data(); G.gid = gid; G.junctionsupport = P.junctionsupport; G.trim = P.trim;
          // bounds
          uint32_t maxlen = 0; int ncap = 2, nbns = 0;
          for (int x = G.startnode; x != -1; x = G.bn_next[x]) {
            const uint32_t len = G.bn_end[x] - G.bn_start[x] + 1; if (len > maxlen) maxlen = len;
            ncap += 2 + (int)(len / 50); nbns++;
          }
          ncap += 2 * nJ + 2 * nJ;
          const int ecap = nJ + 6 * ncap, fcap = 4 * ncap;
          std::vector<uint32_t> tn_s(ncap), tn_e(ncap), tr_pos(maxlen / 50 + 4); std::vector<int32_t> ev_u(ecap), ev_v(ecap), ft1(fcap), ft2(fcap);
          std::vector<float> ftab(fcap), tr_ab(maxlen / 50 + 4); std::vector<uint8_t> tr_st(maxlen / 50 + 4);
          SweepOut o; memset(&o, 0, sizeof(o));
          o.n_start = tn_s.data(); o.n_end = tn_e.data(); o.ev_u = ev_u.data(); o.ev_v = ev_v.data(); o.ft_n1 = ft1.data(); o.ft_n2 = ft2.data();
          o.ft_ab = ftab.data(); o.bn_first = bnfirst[s].data(); o.bn_cnt = bncnt[s].data(); o.tr_pos = tr_pos.data(); o.tr_ab = tr_ab.data();
          o.tr_st = tr_st.data(); o.tr_cap = maxlen / 50 + 4; o.ncap = ncap; o.ecap = ecap; o.fcap = fcap;
          graph_sweep(G, o);
          err |= o.err;
          for (int x = G.startnode; x != -1; x = G.bn_next[x]) bngraph[s][x] = gid;
          GStore& S = store.back();
          const int gno = o.nn + 1, nwn = (gno + 63) / 64;
          S.n_start.resize(gno); S.n_end.resize(gno); S.n_cov.resize(gno); S.ch_off.resize(gno + 1); S.ch_cnt.resize(gno + 1);
          S.pa_off.resize(gno + 1); S.pa_cnt.resize(gno + 1); S.ch.resize(o.ne + 2 * gno + 2); S.pa.resize(o.ne + 2 * gno + 2);
          S.sinkpar.resize(gno + 1); S.reach.assign((size_t)gno * nwn, 0); S.tf_n1.resize(fcap + 2 * gno + 2); S.tf_n2.resize(fcap + 2 * gno + 2);
          S.tf_ab.resize(fcap + 2 * gno + 2);
          g.n_start = S.n_start.data(); g.n_end = S.n_end.data(); g.n_cov = S.n_cov.data(); g.ch_off = S.ch_off.data(); g.ch_cnt = S.ch_cnt.data();
          g.ch = S.ch.data(); g.pa_off = S.pa_off.data(); g.pa_cnt = S.pa_cnt.data(); g.pa = S.pa.data(); g.n_sinkpar = S.sinkpar.data();
          g.reach = S.reach.data(); g.tf_n1 = S.tf_n1.data(); g.tf_n2 = S.tf_n2.data(); g.tf_ab = S.tf_ab.data();
          std::vector<int32_t> stack(2 * (gno + 2)); std::vector<double> ncs(gno + 2); std::vector<uint8_t> visit(gno + 2);
          graph_finish(G, o, g, stack.data(), ncs.data(), visit.data(), P.allowed_nodes, &err);
        }
        graphs.push_back(g);
      }
    }
    first_graph[2] = (int)graphs.size();
    // ---- a10 record
    for (size_t gi = 0; gi < graphs.size(); gi++) {
      const Graph& g = graphs[gi];
      o_gs.push_back((int8_t)g_s[gi]); o_gb.push_back(g_b[gi]); o_gno.push_back(g.gno); o_edgeno.push_back(g.gno ? g.edgeno : 0);
      if (g.gno) {
        for (int i = 0; i < g.gno; i++) {
          o_nstart.push_back(g.n_start[i]); o_nend.push_back(g.n_end[i]);
          if (i == g.gno - 1) { for (int u = 0; u < g.gno; u++) if (g.n_sinkpar[u]) o_par.push_back(u); }
          else for (int k = 0; k < g.pa_cnt[i]; k++) o_par.push_back(g.pa[g.pa_off[i] + k]);
          for (int k = 0; k < g.ch_cnt[i]; k++) o_child.push_back(g.ch[g.ch_off[i] + k]);
          o_par_off.push_back((uint32_t)o_par.size()); o_child_off.push_back((uint32_t)o_child.size());
        }
        for (int t = 0; t < g.ntf0; t++) {
          o_tf_ab.push_back(g.tf_ab[t]); o_tf_nodes.push_back(g.tf_n1[t]); o_tf_nodes.push_back(g.tf_n2[t]);
          o_tf_node_off.push_back((uint32_t)o_tf_nodes.size());
        }
      }
      o_node_off.push_back((uint32_t)o_nstart.size()); o_tf_off.push_back((uint32_t)o_tf_ab.size());
    }
    o_graph_off.push_back((uint32_t)o_gs.size());
    if (stage < 11) continue;
    (void)first_graph;
    // ---- a11: reads -> node coverage + transfrags
    BundleCtx B; memset(&B, 0, sizeof(B));
    B.nr = nr; B.ng = ng;
    B.r_start = r_start + r0; B.r_end = r_end + r0; B.r_nh = r_nh + r0; B.r_flags = r_flags + r0; B.r_count = r_count + r0;
    B.strand = a9_strand + r0; B.seg_off = r_seg_off + r0; B.seg0 = r_seg_off[r0]; B.seg_s = a9_seg_s + r_seg_off[r0]; B.seg_e = a9_seg_e + r_seg_off[r0];
    B.pair_off = r_pair_off + r0; B.pair0 = r_pair_off[r0]; B.pair = a9_pair + r_pair_off[r0]; B.pair_cnt = pair_cnt + r_pair_off[r0];
    B.rj = a9_rjunc + (r_seg_off[r0] - r0);
    B.j_start = j_start + jo; B.j_end = j_end + jo; B.j_strand = j_strand + jo;
    B.rg_off = rg_off + r0; B.rg0 = rg_off[r0]; B.rg = rg_all + rg_off[r0];
    B.merge = merge_all + group_off[b]; B.negp = negp_all + group_off[b]; B.g2b = g2b_all + 3 * (size_t)group_off[b];
    for (int s = 0; s < 2; s++) { B.bn_graph[s] = bngraph[s].data(); B.bn_first[s] = bnfirst[s].data(); B.bn_cnt[s] = bncnt[s].data(); }
    B.graphs = graphs.data();
    std::vector<int32_t> gnode_off(graphs.size() + 1, 0);
    for (size_t gi = 0; gi < graphs.size(); gi++) gnode_off[gi + 1] = gnode_off[gi] + graphs[gi].gno;
    std::vector<uint32_t> c_off(nr + 1, 0), t_off(nr + 1, 0), k_off(nr + 1, 0);
    for (int n = 0; n < nr; n++) {
      FragCount fc{0, 0, 0};
      read_fragments(B, n, fc, &err);
      c_off[n + 1] = c_off[n] + fc.ncov; t_off[n + 1] = t_off[n] + fc.ntf; k_off[n + 1] = k_off[n] + fc.nkey;
    }
    int nt0 = 0; for (auto& g : graphs) nt0 += g.gno ? g.ntf0 : 0;
    const uint32_t NREC = (uint32_t)nt0 + t_off[nr], NKEY = 2u * (uint32_t)nt0 + k_off[nr];
    std::vector<uint32_t> cov_node(c_off[nr] + 1), rec_koff(NREC + 1), rec_klen(NREC + 1), key(NKEY + 1); std::vector<float> cov_val(c_off[nr] + 1), rec_ab(NREC + 1);
    std::vector<int32_t> rec_graph(NREC + 1), rec_tf(NREC + 1), tf_rep(NREC + 1), tf_graph(NREC + 1); std::vector<unsigned long long> rec_hash(NREC + 1);
    std::vector<float> tf_abund(NREC + 1);
    {  // the transfrags create_graph left, as records in front
      FragWrite w; memset(&w, 0, sizeof(w));
      w.graph_node_off = gnode_off.data(); w.cov_node = cov_node.data(); w.cov_val = cov_val.data(); w.tf_graph = rec_graph.data(); w.tf_ab = rec_ab.data();
      w.tf_key_off = rec_koff.data(); w.tf_key_len = rec_klen.data(); w.tf_hash = rec_hash.data(); w.key = key.data();
      for (size_t gi = 0; gi < graphs.size(); gi++) if (graphs[gi].gno) for (int t = 0; t < graphs[gi].ntf0; t++) {
        int32_t nd[2] = {graphs[gi].tf_n1[t], graphs[gi].tf_n2[t]}; uint8_t fl[2] = {0, 1};
        w.tf((int)gi, nd, fl, 2, graphs[gi].tf_ab[t]);
      }
      for (int n = 0; n < nr; n++) {
        w.icov = c_off[n]; w.itf = (uint32_t)nt0 + t_off[n]; w.ikey = 2u * (uint32_t)nt0 + k_off[n];
        read_fragments(B, n, w, &err);
      }
    }
    uint32_t cap = 16; while (cap < 2 * NREC + 2) cap <<= 1;
    std::vector<unsigned long long> slot_hash(cap, 0ull); std::vector<int32_t> slot_id(cap, -1);
    TfTable T; T.slot_hash = slot_hash.data(); T.slot_id = slot_id.data(); T.cap = cap; T.rec_graph = rec_graph.data(); T.rec_ab = rec_ab.data();
    T.rec_key_off = rec_koff.data(); T.rec_key_len = rec_klen.data(); T.rec_hash = rec_hash.data(); T.key = key.data(); T.rec_tf = rec_tf.data();
    T.tf_rep = tf_rep.data(); T.tf_abund = tf_abund.data(); T.tf_graph = tf_graph.data();
    float stagebuf[32];
    int ntf = tf_commit_warp(T, 0, (uint32_t)nt0, stagebuf, 0);
    ntf = tf_commit_warp(T, (uint32_t)nt0, NREC, stagebuf, ntf);
    std::vector<float> ncov_flat(gnode_off.back() + 1, 0.f);
    nodecov_commit_warp(cov_node.data(), cov_val.data(), 0, c_off[nr], ncov_flat.data());
    for (size_t gi = 0; gi < graphs.size(); gi++) for (int i = 0; i < graphs[gi].gno; i++) graphs[gi].n_cov[i] = ncov_flat[gnode_off[gi] + i];
    // per graph transfrag lists, in order of first appearance
    std::vector<std::vector<int>> gtf(graphs.size());
    for (int t = 0; t < ntf; t++) gtf[tf_graph[t]].push_back(t);
    for (size_t gi = 0; gi < graphs.size(); gi++) {
      const Graph& g = graphs[gi];
      for (int i = 0; i < g.gno; i++) q_ncov.push_back(g.n_cov[i]);
      for (int t : gtf[gi]) {
        q_tf_ab.push_back(tf_abund[t]);
        const uint32_t ko = rec_koff[tf_rep[t]], kl = rec_klen[tf_rep[t]];
        int pc = 0;
        for (uint32_t x = 0; x < kl; x++) { q_tf_nodes.push_back((int32_t)(key[ko + x] & ~FR_FLAG)); pc += 1 + ((key[ko + x] & FR_FLAG) ? 1 : 0); }
        q_tf_node_off.push_back((uint32_t)q_tf_nodes.size()); q_tf_pcount.push_back(pc);
      }
      q_node_off.push_back((uint32_t)q_ncov.size()); q_tf_off.push_back((uint32_t)q_tf_ab.size());
    }
    q_graph_off.push_back((uint32_t)(q_node_off.size() - 1));
    if (stage < 12) { h_pred_off.push_back((uint32_t)h_start.size()); continue; }
    // ---- a12: process_transfrags, then a13-a15 per graph, in the reference's processing order
    const uint32_t* a19_off = arr<uint32_t>(m, "a19_pred_off"); const int32_t* a19_geneno = arr<int32_t>(m, "a19_geneno");
    int geneno = a19_off[b + 1] > a19_off[b] ? a19_geneno[a19_off[b + 1] - 1] + 1 : 0;
    for (size_t gi = 0; gi < graphs.size(); gi++) {
      Graph& g = graphs[gi];
      if (!g.gno) continue;
      const int gno = g.gno, nwn = g.nwn;
      const int nedges = g.ch_off[gno - 1] + g.ch_cnt[gno - 1];
      const int nt0 = (int)gtf[gi].size();
      const int cap = nt0 + nedges + 1;
      std::vector<uint32_t> keyx(key);   // this graph's private copy of the key pool + room for the edge transfrags
      const uint32_t xkey0 = (uint32_t)keyx.size();
      keyx.resize(keyx.size() + 2 * (size_t)nedges + 2);
      std::vector<uint32_t> koff(cap), klen(cap), wkoff(cap), wklen(cap); std::vector<float> ab(cap), wab(cap), path_ab; std::vector<uint8_t> real(cap), ecov(nedges + 1);
      std::vector<int32_t> usepath(cap), path_off(cap), path_cnt(cap), path_node, path_cont, order(cap), ntrf_off(gno + 1), ntrf_cnt(gno + 1), ntrf;
      std::vector<unsigned long long> pat((size_t)cap * nwn), skey(cap);
      int gmax = 1; for (int i = 1; i < gno; i++) if (g.ch_cnt[i] > gmax) gmax = g.ch_cnt[i];
      int ngap = 0; size_t reach_sum = 0;
      for (int x = 0; x < nt0; x++) {
        const int t = gtf[gi][x];
        wkoff[x] = rec_koff[tf_rep[t]]; wklen[x] = rec_klen[tf_rep[t]]; wab[x] = tf_abund[t];
        const uint32_t* k = keyx.data() + wkoff[x];
        bool gap = false;
        for (uint32_t i = 1; i < wklen[x]; i++) if (!(k[i] & TF_FLAG)) gap = true;
        if (gap) ngap++;
        reach_sum += (k[wklen[x] - 1] & ~TF_FLAG) - (k[0] & ~TF_FLAG) + 1;
      }
      path_ab.resize((size_t)ngap * gmax + 1); path_node.resize((size_t)ngap * gmax + 1); path_cont.resize((size_t)ngap * gmax + 1);
      ntrf.resize(reach_sum + 2 * (size_t)nedges + 2);
      TfG f; memset(&f, 0, sizeof(f));
      f.nt = nt0; f.cap = cap; f.koff = koff.data(); f.klen = klen.data(); f.key = keyx.data(); f.xkey0 = xkey0; f.ab = ab.data(); f.real = real.data();
      f.usepath = usepath.data(); f.path_off = path_off.data(); f.path_cnt = path_cnt.data(); f.path_node = path_node.data(); f.path_cont = path_cont.data();
      f.path_ab = path_ab.data(); f.path_cap = ngap * gmax; f.pat = pat.data(); f.ntrf_off = ntrf_off.data(); f.ntrf_cnt = ntrf_cnt.data(); f.ntrf = ntrf.data();
      f.ntrf_cap = (int)ntrf.size(); f.order = order.data(); f.w_koff = wkoff.data(); f.w_klen = wklen.data(); f.w_ab = wab.data(); f.skey = skey.data();
      f.e_cov = ecov.data();
      const int nt = process_transfrags_warp(g, f, &err);
      f.nt = nt;
      w_gs.push_back((int8_t)g_s[gi]); w_gb.push_back(g_b[gi]);
      for (int t = 0; t < nt; t++) {
        w_tf_ab.push_back(f.ab[t]); w_tf_real.push_back(f.real[t]);
        for (uint32_t i = 0; i < f.klen[t]; i++) w_tf_nodes.push_back(tf_node(f, t, (int)i));
        w_tf_node_off.push_back((uint32_t)w_tf_nodes.size());
        for (int x = 0; x < f.path_cnt[t]; x++) {
          w_path_node.push_back(f.path_node[f.path_off[t] + x]); w_path_cont.push_back(f.path_cont[f.path_off[t] + x]); w_path_ab.push_back(f.path_ab[f.path_off[t] + x]);
        }
        w_path_off.push_back((uint32_t)w_path_node.size());
      }
      w_tf_off.push_back((uint32_t)w_tf_ab.size());
      for (int i = 0; i < gno; i++) {
        for (int x = 0; x < f.ntrf_cnt[i]; x++) w_ntrf.push_back(f.ntrf[f.ntrf_off[i] + x]);
        w_ntrf_off.push_back((uint32_t)w_ntrf.size());
      }
      w_node_off.push_back((uint32_t)w_ntrf_off.size() - 1);
      if (stage < 15) continue;
      // ---- a13-a15
      std::vector<float> nodecov(gno), capl(gno + 1), capr(gno + 1), suml(gno + 1), sumr(gno + 1), nodeflux(gno + 1), ex_c(gno + 1);
      std::vector<int32_t> pathv(gno + 2), succ(gno), node2path(gno); std::vector<unsigned long long> pathpat(nwn), istr((cap + 63) / 64 + 1);
      std::vector<uint8_t> used(gno), prevn(gno), preve(nedges + 1); std::vector<uint32_t> ex_s(gno + 1), ex_e(gno + 1);
      FlowScratch S; memset(&S, 0, sizeof(S));
      S.nodecov = nodecov.data(); S.path = pathv.data(); S.succ = succ.data(); S.pathpat = pathpat.data(); S.istr = istr.data(); S.usednode = used.data();
      S.prev_node = prevn.data(); S.prev_edge = preve.data(); S.node2path = node2path.data(); S.capl = capl.data(); S.capr = capr.data(); S.suml = suml.data();
      S.sumr = sumr.data(); S.nodeflux = nodeflux.data();
      const uint32_t PC = 4096, EC = 65536;
      std::vector<int32_t> pg(PC), ps(PC), ptl(PC); std::vector<uint32_t> pst(PC), pen(PC), pe0(PC), pnx(PC), est(EC), een(EC); std::vector<double> pcv(PC);
      std::vector<float> ecv(EC); uint32_t ctr[2] = {0, 0};
      PredOut po; po.counters = ctr; po.cap_pred = PC; po.cap_exon = EC; po.p_graph = pg.data(); po.p_seq = ps.data(); po.p_start = pst.data(); po.p_end = pen.data();
      po.p_cov = pcv.data(); po.p_tlen = ptl.data(); po.p_exon0 = pe0.data(); po.p_nex = pnx.data(); po.e_start = est.data(); po.e_end = een.data(); po.e_cov = ecv.data();
      const int ns = find_transcripts_warp(g, f, S, P, po, (int)gi, ex_s.data(), ex_e.data(), ex_c.data(), &err);
      if (ns) geneno++;
      for (uint32_t x = 0; x < ctr[0]; x++) {
        h_start.push_back((int32_t)pst[x]); h_end.push_back((int32_t)pen[x]); h_cov.push_back(pcv[x]); h_tlen.push_back(ptl[x]); h_geneno.push_back(geneno - 1);
        h_strand.push_back((int8_t)(g_s[gi] ? '+' : '-'));
        for (uint32_t e = 0; e < pnx[x]; e++) { h_es.push_back(est[pe0[x] + e]); h_ee.push_back(een[pe0[x] + e]); h_ecov.push_back(ecv[pe0[x] + e]); }
        h_exon_off.push_back((uint32_t)h_es.size());
      }
    }
    w_graph_off.push_back((uint32_t)w_gs.size());
    h_pred_off.push_back((uint32_t)h_start.size());
  }
  if (stage >= 12) {
    out.add("a12_graph_off", STG_DT_U32, w_graph_off); out.add("a12_g_strand", STG_DT_I8, w_gs); out.add("a12_g_bundle", STG_DT_I32, w_gb);
    out.add("a12_tf_off", STG_DT_U32, w_tf_off); out.add("a12_tf_abund", STG_DT_F32, w_tf_ab); out.add("a12_tf_real", STG_DT_U8, w_tf_real);
    out.add("a12_tf_node_off", STG_DT_U32, w_tf_node_off); out.add("a12_tf_nodes", STG_DT_I32, w_tf_nodes); out.add("a12_tf_path_off", STG_DT_U32, w_path_off);
    out.add("a12_tf_path_node", STG_DT_I32, w_path_node); out.add("a12_tf_path_cont", STG_DT_I32, w_path_cont); out.add("a12_tf_path_ab", STG_DT_F32, w_path_ab);
    out.add("a12_node_off", STG_DT_U32, w_node_off); out.add("a12_ntrf_off", STG_DT_U32, w_ntrf_off); out.add("a12_ntrf", STG_DT_I32, w_ntrf);
  }
  if (stage >= 15) {
    out.add("hp_pred_off", STG_DT_U32, h_pred_off); out.add("hp_start", STG_DT_I32, h_start); out.add("hp_end", STG_DT_I32, h_end); out.add("hp_cov", STG_DT_F64, h_cov);
    out.add("hp_tlen", STG_DT_I32, h_tlen); out.add("hp_geneno", STG_DT_I32, h_geneno); out.add("hp_strand", STG_DT_I8, h_strand);
    out.add("hp_exon_off", STG_DT_U32, h_exon_off); out.add("hp_es", STG_DT_U32, h_es); out.add("hp_ee", STG_DT_U32, h_ee); out.add("hp_ecov", STG_DT_F32, h_ecov);
  }
  if (stage >= 11) {
    out.add("a11_graph_off", STG_DT_U32, q_graph_off); out.add("a11_node_off", STG_DT_U32, q_node_off); out.add("a11_n_cov", STG_DT_F32, q_ncov);
    out.add("a11_tf_off", STG_DT_U32, q_tf_off); out.add("a11_tf_abund", STG_DT_F32, q_tf_ab); out.add("a11_tf_node_off", STG_DT_U32, q_tf_node_off);
    out.add("a11_tf_nodes", STG_DT_I32, q_tf_nodes); out.add("a11_tf_pcount", STG_DT_I32, q_tf_pcount);
  }
  out.add("a10_graph_off", STG_DT_U32, o_graph_off); out.add("a10_g_strand", STG_DT_I8, o_gs); out.add("a10_g_bundle", STG_DT_I32, o_gb);
  out.add("a10_g_graphno", STG_DT_I32, o_gno); out.add("a10_g_edgeno", STG_DT_I32, o_edgeno); out.add("a10_node_off", STG_DT_U32, o_node_off);
  out.add("a10_n_start", STG_DT_U32, o_nstart); out.add("a10_n_end", STG_DT_U32, o_nend); out.add("a10_n_par_off", STG_DT_U32, o_par_off);
  out.add("a10_n_par", STG_DT_I32, o_par); out.add("a10_n_child_off", STG_DT_U32, o_child_off); out.add("a10_n_child", STG_DT_I32, o_child);
  out.add("a10_tf_off", STG_DT_U32, o_tf_off); out.add("a10_tf_abund", STG_DT_F32, o_tf_ab); out.add("a10_tf_node_off", STG_DT_U32, o_tf_node_off);
  out.add("a10_tf_nodes", STG_DT_I32, o_tf_nodes);
  out.write(argv[2]);
  if (err) fprintf(stderr, "asm_hostcheck: error flags 0x%x\n", err);
  return err ? 1 : 0;
}
