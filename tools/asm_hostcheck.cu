// tools/asm_hostcheck.cu — DEVELOPMENT AID, not part of libstgpu: compiles the per-graph device code
// (stringtie_b200/csrc/asm_device.cuh, frag_device.cuh) for the HOST (one-lane warps, warp_prims.cuh) and steps through
// it on a dump of the reference (`oracle/_ref/ref_hot --full --dump`, or a golden fixture converted by
// tools/asm_hostcheck.py), starting from the reference's OWN state after the bundles were formed (a9 / a9b arrays), so
// that the graph / transfrag / flow code can be debugged on a CPU-only box against the reference's stage dumps.
// Writes what it computed as a10_* / a11_* / a12_* / hp_* arrays for tools/asm_hostcheck.py to compare.
//   nvcc -std=c++17 -O1 -Xcompiler -ffp-contract=off -o tools/_build/asm_hostcheck tools/asm_hostcheck.cu
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
    std::vector<BundleGraphs> BG(1);
    BundleGraphs& bg = BG[0];
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
    (void)ng; (void)r0; (void)nr; (void)g2b_all; (void)negp_all; (void)merge_all; (void)rg_off; (void)rg_all; (void)r_start; (void)r_count;
    (void)r_nh; (void)r_flags; (void)r_seg_off; (void)r_pair_off; (void)pair_cnt; (void)a9_strand; (void)a9_pair; (void)a9_seg_s; (void)a9_seg_e;
    (void)a9_rjunc; (void)bg; (void)first_graph;
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
