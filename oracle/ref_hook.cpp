// oracle/ref_hook.cpp — TEST INFRASTRUCTURE ONLY (header per task rules: nothing under oracle/ is product code;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may run it).
//
// Compiled ONLY into oracle/_ref/ref_tool by oracle/Makefile.  It pulls the reference's own translation unit
// in by #include from where it lies under /root/reference (nothing is copied into this repo) with
// infer_transcripts() renamed, and supplies:
//   * a hooked infer_transcripts() (rlink.h:380) that serialises every bundle crossing the boundary as a
//     packed batch (include/stgpu.h schema) plus stage outputs: bpcov[3] and the junction table after
//     count_good_junctions() (rlink.cpp:16656), and pred / frag_len / num_fragments at return;
//   * `STG_DUMP=<file> ref_tool <stringtie args…>` — the reference main(), unchanged, with the hook armed;
//   * `ref_hot <in.stgb> [--threads N] [--reps R] [--dump out.stgb] [--limit nbundles]` (built from this file
//     with -DSTG_HOOK_MAIN) — rebuilds
//     BundleData from a packed batch and runs the reference's OWN count_good_junctions() on it: the
//     reference-side oracle for arbitrary packed inputs and the "--impl reference" timing arm.
#define infer_transcripts infer_transcripts_reference
#include "rlink.cpp"
#undef infer_transcripts

#include "../include/stgpu.h"
#include <vector>
#include <string>
#include <unordered_map>
#include <thread>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>


namespace {

struct Arr {
  std::string name;
  uint32_t dtype;
  size_t esize;
  std::vector<char> bytes;
  Arr(const char* n, uint32_t dt, size_t es) : name(n), dtype(dt), esize(es) {}
  template <class T> void push(T v) {
    const char* p = (const char*)&v;
    bytes.insert(bytes.end(), p, p + sizeof(T));
  }
  size_t count() const { return bytes.size() / esize; }
};

struct Dump {
  std::mutex mu;
#define A(name, dt, T) Arr name{#name, dt, sizeof(T)};
  A(b_start, STG_DT_I32, int32_t) A(b_end, STG_DT_I32, int32_t)
  A(b_read_off, STG_DT_U32, uint32_t) A(b_junc_off, STG_DT_U32, uint32_t)
  A(r_start, STG_DT_U32, uint32_t) A(r_end, STG_DT_U32, uint32_t) A(r_len, STG_DT_U32, uint32_t)
  A(r_count, STG_DT_F32, float) A(r_nh, STG_DT_I16, int16_t) A(r_strand, STG_DT_I8, int8_t)
  A(r_flags, STG_DT_U8, uint8_t) A(r_seg_off, STG_DT_U32, uint32_t) A(r_pair_off, STG_DT_U32, uint32_t)
  A(seg_start, STG_DT_U32, uint32_t) A(seg_end, STG_DT_U32, uint32_t) A(rjunc_idx, STG_DT_I32, int32_t)
  A(pair_idx, STG_DT_I32, int32_t) A(pair_cnt, STG_DT_F32, float)
  A(j_start, STG_DT_U32, uint32_t) A(j_end, STG_DT_U32, uint32_t) A(j_strand, STG_DT_I8, int8_t)
  A(j_guide_match, STG_DT_I8, int8_t) A(j_nreads, STG_DT_F64, double) A(j_nm, STG_DT_F64, double)
  A(j_mm, STG_DT_F64, double)
  // stage 1: after count_good_junctions
  A(s1_cov_off, STG_DT_I64, int64_t) A(s1_cov, STG_DT_F32, float)
  A(s1_j_nreads_good, STG_DT_F64, double) A(s1_j_leftsupport, STG_DT_F64, double)
  A(s1_j_rightsupport, STG_DT_F64, double) A(s1_j_nreads, STG_DT_F64, double) A(s1_j_strand, STG_DT_I8, int8_t)
  A(b_cov_stage, STG_DT_U8, uint8_t) A(b_cov_final_same, STG_DT_U8, uint8_t)
  // final: at return of infer_transcripts
  A(f_num_fragments, STG_DT_F64, double) A(f_frag_len, STG_DT_F64, double)
  A(f_pred_off, STG_DT_U32, uint32_t) A(p_start, STG_DT_I32, int32_t) A(p_end, STG_DT_I32, int32_t)
  A(p_cov, STG_DT_F64, double) A(p_longcov, STG_DT_F64, double) A(p_strand, STG_DT_I8, int8_t)
  A(p_tlen, STG_DT_I32, int32_t) A(p_flag, STG_DT_U8, uint8_t) A(p_geneno, STG_DT_I32, int32_t)
  A(p_exon_off, STG_DT_U32, uint32_t) A(pe_start, STG_DT_U32, uint32_t) A(pe_end, STG_DT_U32, uint32_t)
  A(pe_cov, STG_DT_F32, float)
  A(f_junc_off, STG_DT_U32, uint32_t) A(f_j_start, STG_DT_U32, uint32_t) A(f_j_end, STG_DT_U32, uint32_t)
  A(f_j_strand, STG_DT_I8, int8_t) A(f_j_nreads, STG_DT_F64, double) A(f_j_nreads_good, STG_DT_F64, double)
  A(f_j_leftsupport, STG_DT_F64, double) A(f_j_rightsupport, STG_DT_F64, double) A(f_j_mm, STG_DT_F64, double)
#undef A
  std::vector<Arr*> all() {
    return {&b_start, &b_end, &b_read_off, &b_junc_off, &r_start, &r_end, &r_len, &r_count, &r_nh, &r_strand,
            &r_flags, &r_seg_off, &r_pair_off, &seg_start, &seg_end, &rjunc_idx, &pair_idx, &pair_cnt, &j_start,
            &j_end, &j_strand, &j_guide_match, &j_nreads, &j_nm, &j_mm, &s1_cov_off, &s1_cov, &s1_j_nreads_good,
            &s1_j_leftsupport, &s1_j_rightsupport, &s1_j_nreads, &s1_j_strand, &b_cov_stage, &b_cov_final_same,
            &f_num_fragments, &f_frag_len, &f_pred_off, &p_start, &p_end, &p_cov, &p_longcov, &p_strand, &p_tlen,
            &p_flag, &p_geneno, &p_exon_off, &pe_start, &pe_end, &pe_cov, &f_junc_off, &f_j_start, &f_j_end,
            &f_j_strand, &f_j_nreads, &f_j_nreads_good, &f_j_leftsupport, &f_j_rightsupport, &f_j_mm};
  }
  bool started = false;
  void begin() {
    if (started) return;
    started = true;
    b_read_off.push<uint32_t>(0); b_junc_off.push<uint32_t>(0); r_seg_off.push<uint32_t>(0);
    r_pair_off.push<uint32_t>(0); s1_cov_off.push<int64_t>(0); f_pred_off.push<uint32_t>(0);
    p_exon_off.push<uint32_t>(0); f_junc_off.push<uint32_t>(0);
  }
  int write(const char* path) {
    begin();
    FILE* f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "ref_tool: cannot write %s\n", path); return 1; }
    std::vector<Arr*> v = all();
    uint32_t hdr[3] = {STG_FILE_MAGIC, 1u, (uint32_t)v.size()};
    fwrite(hdr, 4, 3, f);
    for (Arr* a : v) {
      char name[24]; memset(name, 0, 24); strncpy(name, a->name.c_str(), 23);
      fwrite(name, 1, 24, f);
      uint32_t dt[2] = {a->dtype, 0}; fwrite(dt, 4, 2, f);
      uint64_t cnt = a->count(); fwrite(&cnt, 8, 1, f);
      if (!a->bytes.empty()) fwrite(a->bytes.data(), 1, a->bytes.size(), f);
      size_t pad = (8 - a->bytes.size() % 8) % 8; char z[8] = {0};
      if (pad) fwrite(z, 1, pad, f);
    }
    fclose(f);
    return 0;
  }
};

Dump* g_dump = NULL;
const char* g_dump_path = NULL;

// ---- serialise the boundary input (what BundleData holds when infer_transcripts is entered) ----------
void dump_input(Dump& d, BundleData* b) {
  d.begin();
  d.b_start.push<int32_t>(b->start); d.b_end.push<int32_t>(b->end);
  std::unordered_map<const CJunction*, int> jidx;
  for (int j = 0; j < b->junction.Count(); j++) {
    CJunction* jn = b->junction[j];
    jidx[jn] = j;
    d.j_start.push<uint32_t>(jn->start); d.j_end.push<uint32_t>(jn->end);
    d.j_strand.push<int8_t>(jn->strand); d.j_guide_match.push<int8_t>(jn->guide_match);
    d.j_nreads.push<double>(jn->nreads); d.j_nm.push<double>(jn->nm); d.j_mm.push<double>(jn->mm);
  }
  d.b_junc_off.push<uint32_t>((uint32_t)d.j_start.count());
  for (int n = 0; n < b->readlist.Count(); n++) {
    CReadAln& rd = *(b->readlist[n]);
    d.r_start.push<uint32_t>(rd.start); d.r_end.push<uint32_t>(rd.end); d.r_len.push<uint32_t>(rd.len);
    d.r_count.push<float>(rd.read_count); d.r_nh.push<int16_t>(rd.nh); d.r_strand.push<int8_t>(rd.strand);
    d.r_flags.push<uint8_t>((rd.unitig ? STG_RF_UNITIG : 0) | (rd.longread ? STG_RF_LONGREAD : 0));
    if (rd.juncs.Count() != rd.segs.Count() - 1) {
      fprintf(stderr, "ref_tool: read %d has %d segs but %d juncs\n", n, rd.segs.Count(), rd.juncs.Count());
      exit(3);
    }
    for (int i = 0; i < rd.segs.Count(); i++) {
      d.seg_start.push<uint32_t>(rd.segs[i].start); d.seg_end.push<uint32_t>(rd.segs[i].end);
    }
    for (int i = 0; i < rd.juncs.Count(); i++) {
      std::unordered_map<const CJunction*, int>::iterator it = jidx.find(rd.juncs[i]);
      d.rjunc_idx.push<int32_t>(it == jidx.end() ? -1 : it->second);
    }
    for (int i = 0; i < rd.pair_idx.Count(); i++) {
      d.pair_idx.push<int32_t>(rd.pair_idx[i]); d.pair_cnt.push<float>(rd.pair_count[i]);
    }
    d.r_seg_off.push<uint32_t>((uint32_t)d.seg_start.count());
    d.r_pair_off.push<uint32_t>((uint32_t)d.pair_idx.count());
  }
  d.b_read_off.push<uint32_t>((uint32_t)d.r_start.count());
}

void dump_cov(Dump& d, BundleData* b, int stage) {
  for (int s = 0; s < 3; s++)
    for (int i = 0; i < b->bpcov[s].Count(); i++) d.s1_cov.push<float>(b->bpcov[s][i]);
  d.s1_cov_off.push<int64_t>((int64_t)d.s1_cov.count());
  d.b_cov_stage.push<uint8_t>((uint8_t)stage);
}

void dump_stage1_junctions(Dump& d, BundleData* b) {
  for (int j = 0; j < b->junction.Count(); j++) {
    CJunction* jn = b->junction[j];
    d.s1_j_nreads_good.push<double>(jn->nreads_good); d.s1_j_leftsupport.push<double>(jn->leftsupport);
    d.s1_j_rightsupport.push<double>(jn->rightsupport); d.s1_j_nreads.push<double>(jn->nreads);
    d.s1_j_strand.push<int8_t>(jn->strand);
  }
}

void dump_final(Dump& d, BundleData* b) {
  d.f_num_fragments.push<double>(b->num_fragments); d.f_frag_len.push<double>(b->frag_len);
  for (int p = 0; p < b->pred.Count(); p++) {
    CPrediction* pr = b->pred[p];
    d.p_start.push<int32_t>(pr->start); d.p_end.push<int32_t>(pr->end); d.p_cov.push<double>(pr->cov);
    d.p_longcov.push<double>(pr->longcov); d.p_strand.push<int8_t>(pr->strand); d.p_tlen.push<int32_t>(pr->tlen);
    d.p_flag.push<uint8_t>(pr->flag ? 1 : 0); d.p_geneno.push<int32_t>(pr->geneno);
    for (int e = 0; e < pr->exons.Count(); e++) {
      d.pe_start.push<uint32_t>(pr->exons[e].start); d.pe_end.push<uint32_t>(pr->exons[e].end);
      d.pe_cov.push<float>(e < pr->exoncov.Count() ? pr->exoncov[e] : 0.f);
    }
    d.p_exon_off.push<uint32_t>((uint32_t)d.pe_start.count());
  }
  d.f_pred_off.push<uint32_t>((uint32_t)d.p_start.count());
  for (int j = 0; j < b->junction.Count(); j++) {
    CJunction* jn = b->junction[j];
    d.f_j_start.push<uint32_t>(jn->start); d.f_j_end.push<uint32_t>(jn->end); d.f_j_strand.push<int8_t>(jn->strand);
    d.f_j_nreads.push<double>(jn->nreads); d.f_j_nreads_good.push<double>(jn->nreads_good);
    d.f_j_leftsupport.push<double>(jn->leftsupport); d.f_j_rightsupport.push<double>(jn->rightsupport);
    d.f_j_mm.push<double>(jn->mm);
  }
  d.f_junc_off.push<uint32_t>((uint32_t)d.f_j_start.count());
}

}  // namespace

// ---- the hooked boundary function --------------------------------------------------------------------
// In plain short-read mode the body of the reference's infer_transcripts (rlink.cpp:17161-17272) reduces to
// count_good_junctions(bundle); build_graphs(bundle); (the long-read exon shortening and the mixed-mode
// re-sort never trigger), so the two calls are made here with a dump in between.  In every other mode the
// renamed reference function is called whole and only input + final state are dumped.
int infer_transcripts(BundleData* bundle) {
  if (!g_dump) return infer_transcripts_reference(bundle);
  std::lock_guard<std::mutex> lk(g_dump->mu);
  Dump& d = *g_dump;
  dump_input(d, bundle);
  int geneno = 0;
  bool split = !mergeMode && !longreads && !mixedMode && (bundle->keepguides.Count() || !eonly);
  if (split) {
    bool anylong = false;
    for (int n = 0; n < bundle->readlist.Count(); n++) if (bundle->readlist[n]->longread) { anylong = true; break; }
    if (anylong) split = false;
  }
  if (split) {
    count_good_junctions(bundle);
    dump_cov(d, bundle, 1);
    dump_stage1_junctions(d, bundle);
    std::vector<float> snap[3];
    for (int s = 0; s < 3; s++) snap[s].assign(&bundle->bpcov[s][0], &bundle->bpcov[s][0] + bundle->bpcov[s].Count());
    geneno = build_graphs(bundle);
    bool same = true;
    for (int s = 0; s < 3 && same; s++) {
      if ((int)snap[s].size() != bundle->bpcov[s].Count()) same = false;
      else if (memcmp(snap[s].data(), &bundle->bpcov[s][0], snap[s].size() * sizeof(float)) != 0) same = false;
    }
    d.b_cov_final_same.push<uint8_t>(same ? 1 : 0);
  } else {
    geneno = infer_transcripts_reference(bundle);
    dump_cov(d, bundle, 2);
    dump_stage1_junctions(d, bundle);
    d.b_cov_final_same.push<uint8_t>(1);
  }
  dump_final(d, bundle);
  return geneno;
}

// ---- .stgb reader for `hot` ---------------------------------------------------------------------------
namespace {
struct FileArr { uint32_t dtype; uint64_t count; std::vector<char> bytes; };
typedef std::unordered_map<std::string, FileArr> FileMap;

bool read_stgb(const char* path, FileMap& m) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  uint32_t hdr[3];
  if (fread(hdr, 4, 3, f) != 3 || hdr[0] != STG_FILE_MAGIC) { fclose(f); return false; }
  static const size_t es[] = {0, 1, 1, 2, 4, 4, 4, 8, 8};
  for (uint32_t i = 0; i < hdr[2]; i++) {
    char name[25]; name[24] = 0; uint32_t dt[2]; uint64_t cnt;
    if (fread(name, 1, 24, f) != 24 || fread(dt, 4, 2, f) != 2 || fread(&cnt, 8, 1, f) != 1) { fclose(f); return false; }
    FileArr a; a.dtype = dt[0]; a.count = cnt;
    size_t nb = cnt * es[dt[0]];
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
  if (it == m.end()) { fprintf(stderr, "ref_tool: array %s missing\n", name); exit(2); }
  return (const T*)it->second.bytes.data();
}
size_t arrn(const FileMap& m, const char* name) {
  FileMap::const_iterator it = m.find(name);
  return it == m.end() ? 0 : (size_t)it->second.count;
}

struct Batch {
  size_t nb;
  const int32_t *b_start, *b_end; const uint32_t *b_read_off, *b_junc_off;
  const uint32_t *r_start, *r_end, *r_len; const float* r_count; const int16_t* r_nh; const int8_t* r_strand;
  const uint8_t* r_flags; const uint32_t *r_seg_off, *r_pair_off, *seg_start, *seg_end; const int32_t* rjunc_idx;
  const int32_t* pair_idx; const float* pair_cnt; const uint32_t *j_start, *j_end; const int8_t *j_strand, *j_guide_match;
  const double *j_nreads, *j_nm, *j_mm;
};

// Rebuild what the producer (processRead, rlink.cpp:659-983) would have left in BundleData.
void build_bundle(const Batch& B, size_t b, BundleData& bd) {
  bd.start = B.b_start[b]; bd.end = B.b_end[b];
  uint32_t j0 = B.b_junc_off[b], j1 = B.b_junc_off[b + 1];
  bd.junction.setSorted(false);  // keep the dumped row order verbatim
  std::vector<CJunction*> jp(j1 - j0);
  for (uint32_t j = j0; j < j1; j++) {
    CJunction* jn = new CJunction(B.j_start[j], B.j_end[j], B.j_strand[j]);
    jn->guide_match = B.j_guide_match[j]; jn->nreads = B.j_nreads[j]; jn->nm = B.j_nm[j]; jn->mm = B.j_mm[j];
    bd.junction.Add(jn);
    jp[j - j0] = jn;
  }
  uint32_t r0 = B.b_read_off[b], r1 = B.b_read_off[b + 1];
  for (uint32_t n = r0; n < r1; n++) {
    CReadAln* rd = new CReadAln(B.r_strand[n], B.r_nh[n], B.r_start[n], B.r_end[n]);
    rd->len = B.r_len[n]; rd->read_count = B.r_count[n];
    rd->unitig = (B.r_flags[n] & STG_RF_UNITIG) != 0; rd->longread = (B.r_flags[n] & STG_RF_LONGREAD) != 0;
    uint32_t s0 = B.r_seg_off[n], s1 = B.r_seg_off[n + 1];
    for (uint32_t s = s0; s < s1; s++) { GSeg sg(B.seg_start[s], B.seg_end[s]); rd->segs.Add(sg); }
    for (uint32_t i = 0; i + 1 < s1 - s0; i++) {
      int32_t ji = B.rjunc_idx[s0 - n + i];
      rd->juncs.Add(ji >= 0 ? jp[ji] : jp[0]);
    }
    for (uint32_t p = B.r_pair_off[n]; p < B.r_pair_off[n + 1]; p++) {
      int pi = B.pair_idx[p]; float pc = B.pair_cnt[p];
      rd->pair_idx.Add(pi); rd->pair_count.Add(pc);
    }
    bd.readlist.Add(rd);
  }
}

int hot_main(int argc, char** argv) {
  const char* in = NULL; const char* dump = NULL; int threads = 1, reps = 1; long limit = -1;
  for (int i = 0; i < argc; i++) {
    if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];
    else if (!strcmp(argv[i], "--limit") && i + 1 < argc) limit = atol(argv[++i]);
    else in = argv[i];
  }
  if (!in) { fprintf(stderr, "usage: ref_hot <in.stgb> [--threads N] [--reps R] [--dump out.stgb] [--limit n]\n"); return 2; }
  FileMap m;
  if (!read_stgb(in, m)) { fprintf(stderr, "ref_tool: cannot read %s\n", in); return 2; }
  Batch B;
  B.nb = arrn(m, "b_start");
  if (limit >= 0 && (size_t)limit < B.nb) B.nb = (size_t)limit;
  B.b_start = arr<int32_t>(m, "b_start"); B.b_end = arr<int32_t>(m, "b_end");
  B.b_read_off = arr<uint32_t>(m, "b_read_off"); B.b_junc_off = arr<uint32_t>(m, "b_junc_off");
  B.r_start = arr<uint32_t>(m, "r_start"); B.r_end = arr<uint32_t>(m, "r_end"); B.r_len = arr<uint32_t>(m, "r_len");
  B.r_count = arr<float>(m, "r_count"); B.r_nh = arr<int16_t>(m, "r_nh"); B.r_strand = arr<int8_t>(m, "r_strand");
  B.r_flags = arr<uint8_t>(m, "r_flags"); B.r_seg_off = arr<uint32_t>(m, "r_seg_off");
  B.r_pair_off = arr<uint32_t>(m, "r_pair_off"); B.seg_start = arr<uint32_t>(m, "seg_start");
  B.seg_end = arr<uint32_t>(m, "seg_end"); B.rjunc_idx = arr<int32_t>(m, "rjunc_idx");
  B.pair_idx = arr<int32_t>(m, "pair_idx"); B.pair_cnt = arr<float>(m, "pair_cnt");
  B.j_start = arr<uint32_t>(m, "j_start"); B.j_end = arr<uint32_t>(m, "j_end"); B.j_strand = arr<int8_t>(m, "j_strand");
  B.j_guide_match = arr<int8_t>(m, "j_guide_match"); B.j_nreads = arr<double>(m, "j_nreads");
  B.j_nm = arr<double>(m, "j_nm"); B.j_mm = arr<double>(m, "j_mm");

  Dump* d = dump ? new Dump() : NULL;
  if (d && threads != 1) { fprintf(stderr, "ref_tool: --dump forces --threads 1\n"); threads = 1; }
  std::vector<double> tsec(threads, 0.0);
  std::vector<double> treads(threads, 0.0);
  auto w0 = std::chrono::steady_clock::now();
  auto worker = [&](int t) {
    for (int rep = 0; rep < reps; rep++) {
      for (size_t b = t; b < B.nb; b += threads) {
        BundleData bd;
        build_bundle(B, b, bd);
        if (d && rep == 0) dump_input(*d, &bd);
        auto t0 = std::chrono::steady_clock::now();
        count_good_junctions(&bd);
        auto t1 = std::chrono::steady_clock::now();
        tsec[t] += std::chrono::duration<double>(t1 - t0).count();
        treads[t] += (double)(B.b_read_off[b + 1] - B.b_read_off[b]);
        if (d && rep == 0) { dump_cov(*d, &bd, 1); dump_stage1_junctions(*d, &bd); d->b_cov_final_same.push<uint8_t>(1); }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
  double maxt = 0, sumreads = 0, agg = 0;
  for (int t = 0; t < threads; t++) {
    if (tsec[t] > maxt) maxt = tsec[t];
    sumreads += treads[t];
    if (tsec[t] > 0) agg += treads[t] / tsec[t];
  }
  printf("{\"tool\": \"ref_tool hot\", \"stage\": \"count_good_junctions\", \"bundles\": %zu, \"reads\": %.0f, "
         "\"threads\": %d, \"reps\": %d, \"compute_s_max\": %.6f, \"wall_s\": %.6f, \"reads_per_s\": %.1f}\n",
         B.nb, sumreads / reps, threads, reps, maxt, wall, agg);
  if (d) return d->write(dump);
  return 0;
}
}  // namespace

#ifdef STG_HOOK_MAIN
// ref_hot: our own main(); the reference main() is linked in renamed and never called (it has no final
// `return`, so it cannot be called as an ordinary function).
int main(int argc, char* argv[]) { return hot_main(argc - 1, argv + 1); }
#else
// ref_tool: the reference's own main() runs unchanged; the dump is armed by env STG_DUMP before main and
// written by an atexit handler after main returns.
namespace {
void write_dump_at_exit() {
  if (g_dump && g_dump_path) g_dump->write(g_dump_path);
}
struct ArmDump {
  ArmDump() {
    g_dump_path = getenv("STG_DUMP");
    if (g_dump_path && *g_dump_path) { g_dump = new Dump(); atexit(write_dump_at_exit); }
  }
} g_arm_dump;
}  // namespace
#endif
