// host/stgpu_adaptor.cpp — the reference-side binding of libstgpu (made real).
//
// Built by host/Makefile into host/_build/stringtie_stgpu: the reference program (stringtie.cpp main, htslib decode,
// bundle queue, build_graphs, printResults — all its own sources, compiled where they lie) with ONE function replaced:
//     void count_good_junctions(BundleData* bdata)        rlink.cpp:16656, called from infer_transcripts (:17256)
// is provided here and runs on the GPU through the C ABI of include/stgpu.h.  For short reads that function is exactly
// the stage the library implements (add_read_to_cov + junction support + the blocked clamped scan; the long-read /
// mixed-mode junction correction and strand inference never run, rlink.cpp:16668, 16902-16975, 17067).  The rest of the
// reference then works on the bpcov and junction counters the GPU produced, so the GTF must come out byte-identical
// to the unmodified reference's — tests/test_host_integration.py checks exactly that.
//
// This is the one-bundle synchronous form (stg_infer_transcripts): a correctness demonstration of the boundary, not
// the fast path — a production host batches thousands of bundles per call.
#include "rlink.h"
#include "stgpu.h"
#include <mutex>
#include <vector>
#include <cstring>
#include <unistd.h>

extern uint junctionsupport;   // rlink.cpp:35 (stringtie.cpp:140)
extern uint sserror;           // rlink.cpp:36
extern int junctionthr;        // rlink.cpp:37
extern uint bundledist;
extern bool longreads, mixedMode, eonly, viral;

void count_good_junctions_reference(BundleData* bdata);   // the reference's own body, renamed by host/Makefile

namespace {
stg_ctx* g_ctx = nullptr;
std::mutex g_mu;          // one context, one stream: worker threads (-p N) take turns
long g_calls = 0;

stg_ctx* context() {
  if (g_ctx) return g_ctx;
  stg_params p;
  stg_params_default(&p);
  p.junctionsupport = junctionsupport; p.sserror = sserror; p.junctionthr = junctionthr; p.bundledist = bundledist;
  p.eonly = eonly ? 1 : 0;
  if (stg_init(&p, 0, &g_ctx) != STG_OK) {   // no GPU, no result: there is no CPU fallback behind this call
    fprintf(stderr, "libstgpu: %s\n", stg_last_error(nullptr));
    _exit(1);                                // (exit() would wait on the reference's worker threads)
  }
  atexit([] { if (getenv("STGPU_VERBOSE")) fprintf(stderr, "libstgpu: %ld bundles went through the GPU\n", g_calls); });
  return g_ctx;
}
}  // namespace

void count_good_junctions(BundleData* bdata) {
  if (longreads || mixedMode || viral) { count_good_junctions_reference(bdata); return; }   // stages not built yet
  GList<CReadAln>& rd = bdata->readlist;
  GList<CJunction>& jn = bdata->junction;
  const int nr = rd.Count(), nj = jn.Count();
  std::vector<uint32_t> r_start(nr), r_end(nr), r_len(nr), seg_off(nr + 1, 0), pair_off(nr + 1, 0), seg_s, seg_e;
  std::vector<float> r_count(nr), pair_cnt;
  std::vector<int16_t> r_nh(nr);
  std::vector<int8_t> r_strand(nr), j_strand(nj), j_gm(nj);
  std::vector<uint8_t> r_flags(nr);
  std::vector<int32_t> rjunc, pair_idx;
  std::vector<uint32_t> j_s(nj), j_e(nj);
  std::vector<double> j_nreads(nj), j_nm(nj), j_mm(nj);
  GHashMap<CJunction*, int> row(false);
  for (int j = 0; j < nj; j++) {                       // CJunction, bundle.h:141-166
    CJunction* x = jn[j];
    row.Add(x, j);
    j_s[j] = x->start; j_e[j] = x->end; j_strand[j] = x->strand; j_gm[j] = x->guide_match;
    j_nreads[j] = x->nreads; j_nm[j] = x->nm; j_mm[j] = x->mm;
  }
  for (int n = 0; n < nr; n++) {                       // CReadAln, bundle.h:169-230
    CReadAln& r = *rd[n];
    r_start[n] = r.start; r_end[n] = r.end; r_len[n] = r.len; r_count[n] = r.read_count; r_nh[n] = r.nh;
    r_strand[n] = r.strand;
    r_flags[n] = (uint8_t)((r.unitig ? STG_RF_UNITIG : 0) | (r.longread ? STG_RF_LONGREAD : 0));
    for (int i = 0; i < r.segs.Count(); i++) { seg_s.push_back(r.segs[i].start); seg_e.push_back(r.segs[i].end); }
    for (int i = 0; i < r.juncs.Count(); i++) {
      int* p = row.Find(r.juncs[i]);
      rjunc.push_back(p ? *p : -1);
    }
    for (int i = 0; i < r.pair_idx.Count(); i++) { pair_idx.push_back(r.pair_idx[i]); pair_cnt.push_back(r.pair_count[i]); }
    seg_off[n + 1] = (uint32_t)seg_s.size(); pair_off[n + 1] = (uint32_t)pair_idx.size();
  }
  if (rjunc.empty()) rjunc.push_back(0);
  stg_bundle_view v;
  memset(&v, 0, sizeof(v));
  v.start = bdata->start; v.end = bdata->end; v.n_reads = nr; v.n_juncs = nj;
  v.r_start = r_start.data(); v.r_end = r_end.data(); v.r_len = r_len.data(); v.r_count = r_count.data();
  v.r_nh = r_nh.data(); v.r_strand = r_strand.data(); v.r_flags = r_flags.data(); v.r_seg_off = seg_off.data();
  v.seg_start = seg_s.data(); v.seg_end = seg_e.data(); v.rjunc_idx = rjunc.data(); v.r_pair_off = pair_off.data();
  v.pair_idx = pair_idx.data(); v.pair_cnt = pair_cnt.data(); v.j_start = j_s.data(); v.j_end = j_e.data();
  v.j_strand = j_strand.data(); v.j_guide_match = j_gm.data(); v.j_nreads = j_nreads.data(); v.j_nm = j_nm.data();
  v.j_mm = j_mm.data();
  const int len = bdata->end - bdata->start + 3;       // rlink.cpp:16875
  std::vector<float> cov(3 * (size_t)len);
  std::vector<double> good(nj), left(nj), right(nj);
  stg_bundle_result res = {cov.data(), good.data(), left.data(), right.data()};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    stg_ctx* ctx = context();
    if (stg_infer_transcripts(ctx, &v, &res) != STG_OK) { fprintf(stderr, "libstgpu: %s\n", stg_last_error(ctx)); _exit(1); }
    g_calls++;
  }
  for (int s = 0; s < 3; s++) {                        // BundleData::bpcov, bundle.h:329
    bdata->bpcov[s].Resize(len);
    memcpy(&bdata->bpcov[s][0], &cov[(size_t)s * len], (size_t)len * sizeof(float));
  }
  for (int j = 0; j < nj; j++) {                       // the reference accumulates into fields that start at 0
    jn[j]->nreads_good += good[j]; jn[j]->leftsupport += left[j]; jn[j]->rightsupport += right[j];
  }
}
