// host/stgpu_adaptor.cpp — the reference-side binding of libstgpu (made real).
//
// Built by host/Makefile into host/_build/stringtie_stgpu: the reference program (stringtie.cpp main, htslib decode,
// processRead, the GThreads bundle queue, printResults, the GTF writer — all its own sources, compiled where they lie)
// with ONE function replaced:
//     int infer_transcripts(BundleData* bundle)        rlink.h:380, body rlink.cpp:17161, call site stringtie.cpp:1477
// is provided here and runs on the GPU through the C ABI of include/stgpu.h: the bundle is flattened into a
// stg_bundle_view, stg_infer_transcripts(assemble = 1) runs count_good_junctions and build_graphs to its end on the
// device, and what the reference function writes back — BundleData::pred, bpcov, num_fragments, frag_len — is put
// back into the BundleData, so that the reference's own printResults / FPKM-TPM pass work on it unchanged.  The GTF
// must come out byte-identical to the unmodified reference's: tests/test_host_integration.py checks exactly that.
// Modes the library does not implement yet (guides / -e, -L, --mix, -N, --merge, --viral, --ptf) stop with an error:
// there is no CPU path behind this call.
//
// This is the one-bundle synchronous form: a correctness demonstration of the boundary, not the fast path — a
// production host batches thousands of bundles per call.
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
extern int mintranscriptlen, allowed_nodes;
extern float isofrac, mcov, readthr, singlethr;
extern bool longreads, mixedMode, eonly, viral, trim, includesource, isnascent, printNascent, rawreads, guided, mergeMode, havePtFeatures;

namespace {
stg_ctx* g_ctx = nullptr;
std::mutex g_mu;          // one context, one stream: worker threads (-p N) take turns
long g_calls = 0, g_preds = 0;

void die(const char* what) {
  fprintf(stderr, "libstgpu: %s\n", what);
  _exit(1);                                  // (exit() would wait on the reference's worker threads)
}

stg_ctx* context() {
  if (g_ctx) return g_ctx;
  stg_params p;
  stg_params_default(&p);
  p.junctionsupport = junctionsupport; p.sserror = sserror; p.junctionthr = junctionthr; p.bundledist = bundledist;
  p.mintranscriptlen = mintranscriptlen; p.allowed_nodes = allowed_nodes; p.isofrac = isofrac; p.mcov = mcov; p.readthr = readthr;
  p.singlethr = singlethr; p.trim = trim ? 1 : 0; p.includesource = includesource ? 1 : 0;
  p.eonly = eonly ? 1 : 0; p.longreads = longreads ? 1 : 0; p.mixedMode = mixedMode ? 1 : 0; p.viral = viral ? 1 : 0;
  p.isnascent = isnascent ? 1 : 0; p.printNascent = printNascent ? 1 : 0; p.rawreads = rawreads ? 1 : 0; p.guided = guided ? 1 : 0;
  p.havePtFeatures = havePtFeatures ? 1 : 0;
  if (stg_init(&p, 0, &g_ctx) != STG_OK) die(stg_last_error(nullptr));   // no GPU, no result
  atexit([] { if (getenv("STGPU_VERBOSE")) fprintf(stderr, "libstgpu: %ld bundles went through the GPU, %ld predictions came back\n", g_calls, g_preds); });
  return g_ctx;
}
}  // namespace

int infer_transcripts(BundleData* bundle) {
  if (mergeMode) die("--merge is not implemented by libstgpu");
  if (bundle->keepguides.Count() == 0 && eonly) return 0;   // rlink.cpp:17177: nothing to do for this bundle
  GList<CReadAln>& rd = bundle->readlist;
  GList<CJunction>& jn = bundle->junction;
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
  v.start = bundle->start; v.end = bundle->end; v.n_reads = nr; v.n_juncs = nj;
  v.r_start = r_start.data(); v.r_end = r_end.data(); v.r_len = r_len.data(); v.r_count = r_count.data();
  v.r_nh = r_nh.data(); v.r_strand = r_strand.data(); v.r_flags = r_flags.data(); v.r_seg_off = seg_off.data();
  v.seg_start = seg_s.data(); v.seg_end = seg_e.data(); v.rjunc_idx = rjunc.data(); v.r_pair_off = pair_off.data();
  v.pair_idx = pair_idx.data(); v.pair_cnt = pair_cnt.data(); v.j_start = j_s.data(); v.j_end = j_e.data();
  v.j_strand = j_strand.data(); v.j_guide_match = j_gm.data(); v.j_nreads = j_nreads.data(); v.j_nm = j_nm.data();
  v.j_mm = j_mm.data();
  const int len = bundle->end - bundle->start + 3;     // rlink.cpp:16875
  std::vector<float> cov(3 * (size_t)len);
  stg_bundle_result res;
  memset(&res, 0, sizeof(res));
  res.bpcov = cov.data();
  res.assemble = 1;
  int geneno = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    stg_ctx* ctx = context();
    if (stg_infer_transcripts(ctx, &v, &res) != STG_OK) die(stg_last_error(ctx));
    g_calls++; g_preds += res.n_pred;
    // BundleData::pred (CPrediction, bundle.h:270-311) — copied out while the context's buffers are ours
    for (int i = 0; i < res.n_pred; i++) {
      CPrediction* p = new CPrediction(res.p_geneno[i], NULL, res.p_start[i], res.p_end[i], res.p_cov[i], (char)res.p_strand[i], res.p_tlen[i]);
      for (uint32_t e = res.p_exon_off[i]; e < res.p_exon_off[i + 1]; e++) {
        GSeg exon(res.e_start[e], res.e_end[e]);
        p->exons.Add(exon);
        float ec = res.e_cov[e];
        p->exoncov.Add(ec);
      }
      bundle->pred.Add(p);
      if (res.p_geneno[i] + 1 > geneno) geneno = res.p_geneno[i] + 1;
    }
  }
  for (int s = 0; s < 3; s++) {                        // BundleData::bpcov, bundle.h:329 (printResults reads it)
    bundle->bpcov[s].Resize(len);
    memcpy(&bundle->bpcov[s][0], &cov[(size_t)s * len], (size_t)len * sizeof(float));
  }
  bundle->num_fragments += res.num_fragments;         // rlink.cpp:15105-15116 (both start at 0 for a fresh bundle)
  bundle->frag_len += res.frag_len;
  return geneno;
}
