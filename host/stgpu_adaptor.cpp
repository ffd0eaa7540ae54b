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
// Calls are COALESCED: the reference runs infer_transcripts on its -p N worker threads at once; every
// caller flattens its own bundle (in parallel), queues it and the first one to find no batch in flight becomes the leader:
// it waits until the callers that are still flattening have queued too, sends all queued bundles to the GPU as ONE device
// batch (stg_infer_transcripts_batch) and hands every waiter its result.  Bundles that arrive meanwhile form the next batch.
// With -p 1 every batch holds one bundle.  STGPU_VERBOSE=1 prints the batch statistics and the time spent flattening.
#include "rlink.h"
#include "stgpu.h"
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <ctime>
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
long g_calls = 0, g_preds = 0, g_batches = 0, g_maxbatch = 0;
double g_flatten_s = 0, g_gpu_s = 0;

void die(const char* what) {
  fprintf(stderr, "libstgpu: %s\n", what);
  _exit(1);                                  // (exit() would wait on the reference's worker threads)
}

double now_s() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

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
  atexit([] {
    if (getenv("STGPU_VERBOSE"))
      fprintf(stderr, "libstgpu: %ld bundles in %ld device batches (largest %ld), %ld predictions; flattening %.3f s (summed over threads), "
                      "GPU calls %.3f s\n", g_calls, g_batches, g_maxbatch, g_preds, g_flatten_s, g_gpu_s);
  });
  return g_ctx;
}

// one caller's bundle: the flattened BundleData (stg_bundle_view points into these vectors) and what comes back
struct Request {
  BundleData* bundle;
  std::vector<uint32_t> r_start, r_end, r_len, seg_off, pair_off, seg_s, seg_e, j_s, j_e;
  std::vector<float> r_count, pair_cnt, cov;
  std::vector<int16_t> r_nh;
  std::vector<int8_t> r_strand, j_strand, j_gm;
  std::vector<uint8_t> r_flags;
  std::vector<int32_t> rjunc, pair_idx;
  std::vector<double> j_nreads, j_nm, j_mm;
  int len = 0, geneno = 0;
  bool done = false;
};

void flatten(BundleData* bundle, Request& q, stg_bundle_view& v) {
  GList<CReadAln>& rd = bundle->readlist;
  GList<CJunction>& jn = bundle->junction;
  const int nr = rd.Count(), nj = jn.Count();
  q.bundle = bundle;
  q.r_start.resize(nr); q.r_end.resize(nr); q.r_len.resize(nr); q.seg_off.assign(nr + 1, 0); q.pair_off.assign(nr + 1, 0);
  q.r_count.resize(nr); q.r_nh.resize(nr); q.r_strand.resize(nr); q.r_flags.resize(nr);
  q.j_s.resize(nj); q.j_e.resize(nj); q.j_strand.resize(nj); q.j_gm.resize(nj); q.j_nreads.resize(nj); q.j_nm.resize(nj); q.j_mm.resize(nj);
  GHashMap<CJunction*, int> row(false);
  for (int j = 0; j < nj; j++) {                       // CJunction, bundle.h:141-166
    CJunction* x = jn[j];
    row.Add(x, j);
    q.j_s[j] = x->start; q.j_e[j] = x->end; q.j_strand[j] = x->strand; q.j_gm[j] = x->guide_match;
    q.j_nreads[j] = x->nreads; q.j_nm[j] = x->nm; q.j_mm[j] = x->mm;
  }
  for (int n = 0; n < nr; n++) {                       // CReadAln, bundle.h:169-230
    CReadAln& r = *rd[n];
    q.r_start[n] = r.start; q.r_end[n] = r.end; q.r_len[n] = r.len; q.r_count[n] = r.read_count; q.r_nh[n] = r.nh;
    q.r_strand[n] = r.strand;
    q.r_flags[n] = (uint8_t)((r.unitig ? STG_RF_UNITIG : 0) | (r.longread ? STG_RF_LONGREAD : 0));
    for (int i = 0; i < r.segs.Count(); i++) { q.seg_s.push_back(r.segs[i].start); q.seg_e.push_back(r.segs[i].end); }
    for (int i = 0; i < r.juncs.Count(); i++) {
      int* p = row.Find(r.juncs[i]);
      q.rjunc.push_back(p ? *p : -1);
    }
    for (int i = 0; i < r.pair_idx.Count(); i++) { q.pair_idx.push_back(r.pair_idx[i]); q.pair_cnt.push_back(r.pair_count[i]); }
    q.seg_off[n + 1] = (uint32_t)q.seg_s.size(); q.pair_off[n + 1] = (uint32_t)q.pair_idx.size();
  }
  if (q.rjunc.empty()) q.rjunc.push_back(0);
  memset(&v, 0, sizeof(v));
  v.start = bundle->start; v.end = bundle->end; v.n_reads = nr; v.n_juncs = nj;
  v.r_start = q.r_start.data(); v.r_end = q.r_end.data(); v.r_len = q.r_len.data(); v.r_count = q.r_count.data();
  v.r_nh = q.r_nh.data(); v.r_strand = q.r_strand.data(); v.r_flags = q.r_flags.data(); v.r_seg_off = q.seg_off.data();
  v.seg_start = q.seg_s.data(); v.seg_end = q.seg_e.data(); v.rjunc_idx = q.rjunc.data(); v.r_pair_off = q.pair_off.data();
  v.pair_idx = q.pair_idx.data(); v.pair_cnt = q.pair_cnt.data(); v.j_start = q.j_s.data(); v.j_end = q.j_e.data();
  v.j_strand = q.j_strand.data(); v.j_guide_match = q.j_gm.data(); v.j_nreads = q.j_nreads.data(); v.j_nm = q.j_nm.data();
  v.j_mm = q.j_mm.data();
  q.len = bundle->end - bundle->start + 3;             // rlink.cpp:16875
  q.cov.resize(3 * (size_t)q.len);
}

std::mutex g_mu;
std::condition_variable g_cv;
std::vector<std::pair<Request*, stg_bundle_view> > g_queue;   // flattened bundles waiting for a batch
bool g_leader = false;                                         // a batch is being collected or is on the GPU
int g_inside = 0;                                              // callers inside infer_transcripts (flattening or queued)

// the leader: everything queued goes to the GPU as one batch; BundleData::pred of every bundle is rebuilt here, while the
// context's result buffers are still this batch's
void run_batch(std::vector<std::pair<Request*, stg_bundle_view> >& batch) {
  const size_t n = batch.size();
  std::vector<stg_bundle_view> views(n);
  std::vector<stg_bundle_result> res(n);
  for (size_t i = 0; i < n; i++) {
    views[i] = batch[i].second;
    memset(&res[i], 0, sizeof(res[i]));
    res[i].bpcov = batch[i].first->cov.data();
    res[i].assemble = 1;
  }
  stg_ctx* ctx = context();
  const double t0 = now_s();
  if (stg_infer_transcripts_batch(ctx, (int64_t)n, views.data(), res.data()) != STG_OK) die(stg_last_error(ctx));
  g_gpu_s += now_s() - t0;
  g_batches++; g_calls += (long)n;
  if ((long)n > g_maxbatch) g_maxbatch = (long)n;
  for (size_t i = 0; i < n; i++) {
    Request& q = *batch[i].first;
    const stg_bundle_result& r = res[i];
    g_preds += r.n_pred;
    for (int k = 0; k < r.n_pred; k++) {                 // CPrediction, bundle.h:270-311
      CPrediction* p = new CPrediction(r.p_geneno[k], NULL, r.p_start[k], r.p_end[k], r.p_cov[k], (char)r.p_strand[k], r.p_tlen[k]);
      for (uint32_t e = r.p_exon_off[k]; e < r.p_exon_off[k + 1]; e++) {
        GSeg exon(r.e_start[e], r.e_end[e]);
        p->exons.Add(exon);
        float ec = r.e_cov[e];
        p->exoncov.Add(ec);
      }
      q.bundle->pred.Add(p);
      if (r.p_geneno[k] + 1 > q.geneno) q.geneno = r.p_geneno[k] + 1;
    }
    q.bundle->num_fragments += r.num_fragments;         // rlink.cpp:15105-15116 (both start at 0 for a fresh bundle)
    q.bundle->frag_len += r.frag_len;
  }
}
}  // namespace

int infer_transcripts(BundleData* bundle) {
  if (mergeMode) die("--merge is not implemented by libstgpu");
  if (bundle->keepguides.Count() == 0 && eonly) return 0;   // rlink.cpp:17177: nothing to do for this bundle
  { std::lock_guard<std::mutex> lk(g_mu); g_inside++; }
  Request q;
  stg_bundle_view v;
  const double t0 = now_s();
  flatten(bundle, q, v);
  const double t_flat = now_s() - t0;
  {
    std::unique_lock<std::mutex> lk(g_mu);
    g_flatten_s += t_flat;
    g_queue.push_back(std::make_pair(&q, v));
    g_cv.notify_all();
    while (!q.done) {
      if (g_leader) { g_cv.wait(lk); continue; }
      g_leader = true;
      // let the callers that are still flattening join this batch (bounded: a caller with a huge bundle is not waited for)
      for (int spins = 0; (int)g_queue.size() < g_inside && spins < 40; spins++) g_cv.wait_for(lk, std::chrono::microseconds(50));
      std::vector<std::pair<Request*, stg_bundle_view> > batch;
      batch.swap(g_queue);
      lk.unlock();
      run_batch(batch);
      lk.lock();
      for (size_t i = 0; i < batch.size(); i++) batch[i].first->done = true;
      g_leader = false;
      g_cv.notify_all();
    }
    g_inside--;
  }
  for (int s = 0; s < 3; s++) {                        // BundleData::bpcov, bundle.h:329 (printResults reads all three lanes)
    bundle->bpcov[s].Resize(q.len);
    memcpy(&bundle->bpcov[s][0], &q.cov[(size_t)s * q.len], (size_t)q.len * sizeof(float));
  }
  return q.geneno;
}
