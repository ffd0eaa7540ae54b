// stgpu_api.cu — host side of libstgpu: the C ABI of include/stgpu.h, the bundle batcher, HBM layout planning,
// kernel sequencing and result fetch.  No CPU compute path exists here: without a CUDA device stg_init fails.
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include "stgpu_internal.h"
#include "asm_device.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

thread_local std::string tl_error;

struct Timed { const char* name; cudaEvent_t e0, e1; };

}  // namespace

// Device memory is recycled through the context: a host that streams batch after batch (the reference's bundle
// queue does exactly that) must not pay cudaMalloc/cudaFree — tens of milliseconds for the ~25 GB of a C2 batch —
// on every one.  Blocks are handed back on stg_free_batch and reused best-fit; stg_shutdown / stg_trim release them.
struct PoolBlock { void* p; size_t bytes; };

struct OneBundleOut {   // BundleData::pred of the last stg_infer_transcripts call (owned by the context)
  std::vector<int32_t> p_start, p_end, p_tlen, p_geneno; std::vector<double> p_cov; std::vector<int8_t> p_strand;
  std::vector<uint32_t> p_exon_off, e_start, e_end; std::vector<float> e_cov;
  std::vector<double> j_good, j_left, j_right;
};

struct StagePending { void* dst; size_t off, bytes; };

struct stg_ctx {
  int device = 0;
  int num_sms = 0;
  stg_params params;
  cudaStream_t stream = nullptr;
  cudaStream_t aux_stream = nullptr;   // second stream of the rows a8 + a9 stage (deep bundles beside the shallow mass)
  cudaStream_t aux3_stream = nullptr;  // third stream: the oversized bundles (one cluster of CTAs each)
  // staging for the small host<->device transfers of the stage functions (sizes, counters, per-bundle tables): they go
  // through mapped pinned memory with a copy kernel on the compute stream instead of the copy engines, whose queues may
  // hold seconds of bulk traffic of the neighbouring batches (stg_upload, stg_fetch_bpcov_all_begin)
  uint8_t* stage_h = nullptr;
  size_t stage_cap = 0, stage_used = 0;
  std::vector<StagePending> stage_pending;
  int64_t cluster_reads = 0;           // bundles with at least this many reads are walked by a cluster of CTAs (row x1)
  cudaStream_t in_stream = nullptr;    // H2D copies of stg_upload (a batch can be uploaded while the previous one computes)
  size_t pool_limit = (size_t)160 << 30;   // set from the device's memory size in stg_init
  int64_t n_malloc = 0, malloc_bytes = 0, n_free = 0, n_pool_flush = 0;   // cudaMalloc / cudaFree calls behind the pool (stg_pool_stats)
  cudaStream_t out_stream = nullptr;   // D2H of bpcov started by stg_fetch_bpcov_all_begin (overlaps rows a8-a15)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr;
  std::string error;
  bool profiling = false;
  std::vector<Timed> timed;      // events of the last stg_run
  std::vector<cudaEvent_t> event_pool;
  size_t event_used = 0;
  int64_t launches = 0;
  std::mutex mu;
  std::vector<PoolBlock> pool_free;
  size_t pool_bytes = 0;          // bytes currently parked in pool_free
  struct stg_builder* one_bundle = nullptr;   // batcher reused by stg_infer_transcripts (keeps its pinned buffers)
  std::mutex one_bundle_mu;
  OneBundleOut one_out;
};

// Growable host array in PINNED memory (cudaMallocHost), so that stg_upload's copies out of the batcher run
// asynchronously at full PCIe rate; it keeps its capacity across stg_builder_reset.  Without a CUDA driver (CPU-only
// test boxes) it falls back to malloc.
template <class T>
struct PinnedVec {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  bool pinned = false;
  PinnedVec() {}
  PinnedVec(const PinnedVec&) = delete;
  PinnedVec& operator=(const PinnedVec&) = delete;
  ~PinnedVec() { drop(); }
  void drop() {
    if (p) { if (pinned) cudaFreeHost(p); else free(p); }
    p = nullptr; n = cap = 0;
  }
  bool reserve(size_t want) {
    if (want <= cap) return true;
    size_t nc = std::max(want, std::max(cap * 2, (size_t)(1 << 16) / sizeof(T)));
    T* q = nullptr;
    bool pin = cudaMallocHost((void**)&q, nc * sizeof(T)) == cudaSuccess;
    if (!pin) { cudaGetLastError(); q = (T*)malloc(nc * sizeof(T)); if (!q) return false; }
    if (n) memcpy(q, p, n * sizeof(T));
    if (p) { if (pinned) cudaFreeHost(p); else free(p); }
    p = q; cap = nc; pinned = pin;
    return true;
  }
  bool push_back(T v) { if (n == cap && !reserve(n + 1)) return false; p[n++] = v; return true; }
  bool append(const T* src, size_t k) {
    if (!k) return true;
    if (n + k > cap && !reserve(n + k)) return false;
    memcpy(p + n, src, k * sizeof(T)); n += k;
    return true;
  }
  T* data() { return p; }
  size_t size() const { return n; }
  void clear() { n = 0; }
};

struct stg_builder {
  stg_ctx* ctx;
  PinnedVec<int32_t> b_start, b_end;
  PinnedVec<uint32_t> b_read_off, b_junc_off;
  PinnedVec<uint32_t> r_start, r_end, r_len;
  PinnedVec<float> r_count;
  PinnedVec<int16_t> r_nh;
  PinnedVec<int8_t> r_strand;
  PinnedVec<uint8_t> r_flags;
  PinnedVec<uint32_t> r_seg_off, r_pair_off, seg_start, seg_end;
  PinnedVec<int32_t> rjunc_idx, pair_idx;
  PinnedVec<float> pair_cnt;
  PinnedVec<uint32_t> j_start, j_end;
  PinnedVec<int8_t> j_strand, j_guide_match;
  PinnedVec<double> j_nreads, j_nm, j_mm;
  void reset() {
    b_start.clear(); b_end.clear(); b_read_off.clear(); b_junc_off.clear(); r_start.clear(); r_end.clear();
    r_len.clear(); r_count.clear(); r_nh.clear(); r_strand.clear(); r_flags.clear(); r_seg_off.clear();
    r_pair_off.clear(); seg_start.clear(); seg_end.clear(); rjunc_idx.clear(); pair_idx.clear(); pair_cnt.clear();
    j_start.clear(); j_end.clear(); j_strand.clear(); j_guide_match.clear(); j_nreads.clear(); j_nm.clear(); j_mm.clear();
    b_read_off.push_back(0); b_junc_off.push_back(0); r_seg_off.push_back(0); r_pair_off.push_back(0);
  }
};

struct stg_dbatch {
  DevBatchView v;
  std::vector<PoolBlock> allocs;   // every device block this batch holds (returned to the context pool on free)
  std::vector<int64_t> h_cov_off;  // [n_bundles] float offset of lane 0 of bundle b (= gbase)
  std::vector<int64_t> h_stride;   // [n_bundles] distance between the strand lanes of bundle b (= total_pos: the lanes are batch-wide)
  std::vector<uint32_t> h_len;
  std::vector<uint32_t> h_tile_off;   // [n_bundles + 1] first sub-tile of every bundle
  uint32_t* pk_desc = nullptr; float* pk_val = nullptr; float* pk_raw = nullptr;   // packed bpcov (stg_pack_bpcov)
  int64_t pk_nraw = -1, pk_raw_cap = 0;
  size_t pk_mark = 0;                  // allocs.size() when the packed form was first allocated
  bool released = false;               // stg_release_device: no device memory behind the handle any more
  std::vector<int32_t> h_start;    // [n_bundles] BundleData::start
  int64_t ivl_capacity = 0;     // intervals the interval / event / chunk arrays can hold
  int64_t cell_capacity = 0;    // counter cells of the binning matrix
  int64_t scan_capacity = 0;    // elements the scan scratch covers
  int64_t sum_len = 0;
  bool ran = false;
  // rows a8 + a9 (stg_build_groups)
  GroupsView gv;
  bool groups_built = false;
  NeutralView nv;                  // row a19 (stg_neutral_predictions)
  bool neutral_done = false;
  std::vector<uint32_t> h_nreads;  // [n_bundles]
  AsmView av;                      // rows a10-a15 (stg_assemble)
  bool assembled = false;
  size_t stage_mark = 0;           // allocs.size() when stg_build_groups started: what stg_rewind gives back
  cudaEvent_t ev_uploaded = nullptr;   // on in_stream after the last H2D copy of stg_upload
  cudaEvent_t ev_cov = nullptr;        // on the context stream after the last kernel that writes bpcov
  bool out_pending = false;            // a copy of stg_fetch_bpcov_all_begin may be in flight on out_stream
};

namespace {

int fail(stg_ctx* ctx, int code, const std::string& msg) {
  tl_error = msg;
  if (ctx) ctx->error = msg;
  return code;
}

#define CK(ctx, call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(ctx, STG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));            \
  } while (0)

void* pool_take(stg_ctx* ctx, size_t bytes, size_t* got) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  int best = -1;
  for (int i = 0; i < (int)ctx->pool_free.size(); i++) {
    const size_t b = ctx->pool_free[i].bytes;
    if (b >= bytes && b <= bytes + bytes / 4 + (1u << 20) && (best < 0 || b < ctx->pool_free[best].bytes)) best = i;
  }
  if (best < 0) return nullptr;
  void* p = ctx->pool_free[best].p;
  *got = ctx->pool_free[best].bytes;
  ctx->pool_bytes -= *got;
  ctx->pool_free.erase(ctx->pool_free.begin() + best);
  return p;
}

void pool_release_all(stg_ctx* ctx) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (auto& b : ctx->pool_free) cudaFree(b.p);
  ctx->pool_free.clear();
  ctx->pool_bytes = 0;
}

template <class T>
int dev_alloc(stg_ctx* ctx, stg_dbatch* d, T** out, int64_t count) {
  size_t bytes = ((size_t)std::max<int64_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
  size_t got = 0;
  void* p = pool_take(ctx, bytes, &got);
  if (!p) {
    cudaError_t e = cudaMalloc(&p, bytes);
    ctx->n_malloc++; ctx->malloc_bytes += (int64_t)bytes;
    if (e != cudaSuccess) {   // make room: drop everything parked in the pool and retry once
      cudaGetLastError();
      pool_release_all(ctx);
      ctx->n_pool_flush++;
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return fail(ctx, STG_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    got = bytes;
  }
  d->allocs.push_back(PoolBlock{p, got});
  *out = (T*)p;
  return STG_OK;
}

// Bulk copies go out in pieces: a copy engine serves its queue one command at a time, so a multi-GB command would hold up
// the 4-byte read-backs of the batch that is computing meanwhile (they travel in the same direction on another stream).
cudaError_t copy_chunked(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  const size_t CH = (size_t)32 << 20;
  for (size_t o = 0; o < bytes; o += CH) {
    const cudaError_t e = cudaMemcpyAsync((char*)dst + o, (const char*)src + o, bytes - o < CH ? bytes - o : CH, kind, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

__global__ void __launch_bounds__(256) copy_bytes_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes) {
  const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = gridDim.x * (size_t)blockDim.x;
  if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) == 0) {
    for (size_t i = i0; i < bytes / 16; i += stride) ((uint4*)dst)[i] = ((const uint4*)src)[i];
  } else if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 3u) == 0) {
    for (size_t i = i0; i < bytes / 4; i += stride) ((uint32_t*)dst)[i] = ((const uint32_t*)src)[i];
  } else {
    for (size_t i = i0; i < bytes; i += stride) dst[i] = src[i];
  }
}

cudaError_t launch_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return cudaSuccess;
  const size_t want = (bytes / 16 + 255) / 256;
  const unsigned grid = (unsigned)std::min<size_t>(std::max<size_t>(want, 1), 148 * 8);
  copy_bytes_kernel<<<grid, 256, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, bytes);
  return cudaGetLastError();
}

// device -> host, delivered by small_sync()
cudaError_t small_d2h(stg_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return cudaSuccess;
  const size_t off = (ctx->stage_used + 15) & ~(size_t)15;
  if (!ctx->stage_h || off + bytes > ctx->stage_cap) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  ctx->stage_used = off + bytes;
  ctx->stage_pending.push_back(StagePending{dst, off, bytes});
  return launch_copy(ctx->stage_h + off, src, bytes, st);
}

// host -> device; the host array may be reused as soon as this returns
cudaError_t small_h2d(stg_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return cudaSuccess;
  const size_t off = (ctx->stage_used + 15) & ~(size_t)15;
  if (!ctx->stage_h || off + bytes > ctx->stage_cap) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    return e != cudaSuccess ? e : cudaStreamSynchronize(st);
  }
  ctx->stage_used = off + bytes;
  memcpy(ctx->stage_h + off, src, bytes);
  return launch_copy(dst, ctx->stage_h + off, bytes, st);
}

struct StageScope {
  stg_ctx* c;
  explicit StageScope(stg_ctx* ctx) : c(ctx) {}
  ~StageScope() { c->stage_pending.clear(); c->stage_used = 0; }
};

// wait for the stream, hand the staged read-backs to their destinations, recycle the staging area
cudaError_t small_sync(stg_ctx* ctx, cudaStream_t st) {
  const cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess)
    for (const StagePending& p : ctx->stage_pending) memcpy(p.dst, ctx->stage_h + p.off, p.bytes);
  ctx->stage_pending.clear();
  ctx->stage_used = 0;
  return e;
}

template <class T>
int dev_upload(stg_ctx* ctx, stg_dbatch* d, const T** out, const T* host, int64_t count) {
  T* p = nullptr;
  int rc = dev_alloc(ctx, d, &p, count);
  if (rc) return rc;
  if (count > 0) CK(ctx, copy_chunked(p, host, (size_t)count * sizeof(T), cudaMemcpyHostToDevice, ctx->in_stream));
  *out = p;
  return STG_OK;
}

void free_batch_allocs(stg_ctx* ctx, stg_dbatch* d) {
  if (ctx) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& b : d->allocs) {
      // no bound of its own: a C2 batch takes ~52 GB and a streaming host has two in flight plus one parked; when the device
      // runs out, dev_alloc drops the whole pool once and retries (cudaFree / cudaMalloc synchronise the device, so the pool
      // must not churn in the steady state: stg_pool_stats)
      if (ctx->pool_bytes + b.bytes > ctx->pool_limit) { cudaFree(b.p); ctx->n_free++; continue; }
      ctx->pool_free.push_back(b); ctx->pool_bytes += b.bytes;
    }
  } else {
    for (auto& b : d->allocs) cudaFree(b.p);
  }
  d->allocs.clear();
}

cudaEvent_t take_event(stg_ctx* ctx) {
  if (ctx->event_used == ctx->event_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->event_pool.push_back(e);
  }
  return ctx->event_pool[ctx->event_used++];
}

struct Scope {  // brackets a group of launches with events when profiling is on
  stg_ctx* ctx; Timed t; bool on;
  Scope(stg_ctx* c, const char* name, int nlaunch) : ctx(c), on(c->profiling) {
    ctx->launches += nlaunch;
    t.name = name;
    if (on) { t.e0 = take_event(ctx); t.e1 = take_event(ctx); cudaEventRecord(t.e0, ctx->stream); }
  }
  ~Scope() { if (on) { cudaEventRecord(t.e1, ctx->stream); ctx->timed.push_back(t); } }
};

int validate(stg_ctx* ctx, const stg_packed_batch* h) {
  if (!h) return fail(ctx, STG_EINVAL, "null batch");
  if (h->n_bundles < 0 || h->n_reads < 0 || h->n_segs < h->n_reads || h->n_pairs < 0 || h->n_juncs < 0)
    return fail(ctx, STG_EINVAL, "negative or inconsistent counts");
  if (h->n_reads >= (int64_t)0x7fffffff || h->n_segs >= (int64_t)0xffffffffLL || h->n_pairs >= (int64_t)0xffffffffLL ||
      h->n_juncs >= (int64_t)0x7fffffff)
    return fail(ctx, STG_ERANGE, "batch exceeds 32-bit index space; split it");
  if (h->n_bundles == 0) {
    if (h->n_reads || h->n_segs || h->n_pairs || h->n_juncs) return fail(ctx, STG_EINVAL, "reads / junctions without a bundle");
    return STG_OK;
  }
  if (!h->b_start || !h->b_end || !h->b_read_off || !h->b_junc_off) return fail(ctx, STG_EINVAL, "null bundle array");
  if (h->n_reads && (!h->r_start || !h->r_end || !h->r_len || !h->r_count || !h->r_nh || !h->r_strand || !h->r_flags || !h->r_seg_off ||
                     !h->r_pair_off || !h->seg_start || !h->seg_end))
    return fail(ctx, STG_EINVAL, "null read array");
  if (h->n_segs > h->n_reads && !h->rjunc_idx) return fail(ctx, STG_EINVAL, "null rjunc_idx");
  if (h->n_pairs && (!h->pair_idx || !h->pair_cnt)) return fail(ctx, STG_EINVAL, "null pair array");
  if (h->n_juncs && (!h->j_start || !h->j_end || !h->j_strand || !h->j_guide_match || !h->j_nreads || !h->j_nm || !h->j_mm))
    return fail(ctx, STG_EINVAL, "null junction array");
  if (h->b_read_off[0] != 0 || h->b_read_off[h->n_bundles] != (uint32_t)h->n_reads)
    return fail(ctx, STG_EINVAL, "b_read_off does not span the reads");
  if (h->b_junc_off[0] != 0 || h->b_junc_off[h->n_bundles] != (uint32_t)h->n_juncs)
    return fail(ctx, STG_EINVAL, "b_junc_off does not span the junctions");
  if (h->n_reads && (h->r_seg_off[0] != 0 || h->r_seg_off[h->n_reads] != (uint32_t)h->n_segs))
    return fail(ctx, STG_EINVAL, "r_seg_off does not span the segments");
  if (h->n_reads && (h->r_pair_off[0] != 0 || h->r_pair_off[h->n_reads] != (uint32_t)h->n_pairs))
    return fail(ctx, STG_EINVAL, "r_pair_off does not span the pair links");
  // every read has at least one segment: the introns of read n sit at r_seg_off[n] - n, which must not wrap
  for (int64_t n = 0; n < h->n_reads; n++)
    if (h->r_seg_off[n + 1] <= h->r_seg_off[n]) return fail(ctx, STG_EINVAL, "read without segments (r_seg_off not strictly increasing)");
  for (int64_t b = 0; b < h->n_bundles; b++) {
    if (h->b_end[b] < h->b_start[b]) return fail(ctx, STG_EINVAL, "bundle with end < start");
    if (h->b_read_off[b + 1] < h->b_read_off[b] || h->b_junc_off[b + 1] < h->b_junc_off[b])
      return fail(ctx, STG_EINVAL, "bundle offsets not monotone");
  }
  return STG_OK;
}

}  // namespace

extern "C" {

void stg_params_default(stg_params* p) {
  memset(p, 0, sizeof(*p));
  p->struct_size = sizeof(stg_params);
  p->junctionsupport = 10; p->sserror = 25; p->junctionthr = 1; p->bundledist = 50;
  p->mintranscriptlen = 200; p->allowed_nodes = 1000;
  p->readthr = 1.f; p->singlethr = 4.75f; p->isofrac = 0.01f; p->mcov = 1.f;
  p->trim = 1; p->isunitig = 1; p->includesource = 1;
}

int stg_abi_version(void) { return STG_ABI_VERSION; }

const char* stg_last_error(const stg_ctx* ctx) { return ctx ? ctx->error.c_str() : tl_error.c_str(); }

int stg_init(const stg_params* params, int device, stg_ctx** out) {
  if (!out) return fail(nullptr, STG_EINVAL, "stg_init: out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, STG_ENODEV, std::string("no CUDA device (libstgpu has no CPU path): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, STG_EINVAL, "stg_init: bad device index");
  stg_ctx* ctx = new stg_ctx();
  ctx->device = device;
  if (params) {
    if (params->struct_size != sizeof(stg_params)) { delete ctx; return fail(nullptr, STG_EINVAL, "stg_params.struct_size mismatch"); }
    ctx->params = *params;
  } else stg_params_default(&ctx->params);
  if (ctx->params.longreads || ctx->params.mixedMode) {
    delete ctx;
    return fail(nullptr, STG_EUNSUPPORTED, "long-read / mixed mode junction correction (rlink.cpp:16668-16975) is not built yet");
  }
  if (ctx->params.have_gseq) {
    delete ctx;
    return fail(nullptr, STG_EUNSUPPORTED, "splice-consensus filtering from the genomic sequence (--ref / CRAM reference, rlink.cpp:14456-14530) is not built: "
                                           "the junction pass would keep junctions the reference drops");
  }
  if (ctx->params.viral || ctx->params.havePtFeatures) {
    delete ctx;
    return fail(nullptr, STG_EUNSUPPORTED, "--viral / --ptf branches of the junction pass (rlink.cpp:14189-14202, 14543-14589) are not built yet");
  }
  auto init_fail = [&](cudaError_t e, const char* what) {
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux3_stream) cudaStreamDestroy(ctx->aux3_stream);
    if (ctx->in_stream) cudaStreamDestroy(ctx->in_stream);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    if (ctx->stage_h) cudaFreeHost(ctx->stage_h);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_join3) cudaEventDestroy(ctx->ev_join3);
    delete ctx;
    return fail(nullptr, STG_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
  };
#define ICK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return init_fail(e__, #call); } while (0)
  ICK(cudaSetDevice(device));
  cudaDeviceProp prop;
  ICK(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    delete ctx;
    return fail(nullptr, STG_ENODEV, "libstgpu is built for sm_100a (B200) only");
  }
  // tuning / test aid: the tests lower the threshold to drive ordinary bundles through the cluster kernel
  ctx->cluster_reads = getenv("STGPU_CLUSTER_READS") ? atoll(getenv("STGPU_CLUSTER_READS")) : groups_cluster_reads();
  if (ctx->cluster_reads < 1) ctx->cluster_reads = 1;
  ICK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  { size_t fr = 0, tot = 0; if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && tot) ctx->pool_limit = tot - tot / 16; }
  ICK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
  ICK(cudaStreamCreateWithFlags(&ctx->aux3_stream, cudaStreamNonBlocking));
  ctx->stage_cap = (size_t)16 << 20;
  ICK(cudaHostAlloc((void**)&ctx->stage_h, ctx->stage_cap, cudaHostAllocMapped | cudaHostAllocPortable));
  ICK(cudaStreamCreateWithFlags(&ctx->in_stream, cudaStreamNonBlocking));
  ICK(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
  ICK(cudaEventCreateWithFlags(&ctx->ev_join3, cudaEventDisableTiming));
  ICK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  ICK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
#undef ICK
  *out = ctx;
  return STG_OK;
}

int stg_shutdown(stg_ctx* ctx) {
  if (!ctx) return STG_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  pool_release_all(ctx);
  if (ctx->one_bundle) stg_builder_destroy(ctx->one_bundle);
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->aux3_stream) { cudaStreamSynchronize(ctx->aux3_stream); cudaStreamDestroy(ctx->aux3_stream); }
  if (ctx->in_stream) { cudaStreamSynchronize(ctx->in_stream); cudaStreamDestroy(ctx->in_stream); }
  if (ctx->out_stream) { cudaStreamSynchronize(ctx->out_stream); cudaStreamDestroy(ctx->out_stream); }
  if (ctx->stage_h) cudaFreeHost(ctx->stage_h);
  if (ctx->ev_join3) cudaEventDestroy(ctx->ev_join3);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  delete ctx;
  return STG_OK;
}

void* stg_stream(stg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int stg_trim(stg_ctx* ctx) {
  if (!ctx) return fail(nullptr, STG_EINVAL, "stg_trim: null ctx");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  pool_release_all(ctx);
  return STG_OK;
}

// ---- bundle batcher -----------------------------------------------------------------------------------------
int stg_builder_create(stg_ctx* ctx, stg_builder** out) {
  if (!out) return fail(ctx, STG_EINVAL, "stg_builder_create: out is null");
  if (ctx) cudaSetDevice(ctx->device);
  stg_builder* b = new stg_builder();
  b->ctx = ctx;
  b->reset();
  *out = b;
  return STG_OK;
}

int stg_builder_add(stg_builder* b, const stg_bundle_view* v) {
  if (!b || !v) return fail(nullptr, STG_EINVAL, "stg_builder_add: null argument");
  if (v->n_reads < 0 || v->n_juncs < 0 || v->end < v->start) return fail(b->ctx, STG_EINVAL, "stg_builder_add: bad view");
  const int32_t nr = v->n_reads;
  if (nr && (v->r_seg_off[0] != 0 || v->r_pair_off[0] != 0))
    return fail(b->ctx, STG_EINVAL, "stg_builder_add: the CSR offsets of a view start at 0 (rjunc_idx is indexed from them)");
  const uint32_t s0 = 0, p0 = 0;
  const uint32_t nseg = nr ? v->r_seg_off[nr] - s0 : 0;
  const uint32_t npair = nr ? v->r_pair_off[nr] - p0 : 0;
  for (int32_t n = 0; n < nr; n++)
    if (v->r_seg_off[n + 1] <= v->r_seg_off[n]) return fail(b->ctx, STG_EINVAL, "stg_builder_add: read without segments");
  const uint32_t seg_base = (uint32_t)b->seg_start.size(), pair_base = (uint32_t)b->pair_idx.size();
  // sizes at entry: a failed append leaves the builder as it was
  const size_t z_b = b->b_start.size(), z_r = b->r_start.size(), z_so = b->r_seg_off.size(), z_po = b->r_pair_off.size(),
               z_i = b->rjunc_idx.size(), z_j = b->j_start.size(), z_bro = b->b_read_off.size(), z_bjo = b->b_junc_off.size();
  bool ok = b->b_start.push_back(v->start) && b->b_end.push_back(v->end);
  // the per-read arrays of a view are contiguous: bulk copies, only the CSR offsets are rebased one by one
  ok = ok && b->r_start.append(v->r_start, nr) && b->r_end.append(v->r_end, nr) && b->r_len.append(v->r_len, nr) &&
       b->r_count.append(v->r_count, nr) && b->r_nh.append(v->r_nh, nr) && b->r_strand.append(v->r_strand, nr) &&
       b->r_flags.append(v->r_flags, nr);
  ok = ok && b->r_seg_off.reserve(b->r_seg_off.size() + nr) && b->r_pair_off.reserve(b->r_pair_off.size() + nr);
  if (ok) {
    uint32_t* so = b->r_seg_off.data() + b->r_seg_off.size();
    uint32_t* po = b->r_pair_off.data() + b->r_pair_off.size();
    for (int32_t n = 0; n < nr; n++) {
      so[n] = seg_base + (v->r_seg_off[n + 1] - s0);
      po[n] = pair_base + (v->r_pair_off[n + 1] - p0);
    }
    b->r_seg_off.n += nr; b->r_pair_off.n += nr;
  }
  ok = ok && b->seg_start.append(v->seg_start + s0, nseg) && b->seg_end.append(v->seg_end + s0, nseg) &&
       b->rjunc_idx.append(v->rjunc_idx, nseg - (uint32_t)nr) && b->pair_idx.append(v->pair_idx + p0, npair) &&
       b->pair_cnt.append(v->pair_cnt + p0, npair) && b->j_start.append(v->j_start, v->n_juncs) &&
       b->j_end.append(v->j_end, v->n_juncs) && b->j_strand.append(v->j_strand, v->n_juncs) &&
       b->j_guide_match.append(v->j_guide_match, v->n_juncs) && b->j_nreads.append(v->j_nreads, v->n_juncs) &&
       b->j_nm.append(v->j_nm, v->n_juncs) && b->j_mm.append(v->j_mm, v->n_juncs);
  ok = ok && b->b_read_off.push_back((uint32_t)b->r_start.size()) && b->b_junc_off.push_back((uint32_t)b->j_start.size());
  if (!ok) {
    b->b_start.n = z_b; b->b_end.n = z_b; b->r_start.n = z_r; b->r_end.n = z_r; b->r_len.n = z_r; b->r_count.n = z_r; b->r_nh.n = z_r;
    b->r_strand.n = z_r; b->r_flags.n = z_r; b->r_seg_off.n = z_so; b->r_pair_off.n = z_po; b->seg_start.n = seg_base; b->seg_end.n = seg_base;
    b->rjunc_idx.n = z_i; b->pair_idx.n = pair_base; b->pair_cnt.n = pair_base; b->j_start.n = z_j; b->j_end.n = z_j; b->j_strand.n = z_j;
    b->j_guide_match.n = z_j; b->j_nreads.n = z_j; b->j_nm.n = z_j; b->j_mm.n = z_j; b->b_read_off.n = z_bro; b->b_junc_off.n = z_bjo;
    return fail(b->ctx, STG_ENOMEM, "stg_builder_add: out of host memory");
  }
  return STG_OK;
}

int stg_builder_packed(stg_builder* b, stg_packed_batch* o) {
  if (!b || !o) return fail(nullptr, STG_EINVAL, "stg_builder_packed: null argument");
  o->n_bundles = (int64_t)b->b_start.size(); o->n_reads = (int64_t)b->r_start.size();
  o->n_segs = (int64_t)b->seg_start.size(); o->n_pairs = (int64_t)b->pair_idx.size(); o->n_juncs = (int64_t)b->j_start.size();
  o->b_start = b->b_start.data(); o->b_end = b->b_end.data(); o->b_read_off = b->b_read_off.data();
  o->b_junc_off = b->b_junc_off.data(); o->r_start = b->r_start.data(); o->r_end = b->r_end.data();
  o->r_len = b->r_len.data(); o->r_count = b->r_count.data(); o->r_nh = b->r_nh.data(); o->r_strand = b->r_strand.data();
  o->r_flags = b->r_flags.data(); o->r_seg_off = b->r_seg_off.data(); o->r_pair_off = b->r_pair_off.data();
  o->seg_start = b->seg_start.data(); o->seg_end = b->seg_end.data(); o->rjunc_idx = b->rjunc_idx.data();
  o->pair_idx = b->pair_idx.data(); o->pair_cnt = b->pair_cnt.data(); o->j_start = b->j_start.data();
  o->j_end = b->j_end.data(); o->j_strand = b->j_strand.data(); o->j_guide_match = b->j_guide_match.data();
  o->j_nreads = b->j_nreads.data(); o->j_nm = b->j_nm.data(); o->j_mm = b->j_mm.data();
  return STG_OK;
}

int stg_builder_reset(stg_builder* b) {   // keeps the pinned buffers for the next batch
  if (!b) return STG_OK;
  b->reset();
  return STG_OK;
}

int stg_builder_destroy(stg_builder* b) { delete b; return STG_OK; }

// ---- upload: validate, plan the layout, copy ----------------------------------------------------------------
int stg_upload(stg_ctx* ctx, const stg_packed_batch* h, stg_dbatch** out) {
  if (!ctx || !out) return fail(ctx, STG_EINVAL, "stg_upload: null argument");
  *out = nullptr;
  int rc = validate(ctx, h);
  if (rc) return rc;
  CK(ctx, cudaSetDevice(ctx->device));
  stg_dbatch* d = new stg_dbatch();
  DevBatchView& v = d->v;
  memset(&v, 0, sizeof(v));
  v.n_bundles = h->n_bundles; v.n_reads = h->n_reads; v.n_segs = h->n_segs; v.n_pairs = h->n_pairs; v.n_juncs = h->n_juncs;
  const int64_t nb = h->n_bundles;
  // layout
  std::vector<uint32_t> len(nb), stride(nb), gbase(nb + 1), tile_off(nb + 1);
  d->h_cov_off.resize(nb); d->h_stride.resize(nb); d->h_len.resize(nb); d->h_start.resize(nb);
  uint64_t g = 0, t = 0;
  for (int64_t b = 0; b < nb; b++) {
    const uint64_t l = (uint64_t)((int64_t)h->b_end[b] - (int64_t)h->b_start[b] + 3);
    const uint64_t s = (l + STG_LANE_ALIGN - 1) / STG_LANE_ALIGN * STG_LANE_ALIGN;
    if (l >= 0x7fffffffull) { delete d; return fail(ctx, STG_ERANGE, "bundle span too large"); }
    len[b] = (uint32_t)l; stride[b] = (uint32_t)s; gbase[b] = (uint32_t)g; tile_off[b] = (uint32_t)t;
    d->h_cov_off[b] = (int64_t)g; d->h_stride[b] = 0; d->h_len[b] = (uint32_t)l; d->h_start[b] = h->b_start[b];
    d->sum_len += (int64_t)l;
    d->h_nreads.push_back(h->b_read_off[b + 1] - h->b_read_off[b]);
    g += s; t += (l + STG_TILE_W - 1) / STG_TILE_W;
    if (g >= 0xffffffffull || t >= 0xffffffffull) { delete d; return fail(ctx, STG_ERANGE, "batch spans more than 2^32 positions; split it"); }
  }
  gbase[nb] = (uint32_t)g; tile_off[nb] = (uint32_t)t;
  d->h_tile_off = tile_off;
  v.total_pos = (int64_t)g; v.n_tiles = (int64_t)t;
  v.cov_pitch = (size_t)g;
  for (int64_t b = 0; b < nb; b++) d->h_stride[b] = (int64_t)g;
  // rounds of the tile schedule (tile_plan_kernel): at least the sub-tile count of the longest bundle
  {
    uint32_t max_nt = 1;
    for (int64_t b = 0; b < nb; b++) max_nt = std::max(max_nt, tile_off[b + 1] - tile_off[b]);
    v.n_rounds = std::max<uint32_t>(max_nt, 2048u);
  }

  // bundle of the first read of every block of STG_READ_BLOCK reads (the per-read kernels start their search there)
  std::vector<uint32_t> cta_bundle((size_t)((h->n_reads + STG_READ_BLOCK - 1) / STG_READ_BLOCK) + 1, 0);
  {
    int64_t b = 0;
    for (size_t i = 0; i + 1 < cta_bundle.size(); i++) {
      const uint32_t n = (uint32_t)(i * STG_READ_BLOCK);
      while (b + 1 < nb && h->b_read_off[b + 1] <= n) b++;
      cta_bundle[i] = (uint32_t)b;
    }
  }
  CK(ctx, cudaEventCreateWithFlags(&d->ev_uploaded, cudaEventDisableTiming));
  CK(ctx, cudaEventCreateWithFlags(&d->ev_cov, cudaEventDisableTiming));
#define UPFAIL do { cudaStreamSynchronize(ctx->in_stream); free_batch_allocs(ctx, d); cudaEventDestroy(d->ev_uploaded); cudaEventDestroy(d->ev_cov); delete d; return rc; } while (0)
#define UPV(field, vec) do { rc = dev_upload(ctx, d, &v.field, vec.data(), (int64_t)vec.size()); if (rc) UPFAIL; } while (0)
  UPV(b_len, len); UPV(b_stride, stride); UPV(b_gbase, gbase); UPV(b_tile_off, tile_off);
  UPV(cta_bundle, cta_bundle);
#undef UPV
  // the layout vectors are pageable locals: make sure their (small) copies have left the host before they die; the
  // caller's arrays follow on the same copy stream and are only waited for by stg_run / stg_upload_wait
  CK(ctx, cudaStreamSynchronize(ctx->in_stream));
#define UP(field, count) do { rc = dev_upload(ctx, d, &v.field, h->field, (int64_t)(count)); if (rc) UPFAIL; } while (0)
  UP(b_start, nb); UP(b_end, nb); UP(b_read_off, nb + 1); UP(b_junc_off, nb + 1);
  UP(r_start, h->n_reads); UP(r_end, h->n_reads); UP(r_len, h->n_reads); UP(r_count, h->n_reads); UP(r_nh, h->n_reads);
  UP(r_strand, h->n_reads); UP(r_flags, h->n_reads); UP(r_seg_off, h->n_reads + 1); UP(r_pair_off, h->n_reads + 1);
  UP(seg_start, h->n_segs); UP(seg_end, h->n_segs); UP(rjunc_idx, h->n_segs - h->n_reads);
  UP(pair_idx, h->n_pairs); UP(pair_cnt, h->n_pairs);
  UP(j_start, h->n_juncs); UP(j_end, h->n_juncs); UP(j_strand, h->n_juncs); UP(j_guide_match, h->n_juncs);
  UP(j_nreads, h->n_juncs); UP(j_nm, h->n_juncs); UP(j_mm, h->n_juncs);
#undef UP
  CK(ctx, cudaEventRecord(d->ev_uploaded, ctx->in_stream));
#define AL(field, count) do { rc = dev_alloc(ctx, d, &v.field, (int64_t)(count)); if (rc) UPFAIL; } while (0)
  AL(r_ivl_cnt, h->n_reads + 1 + 16); AL(scan_state, (h->n_reads + 1) / 4096 + 2);
  d->scan_capacity = std::max<int64_t>(std::max<int64_t>(h->n_reads + 1, v.n_tiles + 1), v.n_rounds + 1);
  AL(scan_partials, scan_partials_needed(d->scan_capacity));
  AL(tile_carry, v.n_tiles * 4); AL(tile_ticket, 6); AL(deep_count, 4); AL(err_flag, 1);
  AL(round_off, v.n_rounds + 1); AL(round_cur, v.n_rounds + 1);
  AL(tile_plan, v.n_tiles); AL(tile_evoff, v.n_tiles + 1); AL(tile_lastne, v.n_tiles); AL(junc_anchor, h->n_juncs);
  AL(bpcov, 3 * v.total_pos);
  AL(j_nreads_good, h->n_juncs); AL(j_leftsupport, h->n_juncs); AL(j_rightsupport, h->n_juncs);
  AL(b_cov_total, nb * 3);
  AL(jp_strand, h->n_juncs); AL(jp_nreads, h->n_juncs); AL(jp_nreads_good, h->n_juncs); AL(jp_left, h->n_juncs);
  AL(jp_right, h->n_juncs); AL(jp_mm, h->n_juncs); AL(jp_ej, h->n_juncs); AL(jp_good, h->n_juncs); AL(jp_good_strand, h->n_juncs);
#undef AL
#undef UPFAIL
  // the padding between a lane's length and its stride is written by the tile kernels; nothing else to clear
  *out = d;
  return STG_OK;
}

int stg_free_batch(stg_ctx* ctx, stg_dbatch* d) {
  if (!d) return STG_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    if (d->ev_uploaded) cudaEventSynchronize(d->ev_uploaded);
    cudaStreamSynchronize(ctx->stream);
    if (d->out_pending) cudaStreamSynchronize(ctx->out_stream);
  }
  if (d->ev_uploaded) cudaEventDestroy(d->ev_uploaded);
  if (d->ev_cov) cudaEventDestroy(d->ev_cov);
  free_batch_allocs(ctx, d);
  delete d;
  return STG_OK;
}

// The device memory of a batch goes back to the pool; the handle stays valid for stg_unpack_bpcov (host-side layout tables)
// until stg_free_batch.
int stg_release_device(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_release_device: null argument");
  cudaSetDevice(ctx->device);
  if (d->ev_uploaded) cudaEventSynchronize(d->ev_uploaded);
  cudaStreamSynchronize(ctx->stream);
  if (d->out_pending) { cudaStreamSynchronize(ctx->out_stream); d->out_pending = false; }
  free_batch_allocs(ctx, d);
  d->ran = false; d->groups_built = false; d->neutral_done = false; d->assembled = false; d->released = true;
  d->pk_desc = nullptr; d->pk_val = nullptr; d->pk_raw = nullptr; d->pk_raw_cap = 0; d->pk_nraw = -1;
  return STG_OK;
}

// ---- run: the kernel sequence of the stage -----------------------------------------------------------------
int stg_run(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_run: null argument");
  if (d->released) return fail(ctx, STG_EINVAL, "stg_run: the batch's device memory has been released");
  static const bool run_dbg = getenv("STGPU_RUN_DEBUG") != nullptr;   // host clock at the stage's sync points
  struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
  CK(ctx, cudaSetDevice(ctx->device));
  DevBatchView& v = d->v;
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);   // an early return leaves no staged read-back behind
  auto mark = [&](const char* what) { if (run_dbg) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); fprintf(stderr, "stg_run: %-28s %8.3f ms\n", what, 1e3 * (double)(t.tv_sec - ts0.tv_sec) + 1e-6 * (double)(t.tv_nsec - ts0.tv_nsec)); } };
  // the caller's host arrays are free again once this returns; a pending D2H of bpcov must not see the rewrite
  mark("device set");
  CK(ctx, cudaEventSynchronize(d->ev_uploaded));
  mark("upload waited for");
  if (d->out_pending) { CK(ctx, cudaStreamSynchronize(ctx->out_stream)); d->out_pending = false; }
  ctx->timed.clear();
  ctx->event_used = 0;
  KernelParams kp{ctx->params.junctionsupport, ctx->params.junctionthr, ctx->params.longreads};
  if (v.n_juncs) {
    CK(ctx, cudaMemsetAsync(v.j_nreads_good, 0, sizeof(double) * v.n_juncs, st));
    CK(ctx, cudaMemsetAsync(v.j_leftsupport, 0, sizeof(double) * v.n_juncs, st));
    CK(ctx, cudaMemsetAsync(v.j_rightsupport, 0, sizeof(double) * v.n_juncs, st));
  }
  if (v.n_tiles) CK(ctx, cudaMemsetAsync(v.tile_carry, 0xff, sizeof(unsigned long long) * 4 * (size_t)v.n_tiles, st));
  CK(ctx, cudaMemsetAsync(v.tile_ticket, 0, 3 * sizeof(uint32_t), st));
  CK(ctx, cudaMemsetAsync(v.tile_ticket + 4, 0, 2 * sizeof(uint32_t), st));
  CK(ctx, cudaMemsetAsync(v.err_flag, 0, sizeof(uint32_t), st));
  { Scope s(ctx, "junc_anchor", 1); CK(ctx, launch_junc_anchor(v, kp, st)); }
  { Scope s(ctx, "expand_count+junction_support", 1); CK(ctx, launch_expand_count(v, st)); }
  { Scope s(ctx, "scan_interval_offsets", 1); CK(ctx, launch_scan_exclusive_add(v.r_ivl_cnt, v.n_reads + 1, v.scan_state, v.tile_ticket + 3, st)); }
  // the interval count sizes the next buffers: a 4-byte read-back (one of the two host syncs inside the stage)
  uint32_t n_ivl = 0;
  if (v.n_reads) {
    uint32_t err = 0;
    CK(ctx, small_d2h(ctx, &n_ivl, v.r_ivl_cnt + v.n_reads, sizeof(uint32_t), st));
    CK(ctx, small_d2h(ctx, &err, v.err_flag, sizeof(uint32_t), st));
    mark("count + scan enqueued");
    CK(ctx, small_sync(ctx, st));
    mark("interval count read back");
    if (err) {
      std::string what;
      if (err & 1u) what += " r_seg_off runs backwards or past n_segs;";
      if (err & 2u) what += " r_pair_off runs backwards or past n_pairs;";
      if (err & 4u) what += " pair_idx points outside its bundle;";
      if (err & 8u) what += " rjunc_idx points outside its bundle's junction rows;";
      return fail(ctx, STG_EINVAL, "malformed batch:" + what);
    }
  }
  if (n_ivl >= 0x7fffffffu) return fail(ctx, STG_ERANGE, "batch deposits more than 2^31 coverage intervals; split it");
  v.n_ivl = n_ivl;
  v.n_chunks = ((int64_t)n_ivl + STG_CHUNK - 1) / STG_CHUNK;
  if ((int64_t)n_ivl > d->ivl_capacity) {
    const int64_t cap = (int64_t)n_ivl + (int64_t)n_ivl / 16 + 1024;
    const int64_t ccap = (cap + STG_CHUNK - 1) / STG_CHUNK + 1;
    int rc;
    const int64_t dcap = 2 * cap / STG_DEEP_EVENTS + 1;
    if ((rc = dev_alloc(ctx, d, &v.ivl, cap)) || (rc = dev_alloc(ctx, d, &v.events, 2 * cap)) ||
        (rc = dev_alloc(ctx, d, &v.events2, 2 * cap)) || (rc = dev_alloc(ctx, d, &v.deep_list, dcap)) ||
        (rc = dev_alloc(ctx, d, &v.deep_d, dcap * 3 * STG_TILE_W)) || (rc = dev_alloc(ctx, d, &v.deep_long, 2 * cap / 64 + 1)) || ((v.deep_long_cap = (uint32_t)(2 * cap / 64 + 1)), 0) ||
        (rc = dev_alloc(ctx, d, &v.ch_tlo, ccap)) || (rc = dev_alloc(ctx, d, &v.ch_thi, ccap)) ||
        (rc = dev_alloc(ctx, d, &v.ch_pmax_thi, ccap)) || (rc = dev_alloc(ctx, d, &v.ch_smin_tlo, ccap)) ||
        (rc = dev_alloc(ctx, d, &v.ch_row_off, ccap + 1)))
      return rc;
    if (ccap + 1 > d->scan_capacity) {  // scan scratch must cover the chunk arrays too
      uint32_t* sp = nullptr;
      if ((rc = dev_alloc(ctx, d, &sp, scan_partials_needed(ccap + 1)))) return rc;
      v.scan_partials = sp;
      d->scan_capacity = ccap + 1;
    }
    d->ivl_capacity = cap;
  }
  mark("interval buffers allocated");
  { Scope s(ctx, "expand_write", 1); CK(ctx, launch_expand_write(v, st)); }
  { Scope s(ctx, "chunk_hull+scans", 10); CK(ctx, launch_chunk_hull(v, st)); }
  // second read-back: the size of the (chunk x tile) counter matrix
  uint32_t n_cells = 0;
  if (v.n_chunks) {
    CK(ctx, small_d2h(ctx, &n_cells, v.ch_row_off + v.n_chunks, sizeof(uint32_t), st));
    CK(ctx, small_sync(ctx, st));
    mark("cell count read back");
  }
  v.n_cells = n_cells;
  if ((int64_t)n_cells > d->cell_capacity) {
    const int64_t cap = (int64_t)n_cells + (int64_t)n_cells / 16 + 1024;
    int rc;
    if ((rc = dev_alloc(ctx, d, &v.cells, cap))) return rc;
    d->cell_capacity = cap;
  }
  if (n_cells) CK(ctx, cudaMemsetAsync(v.cells, 0, sizeof(uint32_t) * (size_t)n_cells, st));
  { Scope s(ctx, "bin_count", 1); CK(ctx, launch_bin_count(v, st)); }
  { Scope s(ctx, "bin_columns+scans", 7); CK(ctx, launch_bin_columns(v, st)); }
  { Scope s(ctx, "bin_scatter", 1); CK(ctx, launch_bin_scatter(v, st)); }
  { Scope s(ctx, "tile_plan", 8); CK(ctx, launch_tile_plan(v, st)); }
  { Scope s(ctx, "deep_deposit", 2); CK(ctx, launch_deep_deposit(v, ctx->num_sms, st)); }
  { Scope s(ctx, "cov_tile", 1); CK(ctx, launch_cov_tiles(v, ctx->num_sms, st)); }
  { Scope s(ctx, "cov_fill", 1); CK(ctx, launch_cov_fill(v, ctx->num_sms, st)); }
  CK(ctx, cudaEventRecord(d->ev_cov, st));   // bpcov is final from here on (nothing later writes it)
  { Scope s(ctx, "cov_totals", 1); CK(ctx, launch_cov_totals(v, st)); }
  mark("tile kernels enqueued");
  {
    JuncParams jp{ctx->params.junctionsupport, ctx->params.sserror, ctx->params.junctionthr, ctx->params.eonly};
    Scope s(ctx, "junction_pass", 1);
    CK(ctx, launch_junc_pass(v, jp, st));
  }
  d->ran = true;
  mark("junction pass enqueued");
  return STG_OK;
}

int stg_sync(stg_ctx* ctx) {
  if (!ctx) return fail(nullptr, STG_EINVAL, "stg_sync: null ctx");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

// ---- results ------------------------------------------------------------------------------------------------
int stg_cov_layout(const stg_dbatch* d, int64_t* cov_off, int64_t* stride, int64_t* total_floats) {
  if (!d) return fail(nullptr, STG_EINVAL, "stg_cov_layout: null batch");
  for (int64_t b = 0; b < d->v.n_bundles; b++) {
    if (cov_off) cov_off[b] = d->h_cov_off[b];
    if (stride) stride[b] = d->h_stride[b];
  }
  if (total_floats) *total_floats = 3 * d->v.total_pos;
  return STG_OK;
}

const float* stg_device_bpcov(const stg_dbatch* d) { return d ? d->v.bpcov : nullptr; }

int stg_fetch_bpcov(stg_ctx* ctx, const stg_dbatch* d, int64_t b, float* out) {
  if (!ctx || !d || !out || b < 0 || b >= d->v.n_bundles) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t len = d->h_len[b];
  CK(ctx, cudaMemcpy2DAsync(out, len * sizeof(float), d->v.bpcov + d->h_cov_off[b], (size_t)d->h_stride[b] * sizeof(float),
                            len * sizeof(float), 3, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

int stg_fetch_bpcov_all(stg_ctx* ctx, const stg_dbatch* d, float* out) {
  if (!ctx || !d || !out) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_all: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_all: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(out, d->v.bpcov, sizeof(float) * 3 * (size_t)d->v.total_pos, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

int stg_upload_wait(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_upload_wait: null argument");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaEventSynchronize(d->ev_uploaded));
  return STG_OK;
}

int stg_fetch_bpcov_all_begin(stg_ctx* ctx, stg_dbatch* d, float* out) {
  if (!ctx || !d || !out) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_all_begin: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_all_begin: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamWaitEvent(ctx->out_stream, d->ev_cov, 0));
  d->out_pending = true;
  CK(ctx, copy_chunked(out, d->v.bpcov, sizeof(float) * 3 * (size_t)d->v.total_pos, cudaMemcpyDeviceToHost, ctx->out_stream));
  return STG_OK;
}

int stg_fetch_bpcov_lane_begin(stg_ctx* ctx, stg_dbatch* d, int lane, float* out) {
  if (!ctx || !d || !out || lane < 0 || lane > 2) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_lane_begin: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_lane_begin: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamWaitEvent(ctx->out_stream, d->ev_cov, 0));
  d->out_pending = true;
  CK(ctx, copy_chunked(out, d->v.bpcov + (size_t)lane * d->v.cov_pitch, sizeof(float) * (size_t)d->v.total_pos, cudaMemcpyDeviceToHost, ctx->out_stream));
  return STG_OK;
}

int stg_fetch_bpcov_bundle_lane(stg_ctx* ctx, const stg_dbatch* d, int64_t b, int lane, float* out) {
  if (!ctx || !d || !out || b < 0 || b >= d->v.n_bundles || lane < 0 || lane > 2)
    return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_bundle_lane: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_bundle_lane: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(out, d->v.bpcov + (size_t)lane * d->v.cov_pitch + d->h_cov_off[b], sizeof(float) * (size_t)d->h_len[b],
                          cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

// bpcov packed for the trip to the host (kernels K11b, cov_kernels.cu)
int stg_pack_bpcov(stg_ctx* ctx, stg_dbatch* d, int64_t* n_tiles, int64_t* n_raw) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_pack_bpcov: null argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_pack_bpcov: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const DevBatchView& v = d->v;
  const int64_t nt3 = 3 * v.n_tiles;
  if (n_tiles) *n_tiles = v.n_tiles;
  if (n_raw) *n_raw = 0;
  d->pk_nraw = 0;
  if (nt3 == 0) return STG_OK;
  StageScope stage_scope(ctx);
  int rc = STG_OK;
  stg_dbatch tmp;
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { cudaStreamSynchronize(c->stream); free_batch_allocs(c, t); } } release{ctx, &tmp};
  uint32_t *slot = nullptr, *ticket = nullptr; unsigned long long* state = nullptr;
  if ((rc = dev_alloc(ctx, &tmp, &slot, nt3 + 1 + 16)) || (rc = dev_alloc(ctx, &tmp, &ticket, 1)) || (rc = dev_alloc(ctx, &tmp, &state, nt3 / 4096 + 4))) return rc;
  if (!d->pk_desc) d->pk_mark = d->allocs.size();
  if (!d->pk_desc && ((rc = dev_alloc(ctx, d, &d->pk_desc, nt3)) || (rc = dev_alloc(ctx, d, &d->pk_val, nt3)))) return rc;
  uint32_t total = 0;
  {
    Scope s(ctx, "cov_pack_flag+scan", 2);
    CK(ctx, cudaMemsetAsync(slot + nt3, 0, sizeof(uint32_t), st));
    CK(ctx, launch_cov_pack_flag(v, slot, d->pk_val, st));
    CK(ctx, launch_scan_exclusive_add(slot, nt3 + 1, state, ticket, st));
  }
  CK(ctx, small_d2h(ctx, &total, slot + nt3, 4, st));
  CK(ctx, small_sync(ctx, st));
  if ((int64_t)total > d->pk_raw_cap) {
    const int64_t cap = (int64_t)total + (int64_t)total / 8 + 64;
    if ((rc = dev_alloc(ctx, d, &d->pk_raw, cap * STG_TILE_W))) return rc;
    d->pk_raw_cap = cap;
  }
  {
    Scope s(ctx, "cov_pack_gather", 1);
    CK(ctx, launch_cov_pack_gather(v, slot, d->pk_desc, d->pk_raw, st));
  }
  CK(ctx, cudaEventRecord(d->ev_cov, st));   // the copy-out stream waits for this event
  d->pk_nraw = total;
  if (n_raw) *n_raw = total;
  return STG_OK;
}

int stg_fetch_bpcov_packed_begin(stg_ctx* ctx, stg_dbatch* d, uint32_t* desc, float* val, float* raw) {
  if (!ctx || !d || !desc || !val) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_packed_begin: bad argument");
  if (d->pk_nraw < 0) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_packed_begin: stg_pack_bpcov has not run on this batch");
  if (d->pk_nraw && !raw) return fail(ctx, STG_EINVAL, "stg_fetch_bpcov_packed_begin: bad argument");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamWaitEvent(ctx->out_stream, d->ev_cov, 0));
  d->out_pending = true;
  const size_t nt3 = 3 * (size_t)d->v.n_tiles;
  if (nt3) {
    CK(ctx, copy_chunked(desc, d->pk_desc, sizeof(uint32_t) * nt3, cudaMemcpyDeviceToHost, ctx->out_stream));
    CK(ctx, copy_chunked(val, d->pk_val, sizeof(float) * nt3, cudaMemcpyDeviceToHost, ctx->out_stream));
  }
  if (d->pk_nraw) CK(ctx, copy_chunked(raw, d->pk_raw, sizeof(float) * (size_t)d->pk_nraw * STG_TILE_W, cudaMemcpyDeviceToHost, ctx->out_stream));
  return STG_OK;
}

// host only: BundleData::bpcov of bundles [b0, b1) from the packed form, laid out like stg_fetch_bpcov_all's buffer
} // extern "C"
namespace {
// one strand lane of one bundle out of the packed form; STREAM: the destination is written once and read much later
// (streaming stores, no read-for-ownership traffic), else it is the caller's working array (plain stores: it stays in cache)
template <bool STREAM>
inline bool unpack_lane(const stg_dbatch* d, int64_t b, int s, const uint32_t* desc, const float* val, const float* raw, float* o) {
  const size_t nt = (size_t)d->v.n_tiles;
  const uint32_t t0 = d->h_tile_off[b], t1 = d->h_tile_off[b + 1], len = d->h_len[b];
  for (uint32_t t = t0; t < t1; t++) {
    const uint32_t p0 = (t - t0) * STG_TILE_W, n = len - p0 < STG_TILE_W ? len - p0 : STG_TILE_W;
    const uint32_t ds = desc[s * nt + t];
    uint32_t i = 0;
    if (ds == 0xffffffffu) {
      const float x = val[s * nt + t];
#if defined(__SSE2__)
      {
        const __m128 x4 = _mm_set1_ps(x);
        if (STREAM && (((uintptr_t)(o + p0)) & 15u) == 0) for (; i + 4 <= n; i += 4) _mm_stream_ps(o + p0 + i, x4);
        else for (; i + 4 <= n; i += 4) _mm_storeu_ps(o + p0 + i, x4);
      }
#endif
      for (; i < n; i++) o[p0 + i] = x;
    } else {
      if (!raw) return false;
      const float* r = raw + (size_t)ds * STG_TILE_W;
#if defined(__SSE2__)
      if (STREAM && ((((uintptr_t)(o + p0)) | (uintptr_t)r) & 15u) == 0) for (; i + 4 <= n; i += 4) _mm_stream_ps(o + p0 + i, _mm_load_ps(r + i));
#endif
      if (i == 0) memcpy(o + p0, r, sizeof(float) * n);
      else for (; i < n; i++) o[p0 + i] = r[i];
    }
  }
  return true;
}
}  // namespace
extern "C" {

int stg_unpack_bpcov(const stg_dbatch* d, int64_t b0, int64_t b1, const uint32_t* desc, const float* val, const float* raw, float* out) {
  if (!d || !desc || !val || !out || b0 < 0 || b1 > d->v.n_bundles || b0 > b1) return fail(nullptr, STG_EINVAL, "stg_unpack_bpcov: bad argument");
  const size_t pitch = (size_t)d->v.total_pos;
  for (int64_t b = b0; b < b1; b++)
    for (int s = 0; s < 3; s++)
      if (!unpack_lane<true>(d, b, s, desc, val, raw, out + (size_t)s * pitch + (size_t)d->h_cov_off[b]))
        return fail(nullptr, STG_EINVAL, "stg_unpack_bpcov: bad argument");
#if defined(__SSE2__)
  _mm_sfence();
#endif
  return STG_OK;
}

int stg_unpack_bpcov_bundle(const stg_dbatch* d, int64_t b, const uint32_t* desc, const float* val, const float* raw, float* out) {
  if (!d || !desc || !val || !out || b < 0 || b >= d->v.n_bundles) return fail(nullptr, STG_EINVAL, "stg_unpack_bpcov_bundle: bad argument");
  for (int s = 0; s < 3; s++)
    if (!unpack_lane<false>(d, b, s, desc, val, raw, out + (size_t)s * d->h_len[b])) return fail(nullptr, STG_EINVAL, "stg_unpack_bpcov_bundle: bad argument");
  return STG_OK;
}

// What ONE worker thread of the host does with its share of a batch: bundle after bundle, BundleData::bpcov of the bundle into
// the thread's own arrays (`work`, reused from bundle to bundle as the reference reuses its BundleData: bundle.h:398-405), where
// printResults would read it.  `check` (may be NULL) receives the sum over the bundles of the last lane-1 element, so that the
// work cannot be skipped.
int stg_unpack_bpcov_stream(const stg_dbatch* d, int64_t b0, int64_t b1, const uint32_t* desc, const float* val, const float* raw,
                            float* work, int64_t work_floats, double* check) {
  if (!d || !desc || !val || !work || b0 < 0 || b1 > d->v.n_bundles || b0 > b1) return fail(nullptr, STG_EINVAL, "stg_unpack_bpcov_stream: bad argument");
  double acc = 0;
  for (int64_t b = b0; b < b1; b++) {
    const size_t len = d->h_len[b];
    if ((int64_t)(3 * len) > work_floats) return fail(nullptr, STG_ERANGE, "stg_unpack_bpcov_stream: a bundle is longer than the working arrays");
    for (int s = 0; s < 3; s++)
      if (!unpack_lane<false>(d, b, s, desc, val, raw, work + (size_t)s * len)) return fail(nullptr, STG_EINVAL, "stg_unpack_bpcov_stream: bad argument");
    acc += work[2 * len - 1];
  }
  if (check) *check = acc;
  return STG_OK;
}

int stg_pool_stats(stg_ctx* ctx, int64_t* out) {
  if (!ctx || !out) return fail(ctx, STG_EINVAL, "stg_pool_stats: null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  out[0] = ctx->n_malloc; out[1] = ctx->malloc_bytes; out[2] = ctx->n_free; out[3] = ctx->n_pool_flush; out[4] = (int64_t)ctx->pool_bytes;
  return STG_OK;
}

int stg_fetch_wait(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_fetch_wait: null argument");
  CK(ctx, cudaSetDevice(ctx->device));
  if (d->out_pending) { CK(ctx, cudaStreamSynchronize(ctx->out_stream)); d->out_pending = false; }
  return STG_OK;
}

int stg_fetch_junction_support(stg_ctx* ctx, const stg_dbatch* d, double* good, double* left, double* right) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_fetch_junction_support: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_junction_support: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t nb = sizeof(double) * (size_t)d->v.n_juncs;
  if (nb) {
    if (good) CK(ctx, cudaMemcpyAsync(good, d->v.j_nreads_good, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (left) CK(ctx, cudaMemcpyAsync(left, d->v.j_leftsupport, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (right) CK(ctx, cudaMemcpyAsync(right, d->v.j_rightsupport, nb, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

int stg_fetch_junction_pass(stg_ctx* ctx, const stg_dbatch* d, int8_t* strand, double* nreads, double* nreads_good,
                            double* leftsupport, double* rightsupport, double* mm, int32_t* ej, int8_t* good,
                            int8_t* good_strand) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_fetch_junction_pass: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_junction_pass: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)d->v.n_juncs;
  if (n) {
    cudaStream_t st = ctx->stream;
    if (strand) CK(ctx, cudaMemcpyAsync(strand, d->v.jp_strand, n, cudaMemcpyDeviceToHost, st));
    if (nreads) CK(ctx, cudaMemcpyAsync(nreads, d->v.jp_nreads, 8 * n, cudaMemcpyDeviceToHost, st));
    if (nreads_good) CK(ctx, cudaMemcpyAsync(nreads_good, d->v.jp_nreads_good, 8 * n, cudaMemcpyDeviceToHost, st));
    if (leftsupport) CK(ctx, cudaMemcpyAsync(leftsupport, d->v.jp_left, 8 * n, cudaMemcpyDeviceToHost, st));
    if (rightsupport) CK(ctx, cudaMemcpyAsync(rightsupport, d->v.jp_right, 8 * n, cudaMemcpyDeviceToHost, st));
    if (mm) CK(ctx, cudaMemcpyAsync(mm, d->v.jp_mm, 8 * n, cudaMemcpyDeviceToHost, st));
    if (ej) CK(ctx, cudaMemcpyAsync(ej, d->v.jp_ej, 4 * n, cudaMemcpyDeviceToHost, st));
    if (good) CK(ctx, cudaMemcpyAsync(good, d->v.jp_good, n, cudaMemcpyDeviceToHost, st));
    if (good_strand) CK(ctx, cudaMemcpyAsync(good_strand, d->v.jp_good_strand, n, cudaMemcpyDeviceToHost, st));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

int stg_fetch_cov_totals(stg_ctx* ctx, const stg_dbatch* d, double* totals) {
  if (!ctx || !d || !totals) return fail(ctx, STG_EINVAL, "stg_fetch_cov_totals: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_fetch_cov_totals: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  if (d->v.n_bundles)
    CK(ctx, cudaMemcpyAsync(totals, d->v.b_cov_total, sizeof(double) * 3 * (size_t)d->v.n_bundles, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

int stg_query_cov(stg_ctx* ctx, const stg_dbatch* d, int64_t n, const int32_t* bundle, const int8_t* lane,
                  const uint32_t* start, const uint32_t* end, int sign, double* out) {
  if (!ctx || !d || n < 0 || (n && (!bundle || !lane || !start || !end || !out)))
    return fail(ctx, STG_EINVAL, "stg_query_cov: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_query_cov: batch has not been run");
  if (n == 0) return STG_OK;
  for (int64_t i = 0; i < n; i++) {
    if (bundle[i] < 0 || bundle[i] >= d->v.n_bundles || lane[i] < 0 || lane[i] > 2 || end[i] < start[i] ||
        (uint64_t)end[i] + 1 >= d->h_len[bundle[i]])
      return fail(ctx, STG_EINVAL, "stg_query_cov: query outside its bundle");
  }
  CK(ctx, cudaSetDevice(ctx->device));
  int32_t* db = nullptr; int8_t* dl = nullptr; uint32_t *ds = nullptr, *de = nullptr; double* dout = nullptr;
  stg_dbatch tmp;   // temporaries come from the context's pool and go back to it on every way out
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { cudaStreamSynchronize(c->stream); free_batch_allocs(c, t); } } release{ctx, &tmp};
  int rc = STG_OK;
  if ((rc = dev_alloc(ctx, &tmp, &db, n)) || (rc = dev_alloc(ctx, &tmp, &dl, n)) || (rc = dev_alloc(ctx, &tmp, &ds, n)) ||
      (rc = dev_alloc(ctx, &tmp, &de, n)) || (rc = dev_alloc(ctx, &tmp, &dout, n))) return rc;
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);   // an early return leaves no staged read-back behind
  CK(ctx, small_h2d(ctx, db, bundle, n * 4, st));
  CK(ctx, small_h2d(ctx, dl, lane, n, st));
  CK(ctx, small_h2d(ctx, ds, start, n * 4, st));
  CK(ctx, small_h2d(ctx, de, end, n * 4, st));
  ctx->launches++;
  CK(ctx, launch_query_cov(d->v, n, db, dl, ds, de, sign, dout, st));
  CK(ctx, small_d2h(ctx, out, dout, n * 8, st));
  CK(ctx, small_sync(ctx, st));
  return STG_OK;
}

int stg_trims_capacity(uint32_t start, uint32_t end) { return end < start ? 0 : (int)((end - start + 1) / 50u) + 2; }

int stg_find_trims(stg_ctx* ctx, const stg_dbatch* d, int64_t n, const int32_t* bundle, const int8_t* sno,
                   const uint32_t* start, const uint32_t* end, const uint32_t* cap_off, uint32_t* count, uint32_t* pos,
                   float* abundance, uint8_t* is_start) {
  if (!ctx || !d || n < 0 || (n && (!bundle || !sno || !start || !end || !cap_off || !count || !pos || !abundance || !is_start)))
    return fail(ctx, STG_EINVAL, "stg_find_trims: bad argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_find_trims: batch has not been run");
  if (ctx->params.havePtFeatures) return fail(ctx, STG_EUNSUPPORTED, "stg_find_trims: point features are not supported");
  if (n == 0) return STG_OK;
  if (cap_off[0] != 0) return fail(ctx, STG_EINVAL, "stg_find_trims: cap_off[0] must be 0");
  for (int64_t i = 0; i < n; i++) {
    if (bundle[i] < 0 || bundle[i] >= d->v.n_bundles || (sno[i] != 0 && sno[i] != 2) || end[i] < start[i])
      return fail(ctx, STG_EINVAL, "stg_find_trims: bad region");
    const int64_t lo = (int64_t)start[i] - (int64_t)d->h_start[bundle[i]], hi = (int64_t)end[i] - (int64_t)d->h_start[bundle[i]];
    if (lo < 0 || hi + 1 >= (int64_t)d->h_len[bundle[i]]) return fail(ctx, STG_EINVAL, "stg_find_trims: region outside its bundle");
    if ((int64_t)cap_off[i + 1] - (int64_t)cap_off[i] < stg_trims_capacity(start[i], end[i]))
      return fail(ctx, STG_EINVAL, "stg_find_trims: cap_off leaves a region less room than stg_trims_capacity");
  }
  CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);   // an early return leaves no staged read-back behind
  const size_t tot = cap_off[n];
  int32_t* db = nullptr; int8_t* ds = nullptr; uint32_t *da = nullptr, *dz = nullptr, *dc = nullptr, *dn = nullptr, *dp = nullptr;
  float* df = nullptr; uint8_t* dt = nullptr;
  stg_dbatch tmp;   // temporaries come from the context's pool and go back to it on every way out
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { cudaStreamSynchronize(c->stream); free_batch_allocs(c, t); } } release{ctx, &tmp};
  int rc = STG_OK;
  if ((rc = dev_alloc(ctx, &tmp, &db, n)) || (rc = dev_alloc(ctx, &tmp, &ds, n)) || (rc = dev_alloc(ctx, &tmp, &da, n)) ||
      (rc = dev_alloc(ctx, &tmp, &dz, n)) || (rc = dev_alloc(ctx, &tmp, &dc, n + 1)) || (rc = dev_alloc(ctx, &tmp, &dn, n)) ||
      (rc = dev_alloc(ctx, &tmp, &dp, (int64_t)tot + 1)) || (rc = dev_alloc(ctx, &tmp, &df, (int64_t)tot + 1)) ||
      (rc = dev_alloc(ctx, &tmp, &dt, (int64_t)tot + 1))) return rc;
  CK(ctx, small_h2d(ctx, db, bundle, n * 4, st));
  CK(ctx, small_h2d(ctx, ds, sno, n, st));
  CK(ctx, small_h2d(ctx, da, start, n * 4, st));
  CK(ctx, small_h2d(ctx, dz, end, n * 4, st));
  CK(ctx, small_h2d(ctx, dc, cap_off, (n + 1) * 4, st));
  CK(ctx, cudaMemsetAsync(dp, 0, (tot + 1) * 4, st));
  CK(ctx, cudaMemsetAsync(df, 0, (tot + 1) * 4, st));
  CK(ctx, cudaMemsetAsync(dt, 0, tot + 1, st));
  ctx->launches++;
  {
    Scope s(ctx, "find_trims", 0);
    CK(ctx, launch_find_trims(d->v, n, db, ds, da, dz, dc, dn, dp, df, dt, st));
  }
  CK(ctx, small_d2h(ctx, count, dn, n * 4, st));
  if (tot) {
    CK(ctx, small_d2h(ctx, pos, dp, tot * 4, st));
    CK(ctx, small_d2h(ctx, abundance, df, tot * 4, st));
    CK(ctx, small_d2h(ctx, is_start, dt, tot, st));
  }
  CK(ctx, small_sync(ctx, st));
  return STG_OK;
}

// Self-test of the integer form of ordered float sums (exact_chain.cuh): list i = x[off[i] .. off[i+1]) added to s0[i] in
// order by one warp; the result must carry the bits of the plain loop.
int stg_selftest_chain(stg_ctx* ctx, int64_t n, const uint32_t* off, const float* x, const float* s0, float* out, int plain, float* kernel_ms) {
  if (!ctx || n < 0 || (n && (!off || !s0 || !out))) return fail(ctx, STG_EINVAL, "stg_selftest_chain: bad argument");
  if (n == 0) return STG_OK;
  if (off[0] != 0) return fail(ctx, STG_EINVAL, "stg_selftest_chain: off[0] must be 0");
  for (int64_t i = 0; i < n; i++) if (off[i + 1] < off[i]) return fail(ctx, STG_EINVAL, "stg_selftest_chain: offsets must not decrease");
  const int64_t tot = off[n];
  if (tot && !x) return fail(ctx, STG_EINVAL, "stg_selftest_chain: bad argument");
  CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);
  stg_dbatch tmp;
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { cudaStreamSynchronize(c->stream); free_batch_allocs(c, t); } } release{ctx, &tmp};
  uint32_t* doff = nullptr; float *dx = nullptr, *ds = nullptr, *dout = nullptr;
  int rc = STG_OK;
  if ((rc = dev_alloc(ctx, &tmp, &doff, n + 1)) || (rc = dev_alloc(ctx, &tmp, &dx, tot + 1)) || (rc = dev_alloc(ctx, &tmp, &ds, n)) ||
      (rc = dev_alloc(ctx, &tmp, &dout, n))) return rc;
  CK(ctx, cudaMemcpyAsync(doff, off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
  if (tot) CK(ctx, cudaMemcpyAsync(dx, x, (size_t)tot * 4, cudaMemcpyHostToDevice, st));
  CK(ctx, cudaMemcpyAsync(ds, s0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  ctx->launches++;
  cudaEvent_t e0 = take_event(ctx), e1 = take_event(ctx);
  CK(ctx, cudaEventRecord(e0, st));
  CK(ctx, launch_chain_selftest(n, doff, dx, ds, dout, plain, st));
  CK(ctx, cudaEventRecord(e1, st));
  CK(ctx, cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(ctx, cudaStreamSynchronize(st));
  if (kernel_ms) CK(ctx, cudaEventElapsedTime(kernel_ms, e0, e1));
  return STG_OK;
}

// ---- rows a8 + a9 ---------------------------------------------------------------------------------------------
int stg_build_groups(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_build_groups: null argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_build_groups: batch has not been run");
  if (d->groups_built) return fail(ctx, STG_EINVAL, "stg_build_groups: groups of this batch are already built");
  const stg_params& pr = ctx->params;
  if (pr.longreads || pr.mixedMode || pr.viral || pr.guided)
    return fail(ctx, STG_EUNSUPPORTED, "stg_build_groups: only the short-read de novo path is implemented");
  CK(ctx, cudaSetDevice(ctx->device));
  { static int dbg = -1; if (dbg < 0) { dbg = getenv("STGPU_GROUPS_DEBUG") ? 1 : 0; groups_set_debug(dbg); } }
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);   // an early return leaves no staged read-back behind
  DevBatchView& v = d->v;
  GroupsView& g = d->gv;
  memset(&g, 0, sizeof(g));
  d->stage_mark = d->allocs.size();
  g.bundledist = pr.bundledist;
  const int64_t nb = v.n_bundles, nr = v.n_reads, ns = v.n_segs, np = v.n_pairs, nj = v.n_juncs, ni = ns - nr;
  int rc = STG_OK;
  stg_dbatch tmp;   // scratch of this call: handed back to the pool at the end
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { free_batch_allocs(c, t); } } release{ctx, &tmp};
#define KEEP(ptr, count) do { if ((rc = dev_alloc(ctx, d, &(ptr), (int64_t)(count)))) return rc; } while (0)
#define TEMP(ptr, count) do { if ((rc = dev_alloc(ctx, &tmp, &(ptr), (int64_t)(count)))) return rc; } while (0)
  // the longest walks first
  std::vector<uint32_t> order((size_t)nb);
  for (int64_t b = 0; b < nb; b++) order[b] = (uint32_t)b;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return d->h_nreads[a] > d->h_nreads[b]; });
  uint32_t* d_order = nullptr;
  KEEP(d_order, nb);
  if (nb) CK(ctx, small_h2d(ctx, d_order, order.data(), sizeof(uint32_t) * nb, st));
  g.order = d_order;
  // capacities
  uint32_t* rcap = nullptr; unsigned long long* d_total = nullptr; unsigned long long* scan_state = nullptr; uint32_t* ticket = nullptr;
  TEMP(rcap, nr + 1 + 16); TEMP(d_total, 1); TEMP(ticket, 1);
  TEMP(scan_state, std::max(nr, nb) / 4096 + 4);
  KEEP(g.rj, ni);
  CK(ctx, cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
  { Scope s(ctx, "groups_caps+scan", 2);
    CK(ctx, launch_group_caps(v, rcap, d_total, g.rj, st));
    CK(ctx, launch_scan_exclusive_add(rcap, nr + 1, scan_state, ticket, st)); }
  g.rcap_off = rcap;
  unsigned long long total = 0;
  CK(ctx, small_d2h(ctx, &total, d_total, sizeof(total), st));
  CK(ctx, small_sync(ctx, st));   // also: `order` has left the host
  if (total + 3ull * (unsigned long long)nr >= 0x7fffffffull)
    return fail(ctx, STG_ERANGE, "stg_build_groups: batch needs more than 2^31 group slots; split it");
  const int64_t R = (int64_t)total;
  // mutable copies of the reads and of the junction-pass state
  KEEP(g.strand, nr); KEEP(g.seg_start, ns); KEEP(g.seg_end, ns); KEEP(g.pair, np);
  TEMP(g.jt_strand, nj); TEMP(g.jt_mm, nj);
  TEMP(g.ap_start, ni); TEMP(g.ap_end, ni); TEMP(g.ap_strand, ni); TEMP(g.ap_mm, ni);
  TEMP(g.jt_perm, nj + ni); TEMP(g.jt_inv, nj + ni);
  if (nr) CK(ctx, launch_copy(g.strand, v.r_strand, (size_t)nr, st));
  if (ns) CK(ctx, launch_copy(g.seg_start, v.seg_start, sizeof(uint32_t) * ns, st));
  if (ns) CK(ctx, launch_copy(g.seg_end, v.seg_end, sizeof(uint32_t) * ns, st));
  if (np) CK(ctx, launch_copy(g.pair, v.pair_idx, sizeof(int32_t) * np, st));
  if (nj) CK(ctx, launch_copy(g.jt_strand, v.jp_strand, (size_t)nj, st));
  if (nj) CK(ctx, launch_copy(g.jt_mm, v.jp_mm, sizeof(double) * nj, st));
  // sparse groups / colours / readgroup
  TEMP(g.g_start, R); TEMP(g.g_end, R); TEMP(g.g_color, R); TEMP(g.g_next, R); TEMP(g.g_merge, R); TEMP(g.g_grid, R);
  TEMP(g.g_cov, R); TEMP(g.g_multi, R); TEMP(g.g_alive, R); TEMP(g.eqcol, R + 3 * nr); TEMP(g.rg, R);
  KEEP(g.rg_off, nr + 1 + 16);
  g.rg_cnt = g.rg_off;
  KEEP(g.b_ng, nb + 1 + 16); KEEP(g.b_ncol, nb + 1 + 16); KEEP(g.b_napp, nb + 1 + 16);
  KEEP(g.b_startgroup, 3 * nb); KEEP(g.b_numfrag, nb); KEEP(g.b_fraglen, nb); KEEP(g.b_nbun, 3 * nb); KEEP(g.b_nbn, 3 * nb);
  CK(ctx, cudaMemsetAsync(g.rg_off, 0, sizeof(uint32_t) * (nr + 1), st));
  CK(ctx, cudaMemsetAsync(g.b_ng, 0, sizeof(uint32_t) * (nb + 1), st));
  CK(ctx, cudaMemsetAsync(g.b_ncol, 0, sizeof(uint32_t) * (nb + 1), st));
  CK(ctx, cudaMemsetAsync(g.b_napp, 0, sizeof(uint32_t) * (nb + 1), st));
  if (nb) {
    CK(ctx, cudaMemsetAsync(g.b_startgroup, 0xff, sizeof(int32_t) * 3 * nb, st));
    CK(ctx, cudaMemsetAsync(g.b_numfrag, 0, sizeof(double) * nb, st));
    CK(ctx, cudaMemsetAsync(g.b_fraglen, 0, sizeof(double) * nb, st));
    CK(ctx, cudaMemsetAsync(g.b_nbun, 0, sizeof(uint32_t) * 3 * nb, st));
    CK(ctx, cudaMemsetAsync(g.b_nbn, 0, sizeof(uint32_t) * 3 * nb, st));
  }
  CK(ctx, cudaMemsetAsync(v.err_flag, 0, sizeof(uint32_t), st));
  JuncParams jp{pr.junctionsupport, pr.sserror, pr.junctionthr, pr.eonly};
  int64_t n_deep = 0;   // `order` is sorted by decreasing read count
  while (n_deep < nb && (int64_t)d->h_nreads[order[n_deep]] >= groups_deep_reads()) n_deep++;
  int64_t n_cluster = 0;   // the oversized bundles are walked by a cluster of CTAs each; their ordered sums go through a log
  const int64_t cluster_reads = ctx->cluster_reads;
  while (n_cluster < n_deep && (int64_t)d->h_nreads[order[n_cluster]] >= cluster_reads) n_cluster++;
  {
    std::vector<uint32_t> voff((size_t)n_cluster + 1, 0);
    for (int64_t i = 0; i < n_cluster; i++) voff[i + 1] = voff[i] + d->h_nreads[order[i]];
    const int64_t vr = voff[n_cluster];
    uint32_t* d_voff = nullptr;
    TEMP(d_voff, n_cluster + 1);
    CK(ctx, small_h2d(ctx, d_voff, voff.data(), sizeof(uint32_t) * (n_cluster + 1), st));
    CK(ctx, small_sync(ctx, st));   // voff leaves the host
    g.vlog_off = d_voff;
    TEMP(g.vlog_grp, 6 * vr); TEMP(g.vlog_val, 6 * vr); TEMP(g.vlog_sorted, 6 * vr); TEMP(g.vlog_fl, vr); TEMP(g.vlog_nf, vr);
    uint8_t *ctl_raw = nullptr, *rec_raw = nullptr;
    TEMP(ctl_raw, (int64_t)groups_cluster_ctl_bytes() * std::max<int64_t>(n_cluster, 1));
    TEMP(rec_raw, (int64_t)groups_cluster_rec_bytes() * std::max<int64_t>(n_cluster, 1));
    g.cl_ctl = (void*)ctl_raw; g.cl_rec = (FastRec*)(void*)rec_raw;
    TEMP(g.cl_claim, std::max<int64_t>(vr, 1));
    TEMP(g.cl_ticket, 1 + 256);
    CK(ctx, cudaMemsetAsync(g.cl_ticket, 0, sizeof(uint32_t) * 257, st));
    CK(ctx, cudaMemsetAsync(g.cl_claim, 0x7f, sizeof(int32_t) * (size_t)std::max<int64_t>(vr, 1), st));
  }
  { Scope s(ctx, "groups_walk", 3);
    CK(ctx, launch_groups_walk(v, g, jp, n_cluster, n_deep, st, ctx->aux_stream, ctx->aux3_stream, ctx->ev_fork, ctx->ev_join, ctx->ev_join3)); }
  { Scope s(ctx, "groups_scans", 4);
    CK(ctx, launch_scan_exclusive_add(g.b_ng, nb + 1, scan_state, ticket, st));
    CK(ctx, launch_scan_exclusive_add(g.b_ncol, nb + 1, scan_state, ticket, st));
    CK(ctx, launch_scan_exclusive_add(g.b_napp, nb + 1, scan_state, ticket, st));
    CK(ctx, launch_scan_exclusive_add(g.rg_off, nr + 1, scan_state, ticket, st)); }
  uint32_t tot[4] = {0, 0, 0, 0}, err = 0;
  CK(ctx, small_d2h(ctx, &tot[0], g.b_ng + nb, 4, st));
  CK(ctx, small_d2h(ctx, &tot[1], g.b_ncol + nb, 4, st));
  CK(ctx, small_d2h(ctx, &tot[2], g.b_napp + nb, 4, st));
  CK(ctx, small_d2h(ctx, &tot[3], g.rg_off + nr, 4, st));
  CK(ctx, small_d2h(ctx, &err, v.err_flag, 4, st));
  CK(ctx, small_sync(ctx, st));
  if (err) return fail(ctx, STG_ERANGE, "stg_build_groups: a bundle exceeded the capacity bound of its group arrays");
  g.n_groups = tot[0]; g.n_colors = tot[1]; g.n_app = tot[2]; g.n_rg = tot[3];
  const int64_t NG = g.n_groups, NC = g.n_colors, NJ = nj + g.n_app;
  KEEP(g.d_g_start, NG); KEEP(g.d_g_end, NG); KEEP(g.d_g_color, NG); KEEP(g.d_g_next, NG); KEEP(g.d_g_merge, NG);
  KEEP(g.d_g_cov, NG); KEEP(g.d_g_multi, NG); KEEP(g.d_g_alive, NG); KEEP(g.d_color2, NG); KEEP(g.d_negp, NG);
  KEEP(g.d_eqcol, NC); KEEP(g.d_eq2, NC); KEEP(g.d_eqneg, NC); KEEP(g.d_eqpos, NC); TEMP(g.d_bundlecol, NC);
  KEEP(g.d_g2b, 3 * NG);
  KEEP(g.bun_len, 3 * NG); KEEP(g.bun_start, 3 * NG); KEEP(g.bun_last, 3 * NG); KEEP(g.bun_cov, 3 * NG); KEEP(g.bun_multi, 3 * NG);
  KEEP(g.bn_start, 3 * NG); KEEP(g.bn_end, 3 * NG); KEEP(g.bn_cov, 3 * NG); KEEP(g.bn_next, 3 * NG);
  KEEP(g.d_rg, g.n_rg);
  KEEP(g.oj_start, NJ); KEEP(g.oj_end, NJ); KEEP(g.oj_strand, NJ); KEEP(g.oj_nreads, NJ); KEEP(g.oj_nreads_good, NJ);
  KEEP(g.oj_left, NJ); KEEP(g.oj_right, NJ); KEEP(g.oj_nm, NJ); KEEP(g.oj_mm, NJ);
  { Scope s(ctx, "groups_colour+compact", 4); CK(ctx, launch_groups_finish(v, g, st)); }
#undef KEEP
#undef TEMP
  d->groups_built = true;
  return STG_OK;   // `release` parks the scratch in the pool; everything that reuses it runs on the same stream
}

// ---- row a19, de novo part ------------------------------------------------------------------------------------
int stg_neutral_predictions(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_neutral_predictions: null argument");
  if (!d->groups_built) return fail(ctx, STG_EINVAL, "stg_neutral_predictions: groups of this batch are not built");
  if (d->neutral_done) return fail(ctx, STG_EINVAL, "stg_neutral_predictions: already done for this batch");
  const stg_params& pr = ctx->params;
  if (pr.guided || pr.rawreads || pr.longreads || pr.havePtFeatures)
    return fail(ctx, STG_EUNSUPPORTED, "stg_neutral_predictions: only the short-read de novo path is implemented");
  CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);   // an early return leaves no staged read-back behind
  DevBatchView& v = d->v;
  GroupsView& g = d->gv;
  NeutralView& n = d->nv;
  memset(&n, 0, sizeof(n));
  n.mcov = pr.mcov; n.mintranscriptlen = pr.mintranscriptlen;
  const int64_t nb = v.n_bundles;
  int rc = STG_OK;
  stg_dbatch tmp;
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { free_batch_allocs(c, t); } } release{ctx, &tmp};
#define KEEP(ptr, count) do { if ((rc = dev_alloc(ctx, d, &(ptr), (int64_t)(count)))) return rc; } while (0)
#define TEMP(ptr, count) do { if ((rc = dev_alloc(ctx, &tmp, &(ptr), (int64_t)(count)))) return rc; } while (0)
  unsigned long long* scan_state = nullptr; uint32_t* ticket = nullptr;
  TEMP(ticket, 1);
  TEMP(n.reg_off, nb + 1 + 16);
  TEMP(scan_state, nb / 4096 + 4);
  { Scope s(ctx, "neutral_regions", 4);
    CK(ctx, launch_neutral_count(v, g, n.reg_off, st));
    CK(ctx, launch_scan_exclusive_add(n.reg_off, nb + 1, scan_state, ticket, st));
    uint32_t nreg = 0;
    CK(ctx, small_d2h(ctx, &nreg, n.reg_off + nb, 4, st));
    CK(ctx, small_sync(ctx, st));
    n.n_regions = nreg;
    TEMP(n.reg_cnt, nb); TEMP(n.reg_bundle, nreg); TEMP(n.reg_sno, nreg); TEMP(n.reg_start, nreg); TEMP(n.reg_end, nreg);
    TEMP(n.reg_first, nreg); TEMP(n.cap_off, (int64_t)nreg + 1 + 16); TEMP(n.trim_cnt, nreg);
    CK(ctx, cudaMemsetAsync(n.cap_off, 0, sizeof(uint32_t) * ((size_t)nreg + 1), st));
    CK(ctx, launch_neutral_regions(v, g, n, st));
    unsigned long long* scan_state2 = nullptr;
    TEMP(scan_state2, (int64_t)nreg / 4096 + 4);
    CK(ctx, launch_scan_exclusive_add(n.cap_off, (int64_t)nreg + 1, scan_state2, ticket, st)); }
  uint32_t tcap = 0;
  CK(ctx, small_d2h(ctx, &tcap, n.cap_off + n.n_regions, 4, st));
  CK(ctx, small_sync(ctx, st));
  TEMP(n.trim_pos, (int64_t)tcap + 1); TEMP(n.trim_ab, (int64_t)tcap + 1); TEMP(n.trim_st, (int64_t)tcap + 1);
  CK(ctx, cudaMemsetAsync(n.trim_cnt, 0, sizeof(uint32_t) * (size_t)std::max<int64_t>(n.n_regions, 1), st));
  CK(ctx, cudaMemsetAsync(n.trim_pos, 0, sizeof(uint32_t) * ((size_t)tcap + 1), st));
  CK(ctx, cudaMemsetAsync(n.trim_ab, 0, sizeof(float) * ((size_t)tcap + 1), st));
  CK(ctx, cudaMemsetAsync(n.trim_st, 0, (size_t)tcap + 1, st));
  { Scope s(ctx, "find_trims", 1);
    CK(ctx, launch_find_trims(v, n.n_regions, n.reg_bundle, n.reg_sno, n.reg_start, n.reg_end, n.cap_off, n.trim_cnt, n.trim_pos,
                              n.trim_ab, n.trim_st, st)); }
  const int64_t pcap = (int64_t)tcap + n.n_regions + 1;
  TEMP(n.p_start, pcap); TEMP(n.p_end, pcap); TEMP(n.p_cov, pcap); TEMP(n.p_tlen, pcap); TEMP(n.p_geneno, pcap);
  KEEP(n.pred_off, nb + 1 + 16);
  KEEP(n.b_geneno, nb);
  CK(ctx, cudaMemsetAsync(n.pred_off, 0, sizeof(uint32_t) * (nb + 1), st));
  { Scope s(ctx, "neutral_preds", 3);
    CK(ctx, launch_neutral_preds(v, g, n, st));
    CK(ctx, launch_scan_exclusive_add(n.pred_off, nb + 1, scan_state, ticket, st));
    uint32_t np = 0;
    CK(ctx, small_d2h(ctx, &np, n.pred_off + nb, 4, st));
    CK(ctx, small_sync(ctx, st));
    n.n_preds = np;
    KEEP(n.o_start, np); KEEP(n.o_end, np); KEEP(n.o_cov, np); KEEP(n.o_tlen, np); KEEP(n.o_geneno, np);
    CK(ctx, launch_neutral_compact(v, n, st)); }
#undef KEEP
#undef TEMP
  d->neutral_done = true;
  return STG_OK;
}

namespace {
struct GArr { const void* p; int64_t n; int32_t sz; };
GArr groups_array(const stg_dbatch* d, int which) {
  const DevBatchView& v = d->v;
  const GroupsView& g = d->gv;
  const int64_t nb = v.n_bundles, nr = v.n_reads, NG = g.n_groups, NC = g.n_colors, NJ = v.n_juncs + g.n_app;
  switch (which) {
    case STG_GA_R_STRAND: return {g.strand, nr, 1};
    case STG_GA_PAIR_IDX: return {g.pair, v.n_pairs, 4};
    case STG_GA_SEG_START: return {g.seg_start, v.n_segs, 4};
    case STG_GA_SEG_END: return {g.seg_end, v.n_segs, 4};
    case STG_GA_RJUNC: return {g.rj, v.n_segs - nr, 4};
    case STG_GA_GROUP_OFF: return {g.b_ng, nb + 1, 4};
    case STG_GA_COLOR_OFF: return {g.b_ncol, nb + 1, 4};
    case STG_GA_APP_OFF: return {g.b_napp, nb + 1, 4};
    case STG_GA_RG_OFF: return {g.rg_off, nr + 1, 4};
    case STG_GA_STARTGROUP: return {g.b_startgroup, 3 * nb, 4};
    case STG_GA_NUM_FRAGMENTS: return {g.b_numfrag, nb, 8};
    case STG_GA_FRAG_LEN: return {g.b_fraglen, nb, 8};
    case STG_GA_G_START: return {g.d_g_start, NG, 4};
    case STG_GA_G_END: return {g.d_g_end, NG, 4};
    case STG_GA_G_COLOR: return {g.d_g_color, NG, 4};
    case STG_GA_G_NEXT: return {g.d_g_next, NG, 4};
    case STG_GA_G_MERGE: return {g.d_g_merge, NG, 4};
    case STG_GA_G_COV: return {g.d_g_cov, NG, 4};
    case STG_GA_G_MULTI: return {g.d_g_multi, NG, 4};
    case STG_GA_G_ALIVE: return {g.d_g_alive, NG, 1};
    case STG_GA_G_COLOR2: return {g.d_color2, NG, 4};
    case STG_GA_G_NEG_PROP: return {g.d_negp, NG, 4};
    case STG_GA_EQCOL: return {g.d_eqcol, NC, 4};
    case STG_GA_EQCOL2: return {g.d_eq2, NC, 4};
    case STG_GA_EQNEG: return {g.d_eqneg, NC, 4};
    case STG_GA_EQPOS: return {g.d_eqpos, NC, 4};
    case STG_GA_G2B: return {g.d_g2b, 3 * NG, 4};
    case STG_GA_BUN_LEN: return {g.bun_len, 3 * NG, 4};
    case STG_GA_BUN_START: return {g.bun_start, 3 * NG, 4};
    case STG_GA_BUN_LAST: return {g.bun_last, 3 * NG, 4};
    case STG_GA_BUN_COV: return {g.bun_cov, 3 * NG, 4};
    case STG_GA_BUN_MULTI: return {g.bun_multi, 3 * NG, 4};
    case STG_GA_BN_START: return {g.bn_start, 3 * NG, 4};
    case STG_GA_BN_END: return {g.bn_end, 3 * NG, 4};
    case STG_GA_BN_COV: return {g.bn_cov, 3 * NG, 4};
    case STG_GA_BN_NEXT: return {g.bn_next, 3 * NG, 4};
    case STG_GA_N_BUN: return {g.b_nbun, 3 * nb, 4};
    case STG_GA_N_BN: return {g.b_nbn, 3 * nb, 4};
    case STG_GA_RG: return {g.d_rg, g.n_rg, 4};
    case STG_GA_J_START: return {g.oj_start, NJ, 4};
    case STG_GA_J_END: return {g.oj_end, NJ, 4};
    case STG_GA_J_STRAND: return {g.oj_strand, NJ, 1};
    case STG_GA_J_NREADS: return {g.oj_nreads, NJ, 8};
    case STG_GA_J_NREADS_GOOD: return {g.oj_nreads_good, NJ, 8};
    case STG_GA_J_LEFT: return {g.oj_left, NJ, 8};
    case STG_GA_J_RIGHT: return {g.oj_right, NJ, 8};
    case STG_GA_J_NM: return {g.oj_nm, NJ, 8};
    case STG_GA_J_MM: return {g.oj_mm, NJ, 8};
  }
  if (d->neutral_done) {
    const NeutralView& n = d->nv;
    switch (which) {
      case STG_GA_NP_OFF: return {n.pred_off, nb + 1, 4};
      case STG_GA_NP_START: return {n.o_start, n.n_preds, 4};
      case STG_GA_NP_END: return {n.o_end, n.n_preds, 4};
      case STG_GA_NP_COV: return {n.o_cov, n.n_preds, 8};
      case STG_GA_NP_TLEN: return {n.o_tlen, n.n_preds, 4};
      case STG_GA_NP_GENENO: return {n.o_geneno, n.n_preds, 4};
    }
  }
  return {nullptr, -1, 0};
}
}  // namespace

int stg_groups_array_info(const stg_dbatch* d, int which, int64_t* count, int32_t* elem_bytes) {
  if (!d || !d->groups_built) return fail(nullptr, STG_EINVAL, "stg_groups_array_info: groups of this batch are not built");
  const GArr a = groups_array(d, which);
  if (a.n < 0) return fail(nullptr, STG_EINVAL, "stg_groups_array_info: unknown array");
  if (count) *count = a.n;
  if (elem_bytes) *elem_bytes = a.sz;
  return STG_OK;
}

int stg_fetch_groups_array(stg_ctx* ctx, const stg_dbatch* d, int which, int64_t first, int64_t count, void* out) {
  if (!ctx || !d || !d->groups_built) return fail(ctx, STG_EINVAL, "stg_fetch_groups_array: groups of this batch are not built");
  const GArr a = groups_array(d, which);
  if (a.n < 0) return fail(ctx, STG_EINVAL, "stg_fetch_groups_array: unknown array");
  if (first < 0 || count < 0 || first + count > a.n || (count && !out)) return fail(ctx, STG_EINVAL, "stg_fetch_groups_array: bad range");
  if (count == 0) return STG_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(out, (const char*)a.p + (size_t)first * a.sz, (size_t)count * a.sz, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

// Forget the results of stg_build_groups / stg_neutral_predictions / stg_assemble of a resident batch (their device
// blocks go back to the context's pool) so that the stages can run again on it.
int stg_rewind(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_rewind: null argument");
  if (!d->groups_built) return STG_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (size_t i = d->stage_mark; i < d->allocs.size(); i++) { ctx->pool_free.push_back(d->allocs[i]); ctx->pool_bytes += d->allocs[i].bytes; }
  }
  d->allocs.resize(d->stage_mark);
  if (d->pk_desc && d->pk_mark >= d->stage_mark) { d->pk_desc = nullptr; d->pk_val = nullptr; d->pk_raw = nullptr; d->pk_raw_cap = 0; d->pk_nraw = -1; }
  d->groups_built = false; d->neutral_done = false; d->assembled = false;
  return STG_OK;
}

// sums over the batch of BundleData::num_fragments, frag_len (after stg_build_groups; 0 before) and of the clamped lane-1
// coverage, written to three doubles in DEVICE memory (asynchronous on the context stream): what a multi-GPU host
// all-reduces for the FPKM / TPM pass (stringtie.cpp:1490-1497, 915-917)
int stg_reduce_totals(stg_ctx* ctx, stg_dbatch* d, void* dev_out3) {
  if (!ctx || !d || !dev_out3) return fail(ctx, STG_EINVAL, "stg_reduce_totals: null argument");
  if (!d->ran) return fail(ctx, STG_EINVAL, "stg_reduce_totals: batch has not been run");
  CK(ctx, cudaSetDevice(ctx->device));
  ctx->launches += 1;
  CK(ctx, launch_asm_totals(d->v.n_bundles, d->groups_built ? d->gv.b_numfrag : nullptr, d->groups_built ? d->gv.b_fraglen : nullptr,
                            d->v.b_cov_total, (double*)dev_out3, ctx->stream));
  return STG_OK;
}

// ---- rows a10-a15 -----------------------------------------------------------------------------------------------
int stg_assemble(stg_ctx* ctx, stg_dbatch* d) {
  if (!ctx || !d) return fail(ctx, STG_EINVAL, "stg_assemble: null argument");
  if (!d->groups_built || !d->neutral_done) return fail(ctx, STG_EINVAL, "stg_assemble: needs stg_build_groups and stg_neutral_predictions");
  if (d->assembled) return fail(ctx, STG_EINVAL, "stg_assemble: already done for this batch");
  const stg_params& pr = ctx->params;
  if (pr.guided || pr.eonly || pr.longreads || pr.mixedMode || pr.isnascent || pr.printNascent || pr.viral || pr.rawreads || pr.havePtFeatures)
    return fail(ctx, STG_EUNSUPPORTED, "stg_assemble: only the short-read de novo path is implemented");
  CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  StageScope stage_scope(ctx);   // an early return leaves no staged read-back behind
  DevBatchView& v = d->v;
  GroupsView& g = d->gv;
  AsmView& A = d->av;
  memset(&A, 0, sizeof(A));
  AsmParams P;
  P.junctionsupport = pr.junctionsupport; P.trim = pr.trim; P.mcov = pr.mcov; P.mintranscriptlen = pr.mintranscriptlen;
  P.allowed_nodes = pr.allowed_nodes; P.isofrac = pr.isofrac; P.includesource = pr.includesource;
  const int64_t nb = v.n_bundles, nr = v.n_reads, NG = g.n_groups;
  int rc = STG_OK;
  stg_dbatch tmp;   // scratch of this call
  struct Release { stg_ctx* c; stg_dbatch* t; ~Release() { free_batch_allocs(c, t); } } release{ctx, &tmp};
#define KEEP(ptr, count) do { if ((rc = dev_alloc(ctx, d, &(ptr), (int64_t)(count)))) return rc; } while (0)
#define TEMP(ptr, count) do { if ((rc = dev_alloc(ctx, &tmp, &(ptr), (int64_t)(count)))) return rc; } while (0)
#define ZERO(ptr, count) CK(ctx, cudaMemsetAsync((ptr), 0, sizeof(*(ptr)) * (size_t)std::max<int64_t>((count), 1), st))
#define ONES(ptr, count) CK(ctx, cudaMemsetAsync((ptr), 0xff, sizeof(*(ptr)) * (size_t)std::max<int64_t>((count), 1), st))
#define SCAN(ptr, n1) CK(ctx, launch_scan_exclusive_add((ptr), (n1), scan_state, ticket, st))
#define TOTAL(dst, ptr, n) CK(ctx, small_d2h(ctx, &(dst), (ptr) + (n), 4, st))
  // graph ids and the rows of the junction table, from the per-bundle counts of rows a8 + a9
  std::vector<uint32_t> h_nbun((size_t)3 * nb + 1), h_napp((size_t)nb + 1), h_joff((size_t)nb + 1);
  if (nb) {
    CK(ctx, small_d2h(ctx, h_nbun.data(), g.b_nbun, sizeof(uint32_t) * 3 * nb, st));
    CK(ctx, small_d2h(ctx, h_joff.data(), v.b_junc_off, sizeof(uint32_t) * (nb + 1), st));
  }
  CK(ctx, small_d2h(ctx, h_napp.data(), g.b_napp, sizeof(uint32_t) * (nb + 1), st));
  CK(ctx, small_sync(ctx, st));
  std::vector<uint32_t> h_first((size_t)2 * nb + 1), h_jrow((size_t)nb + 1);
  uint32_t G = 0;
  for (int64_t b = 0; b < nb; b++) {
    h_first[2 * b] = G; G += h_nbun[3 * b];
    h_first[2 * b + 1] = G; G += h_nbun[3 * b + 2];
    h_jrow[b] = h_joff[b] + h_napp[b];
  }
  h_first[2 * nb] = G;
  h_jrow[nb] = (nb ? h_joff[nb] : 0) + h_napp[nb];
  A.n_graphs = G; A.n_junc_rows = h_jrow[nb];
  const int64_t NJ = A.n_junc_rows;
  uint32_t *d_first = nullptr, *d_jrow = nullptr;
  KEEP(d_first, 2 * nb + 1); KEEP(d_jrow, nb + 1);
  CK(ctx, small_h2d(ctx, d_first, h_first.data(), sizeof(uint32_t) * (2 * nb + 1), st));
  CK(ctx, small_h2d(ctx, d_jrow, h_jrow.data(), sizeof(uint32_t) * (nb + 1), st));
  A.gr_first = d_first; A.jrow_off = d_jrow;
  unsigned long long* scan_state = nullptr; uint32_t* ticket = nullptr;
  TEMP(ticket, 1);
  TEMP(scan_state, std::max<int64_t>(std::max<int64_t>(nr, nb), (int64_t)G) / 2048 + 8);
  KEEP(A.g_bundle, G); KEEP(A.g_k, G); KEEP(A.g_s, G);
  TEMP(A.js_start, 2 * NJ); TEMP(A.js_end, 2 * NJ); TEMP(A.je_end, 2 * NJ); TEMP(A.je_k, 2 * NJ); TEMP(A.reg_node, 2 * NJ); TEMP(A.reg_graph, 2 * NJ);
  TEMP(A.jl_cnt, 2 * nb);
  TEMP(A.cap_n, (int64_t)G + 1 + 16); TEMP(A.cap_e, (int64_t)G + 1 + 16); TEMP(A.cap_f, (int64_t)G + 1 + 16); TEMP(A.cap_t, (int64_t)G + 1 + 16);
  TEMP(A.sw_nn, G); TEMP(A.sw_ne, G); TEMP(A.sw_nf, G); TEMP(A.sw_edgeno, G);
  KEEP(A.bn_first, 3 * NG); KEEP(A.bn_cnt, 3 * NG); KEEP(A.bn_graph, 3 * NG);
  ONES(A.reg_node, 2 * NJ); ONES(A.reg_graph, 2 * NJ); ONES(A.bn_first, 3 * NG); ONES(A.bn_graph, 3 * NG); ZERO(A.bn_cnt, 3 * NG);
  CK(ctx, cudaMemsetAsync(v.err_flag, 0, sizeof(uint32_t), st));
  uint32_t tn = 0, te = 0, tf = 0, tt = 0, err = 0;
  { Scope s(ctx, "asm_lists+bounds", 7);
    CK(ctx, launch_asm_enum(v, A, st));
    CK(ctx, launch_asm_junc_lists(v, g, A, st));
    CK(ctx, launch_asm_bounds(v, g, A, P, st));
    SCAN(A.cap_n, (int64_t)G + 1); SCAN(A.cap_e, (int64_t)G + 1); SCAN(A.cap_f, (int64_t)G + 1); SCAN(A.cap_t, (int64_t)G + 1); }
  TOTAL(tn, A.cap_n, G); TOTAL(te, A.cap_e, G); TOTAL(tf, A.cap_f, G); TOTAL(tt, A.cap_t, G);
  CK(ctx, small_sync(ctx, st));
  TEMP(A.t_nstart, tn); TEMP(A.t_nend, tn); TEMP(A.t_evu, te); TEMP(A.t_evv, te); TEMP(A.t_ft1, tf); TEMP(A.t_ft2, tf); TEMP(A.t_ftab, tf);
  TEMP(A.t_trpos, tt); TEMP(A.t_trab, tt); TEMP(A.t_trst, tt);
  KEEP(A.x_node, (int64_t)G + 1 + 16); KEEP(A.x_edge, (int64_t)G + 1 + 16); TEMP(A.x_tf0, (int64_t)G + 1 + 16); KEEP(A.x_reach, (int64_t)G + 1 + 16);
  { Scope s(ctx, "asm_sweep", 6);
    CK(ctx, launch_asm_sweep(v, g, A, P, st));
    CK(ctx, launch_asm_sizes(A, st));
    SCAN(A.x_node, (int64_t)G + 1); SCAN(A.x_edge, (int64_t)G + 1); SCAN(A.x_tf0, (int64_t)G + 1); SCAN(A.x_reach, (int64_t)G + 1); }
  uint32_t xn = 0, xe = 0, xt = 0, xr = 0;
  TOTAL(xn, A.x_node, G); TOTAL(xe, A.x_edge, G); TOTAL(xt, A.x_tf0, G); TOTAL(xr, A.x_reach, G);
  CK(ctx, small_d2h(ctx, &err, v.err_flag, 4, st));
  CK(ctx, small_sync(ctx, st));
  if (err) return fail(ctx, STG_ERANGE, "stg_assemble: a graph exceeded the capacity bound of its sweep");
  A.n_nodes = xn; A.n_edges = xe; A.n_tf0cap = xt; A.n_reach = xr;
  const int64_t NN = (int64_t)xn + 2 * (int64_t)G + 2;   // per-node scratch: gno + 2 entries per graph
  { void* gp = nullptr; uint8_t* raw = nullptr; KEEP(raw, (int64_t)asm_graph_struct_bytes() * std::max<int64_t>(G, 1)); gp = raw; A.graphs = (Graph*)gp; }
  KEEP(A.n_start, xn); KEEP(A.n_end, xn); KEEP(A.n_cov, xn); KEEP(A.ch_off, xn); KEEP(A.ch_cnt, xn); KEEP(A.ch, xe); KEEP(A.pa_off, xn);
  KEEP(A.pa_cnt, xn); KEEP(A.pa, xe); KEEP(A.n_sinkpar, xn); KEEP(A.reach, xr); TEMP(A.tf_n1, xt); TEMP(A.tf_n2, xt); TEMP(A.tf_ab, xt);
  TEMP(A.fin_stack, 2 * NN); TEMP(A.fin_ncs, NN); TEMP(A.fin_visit, NN);
  KEEP(A.tf0_off, (int64_t)G + 1 + 16); KEEP(A.g_node_off, G);
  KEEP(A.r_ncov, nr + 1 + 16); KEEP(A.r_ntf, nr + 1 + 16); TEMP(A.r_nkey, nr + 1 + 16);
  { Scope s(ctx, "asm_finish+frag_count", 6);
    CK(ctx, launch_asm_finish(v, g, A, P, st));
    SCAN(A.tf0_off, (int64_t)G + 1);
    CK(ctx, launch_asm_frag_count(v, g, A, st));
    SCAN(A.r_ncov, nr + 1); SCAN(A.r_ntf, nr + 1); SCAN(A.r_nkey, nr + 1); }
  uint32_t t0 = 0, tc = 0, tr = 0, tk = 0;
  TOTAL(t0, A.tf0_off, G); TOTAL(tc, A.r_ncov, nr); TOTAL(tr, A.r_ntf, nr); TOTAL(tk, A.r_nkey, nr);
  CK(ctx, small_d2h(ctx, &err, v.err_flag, 4, st));
  CK(ctx, small_sync(ctx, st));
  if (err & 0x800u) return fail(ctx, STG_EUNSUPPORTED, "stg_assemble: the batch carries unitigs (super-reads), which are not supported");
  if (err & 0x400u) return fail(ctx, STG_ERANGE, "stg_assemble: a graph has more than allowed_nodes nodes (prune_graph_nodes is not built)");
  if (err) return fail(ctx, STG_ERANGE, "stg_assemble: a read spans more graph nodes than the kernels allow");
  A.n_tf0 = t0; A.n_covrec = tc; A.n_rec = tr; A.n_key = tk;
  const int64_t NR = (int64_t)t0 + tr;
  A.xkey_base = 2 * t0 + tk;
  TEMP(A.cov_node, tc); TEMP(A.cov_val, tc);
  TEMP(A.rec_graph, NR); TEMP(A.rec_ab, NR); TEMP(A.rec_koff, NR); TEMP(A.rec_klen, NR); TEMP(A.rec_hash, NR); TEMP(A.rec_tf, NR);
  TEMP(A.key, (int64_t)A.xkey_base + 2 * (int64_t)xe + 2 * (int64_t)G + 16);
  TEMP(A.tab_off, nb + 1 + 16);
  { Scope s(ctx, "asm_frag_write", 4);
    CK(ctx, launch_asm_frag_write(v, g, A, st));
    CK(ctx, launch_asm_table_caps(v, A, st));
    SCAN(A.tab_off, nb + 1); }
  uint32_t ts = 0;
  TOTAL(ts, A.tab_off, nb);
  CK(ctx, small_sync(ctx, st));
  TEMP(A.slot_hash, ts); TEMP(A.slot_minrec, ts); TEMP(A.slot_tf0, ts);
  ZERO(A.slot_hash, ts); ONES(A.slot_minrec, ts); ONES(A.slot_tf0, ts);
  TEMP(A.tf_rep, NR); TEMP(A.tf_abund, NR); TEMP(A.tf_graph, NR); TEMP(A.b_ntf, nb); TEMP(A.b_deep, nb);
  TEMP(A.rec_slot, tr); TEMP(A.rec_first, (int64_t)tr + 1 + 16); TEMP(A.ab_sorted, tr); TEMP(A.cov_sorted, tc);
  A.pool_cap = (unsigned long long)(((int64_t)tr + tc) / 2 + (4 << 20));
  A.deep_cap = 8192;
  TEMP(A.pool, (int64_t)A.pool_cap); TEMP(A.pool_ctr, 1); TEMP(A.deep_cnt, 2);
  TEMP(A.deep_tf_bundle, A.deep_cap); TEMP(A.deep_cov_bundle, A.deep_cap); TEMP(A.deep_tf_off, A.deep_cap); TEMP(A.deep_cov_off, A.deep_cap);
  ZERO(A.pool_ctr, 1); ZERO(A.deep_cnt, 2);
  KEEP(A.g_nt, G); TEMP(A.g_ngap, G); TEMP(A.g_reachsum, G); TEMP(A.g_fill, G);
  ZERO(A.g_nt, G); ZERO(A.g_ngap, G); ZERO(A.g_reachsum, G); ZERO(A.g_fill, G);
  TEMP(A.y_tf, (int64_t)G + 1 + 16); TEMP(A.y_pat, (int64_t)G + 1 + 16); TEMP(A.y_path, (int64_t)G + 1 + 16); TEMP(A.y_ntrf, (int64_t)G + 1 + 16);
  TEMP(A.y_istr, (int64_t)G + 1 + 16);
  unsigned long long* scan_state3 = nullptr;
  TEMP(scan_state3, (int64_t)tr / 2048 + 8);
  { Scope s(ctx, "asm_tf_table", 6);
    CK(ctx, launch_asm_tf_commit(v, A, st));
    CK(ctx, launch_scan_exclusive_add(A.rec_first, (int64_t)tr + 1, scan_state3, ticket, st));
    CK(ctx, launch_asm_tf_ids(v, A, st)); }
  { Scope s(ctx, "asm_tf_sums", 2); CK(ctx, launch_asm_tf_sums(v, A, st)); }
  { Scope s(ctx, "asm_nodecov_commit", 2); CK(ctx, launch_asm_nodecov_commit(v, A, st)); }
  { Scope s(ctx, "asm_tf_stats+sizes", 7);
    CK(ctx, launch_asm_tf_stats(v, A, st));
    SCAN(A.y_tf, (int64_t)G + 1); SCAN(A.y_pat, (int64_t)G + 1); SCAN(A.y_path, (int64_t)G + 1); SCAN(A.y_ntrf, (int64_t)G + 1); SCAN(A.y_istr, (int64_t)G + 1); }
  uint32_t yt = 0, yp = 0, yh = 0, yn = 0, yi = 0;
  TOTAL(yt, A.y_tf, G); TOTAL(yp, A.y_pat, G); TOTAL(yh, A.y_path, G); TOTAL(yn, A.y_ntrf, G); TOTAL(yi, A.y_istr, G);
  CK(ctx, small_sync(ctx, st));
  A.n_ytf = yt; A.n_ypat = yp; A.n_ypath = yh; A.n_yntrf = yn; A.n_yistr = yi;
  TEMP(A.f_koff, yt); TEMP(A.f_klen, yt); TEMP(A.f_wkoff, yt); TEMP(A.f_wklen, yt); TEMP(A.f_ab, yt); TEMP(A.f_wab, yt); TEMP(A.f_real, yt);
  TEMP(A.f_usepath, yt); TEMP(A.f_path_off, yt); TEMP(A.f_path_cnt, yt); TEMP(A.f_order, yt); TEMP(A.f_pat, yp); TEMP(A.f_skey, yt);
  TEMP(A.f_path_node, yh); TEMP(A.f_path_cont, yh); TEMP(A.f_path_ab, yh); TEMP(A.f_ntrf_off, NN); TEMP(A.f_ntrf_cnt, NN); TEMP(A.f_ntrf, yn);
  TEMP(A.f_ecov, xe);
  TEMP(A.s_nodecov, NN); TEMP(A.s_capl, NN); TEMP(A.s_capr, NN); TEMP(A.s_suml, NN); TEMP(A.s_sumr, NN); TEMP(A.s_flux, NN); TEMP(A.s_exc, NN);
  TEMP(A.s_path, NN); TEMP(A.s_succ, NN); TEMP(A.s_n2p, NN); TEMP(A.s_exs, NN); TEMP(A.s_exe, NN); TEMP(A.s_pathpat, NN); TEMP(A.s_istr, yi);
  TEMP(A.s_used, NN); TEMP(A.s_prevn, NN); TEMP(A.s_preve, xe);
  A.cap_pred = yt + xn + 16; A.cap_exon = 8 * A.cap_pred + 1024;
  TEMP(A.p_counters, 2); ZERO(A.p_counters, 2);
  TEMP(A.pp_graph, A.cap_pred); TEMP(A.pp_seq, A.cap_pred); TEMP(A.pp_tlen, A.cap_pred); TEMP(A.pp_start, A.cap_pred); TEMP(A.pp_end, A.cap_pred);
  TEMP(A.pp_exon0, A.cap_pred); TEMP(A.pp_nex, A.cap_pred); TEMP(A.pp_cov, A.cap_pred);
  TEMP(A.pe_start, A.cap_exon); TEMP(A.pe_end, A.cap_exon); TEMP(A.pe_cov, A.cap_exon);
  KEEP(A.g_npred, (int64_t)G + 1 + 16); KEEP(A.g_geneno, G); KEEP(A.o_pred_off, nb + 1);
  { Scope s(ctx, "asm_gather", 1); CK(ctx, launch_asm_gather(v, A, st)); }
  uint32_t *flow_bins = nullptr, *flow_order = nullptr;
  TEMP(flow_bins, (int64_t)asm_flow_bins_words()); TEMP(flow_order, G);
  { static int dbg = -1; if (dbg < 0) { dbg = getenv("STGPU_ASM_DEBUG") ? 1 : 0; asm_set_debug(dbg); } }
  { Scope s(ctx, "asm_flow", 4); CK(ctx, launch_asm_flow(v, A, P, flow_bins, flow_order, ctx->num_sms, st)); }
  uint32_t ctr[2] = {0, 0};
  CK(ctx, small_d2h(ctx, ctr, A.p_counters, 8, st));
  CK(ctx, small_d2h(ctx, &err, v.err_flag, 4, st));
  CK(ctx, small_sync(ctx, st));
  if (err & 0x1000u) return fail(ctx, STG_ERANGE, "stg_assemble: two transfrag keys of one bundle share a 64-bit hash");
  if (err) return fail(ctx, STG_ERANGE, "stg_assemble: a capacity bound of the flow stage was exceeded");
  A.n_pred = ctr[0]; A.n_exon = ctr[1];
  const int64_t NP = A.n_pred;
  KEEP(A.o_start, NP); KEEP(A.o_end, NP); KEEP(A.o_exon_off, NP + 1 + 16); KEEP(A.o_cov, NP); KEEP(A.o_tlen, NP); KEEP(A.o_geneno, NP);
  TEMP(A.o_src, NP); KEEP(A.o_strand, NP); KEEP(A.oe_start, A.n_exon); KEEP(A.oe_end, A.n_exon); KEEP(A.oe_cov, A.n_exon);
  unsigned long long* scan_state2 = nullptr;
  TEMP(scan_state2, NP / 2048 + 8);
  { Scope s(ctx, "asm_predictions", 5);
    SCAN(A.g_npred, (int64_t)G + 1);
    CK(ctx, launch_asm_geneno(v, A, d->nv.b_geneno, st));
    CK(ctx, launch_asm_pred_scatter(A, (uint32_t)NP, st));
    CK(ctx, launch_scan_exclusive_add(A.o_exon_off, NP + 1, scan_state2, ticket, st));
    CK(ctx, launch_asm_exon_scatter(A, (uint32_t)NP, st)); }
  CK(ctx, small_sync(ctx, st));
#undef KEEP
#undef TEMP
#undef ZERO
#undef ONES
#undef SCAN
#undef TOTAL
  d->assembled = true;
  return STG_OK;
}

namespace {
GArr asm_array(const stg_dbatch* d, int which) {
  const AsmView& A = d->av;
  const int64_t nb = d->v.n_bundles, G = A.n_graphs, NP = A.n_pred;
  switch (which) {
    case STG_AA_PRED_OFF: return {A.o_pred_off, nb + 1, 4};
    case STG_AA_P_START: return {A.o_start, NP, 4};
    case STG_AA_P_END: return {A.o_end, NP, 4};
    case STG_AA_P_COV: return {A.o_cov, NP, 8};
    case STG_AA_P_TLEN: return {A.o_tlen, NP, 4};
    case STG_AA_P_GENENO: return {A.o_geneno, NP, 4};
    case STG_AA_P_STRAND: return {A.o_strand, NP, 1};
    case STG_AA_P_EXON_OFF: return {A.o_exon_off, NP + 1, 4};
    case STG_AA_E_START: return {A.oe_start, A.n_exon, 4};
    case STG_AA_E_END: return {A.oe_end, A.n_exon, 4};
    case STG_AA_E_COV: return {A.oe_cov, A.n_exon, 4};
    case STG_AA_G_BUNDLE: return {A.g_bundle, G, 4};
    case STG_AA_G_STRAND: return {A.g_s, G, 1};
    case STG_AA_G_K: return {A.g_k, G, 4};
    case STG_AA_G_NODE_OFF: return {A.x_node, G + 1, 4};
    case STG_AA_N_START: return {A.n_start, A.n_nodes, 4};
    case STG_AA_N_END: return {A.n_end, A.n_nodes, 4};
    case STG_AA_N_COV: return {A.n_cov, A.n_nodes, 4};
    case STG_AA_G_NTF: return {A.g_nt, G, 4};
  }
  return {nullptr, -1, 0};
}
}  // namespace

int stg_asm_array_info(const stg_dbatch* d, int which, int64_t* count, int32_t* elem_bytes) {
  if (!d || !d->assembled) return fail(nullptr, STG_EINVAL, "stg_asm_array_info: batch is not assembled");
  const GArr a = asm_array(d, which);
  if (a.n < 0) return fail(nullptr, STG_EINVAL, "stg_asm_array_info: unknown array");
  if (count) *count = a.n;
  if (elem_bytes) *elem_bytes = a.sz;
  return STG_OK;
}

int stg_fetch_asm_array(stg_ctx* ctx, const stg_dbatch* d, int which, int64_t first, int64_t count, void* out) {
  if (!ctx || !d || !d->assembled) return fail(ctx, STG_EINVAL, "stg_fetch_asm_array: batch is not assembled");
  const GArr a = asm_array(d, which);
  if (a.n < 0) return fail(ctx, STG_EINVAL, "stg_fetch_asm_array: unknown array");
  if (first < 0 || count < 0 || first + count > a.n || (count && !out)) return fail(ctx, STG_EINVAL, "stg_fetch_asm_array: bad range");
  if (count == 0) return STG_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(out, (const char*)a.p + (size_t)first * a.sz, (size_t)count * a.sz, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return STG_OK;
}

// n bundles through the whole path in ONE device batch; what infer_transcripts(BundleData*) writes back comes out per bundle
int stg_infer_transcripts_batch(stg_ctx* ctx, int64_t n, const stg_bundle_view* views, stg_bundle_result* results) {
  if (!ctx || n < 0 || (n && (!views || !results))) return fail(ctx, STG_EINVAL, "stg_infer_transcripts_batch: bad argument");
  std::lock_guard<std::mutex> lk(ctx->one_bundle_mu);   // one context, one stream: callers take turns
  int rc = STG_OK;
  if (!ctx->one_bundle && (rc = stg_builder_create(ctx, &ctx->one_bundle))) return rc;
  stg_builder* b = ctx->one_bundle;
  stg_builder_reset(b);
  stg_dbatch* d = nullptr;
  stg_packed_batch pb;
  bool assemble = false;
  for (int64_t i = 0; i < n; i++) {
    results[i].n_pred = 0; results[i].n_exon = 0; results[i].num_fragments = 0; results[i].frag_len = 0;
    assemble = assemble || results[i].assemble;
    if ((rc = stg_builder_add(b, &views[i]))) return rc;
  }
  if (n == 0) return STG_OK;
  if (!(rc = stg_builder_packed(b, &pb)) && !(rc = stg_upload(ctx, &pb, &d)) && !(rc = stg_run(ctx, d))) {
    // bpcov and the junction support of every bundle that asked for them: the copies are queued together, one wait
    // (all three strand lanes: the reference's printResults reads lane 1 nearly everywhere, but its gene-coverage pass sums every
    // lane over the exons of every gene, rlink.cpp:20232-20236)
    for (int64_t i = 0; i < n && !rc; i++)
      if (results[i].bpcov) {
        const size_t len = d->h_len[i];
        if (cudaMemcpy2DAsync(results[i].bpcov, len * sizeof(float), d->v.bpcov + d->h_cov_off[i], (size_t)d->h_stride[i] * sizeof(float),
                              len * sizeof(float), 3, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
          rc = fail(ctx, STG_ECUDA, "stg_infer_transcripts_batch: copy of bpcov failed");
      }
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, STG_ECUDA, "stg_infer_transcripts_batch: stream failed");
    OneBundleOut& o = ctx->one_out;
    if (!rc && pb.n_juncs) {
      bool want = false;
      for (int64_t i = 0; i < n; i++) want = want || results[i].j_nreads_good || results[i].j_leftsupport || results[i].j_rightsupport;
      if (want) {
        o.j_good.resize(pb.n_juncs); o.j_left.resize(pb.n_juncs); o.j_right.resize(pb.n_juncs);
        rc = stg_fetch_junction_support(ctx, d, o.j_good.data(), o.j_left.data(), o.j_right.data());
        for (int64_t i = 0; i < n && !rc; i++) {
          const uint32_t j0 = pb.b_junc_off[i], nj = pb.b_junc_off[i + 1] - j0;
          if (results[i].j_nreads_good) memcpy(results[i].j_nreads_good, o.j_good.data() + j0, sizeof(double) * nj);
          if (results[i].j_leftsupport) memcpy(results[i].j_leftsupport, o.j_left.data() + j0, sizeof(double) * nj);
          if (results[i].j_rightsupport) memcpy(results[i].j_rightsupport, o.j_right.data() + j0, sizeof(double) * nj);
        }
      }
    }
    if (!rc && assemble && !(rc = stg_build_groups(ctx, d)) && !(rc = stg_neutral_predictions(ctx, d)) && !(rc = stg_assemble(ctx, d))) {
      // BundleData::pred of every bundle: its neutral bundles' predictions, then its graphs'
      const int64_t nn = d->nv.n_preds, np = d->av.n_pred, ne = d->av.n_exon;
      const int64_t NP = nn + np, NE = nn + ne;
      o.p_start.resize(NP); o.p_end.resize(NP); o.p_cov.resize(NP); o.p_tlen.resize(NP); o.p_geneno.resize(NP); o.p_strand.resize(NP);
      o.p_exon_off.resize(NP + n); o.e_start.resize(NE); o.e_end.resize(NE); o.e_cov.resize(NE);
      std::vector<uint32_t> noff(n + 1, 0), poff(n + 1, 0), us(std::max<int64_t>(nn, 1)), ue(std::max<int64_t>(nn, 1)), gs(std::max<int64_t>(np, 1)),
          ge(std::max<int64_t>(np, 1)), xo(np + 1, 0), es(std::max<int64_t>(ne, 1)), ee(std::max<int64_t>(ne, 1));
      std::vector<double> ncov(std::max<int64_t>(nn, 1)), gcov(std::max<int64_t>(np, 1)), nf(n), fl(n);
      std::vector<int32_t> ntl(std::max<int64_t>(nn, 1)), ngn(std::max<int64_t>(nn, 1)), gtl(std::max<int64_t>(np, 1)), ggn(std::max<int64_t>(np, 1));
      std::vector<int8_t> gst(std::max<int64_t>(np, 1));
      std::vector<float> ec(std::max<int64_t>(ne, 1));
      rc = stg_fetch_groups_array(ctx, d, STG_GA_NP_OFF, 0, n + 1, noff.data());
      if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_PRED_OFF, 0, n + 1, poff.data());
      if (nn) {
        if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_NP_START, 0, nn, us.data());
        if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_NP_END, 0, nn, ue.data());
        if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_NP_COV, 0, nn, ncov.data());
        if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_NP_TLEN, 0, nn, ntl.data());
        if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_NP_GENENO, 0, nn, ngn.data());
      }
      if (np) {
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_START, 0, np, gs.data());
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_END, 0, np, ge.data());
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_COV, 0, np, gcov.data());
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_TLEN, 0, np, gtl.data());
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_GENENO, 0, np, ggn.data());
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_STRAND, 0, np, gst.data());
        if (!rc) rc = stg_fetch_asm_array(ctx, d, STG_AA_P_EXON_OFF, 0, np + 1, xo.data());
        if (!rc && ne) rc = stg_fetch_asm_array(ctx, d, STG_AA_E_START, 0, ne, es.data());
        if (!rc && ne) rc = stg_fetch_asm_array(ctx, d, STG_AA_E_END, 0, ne, ee.data());
        if (!rc && ne) rc = stg_fetch_asm_array(ctx, d, STG_AA_E_COV, 0, ne, ec.data());
      }
      if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_NUM_FRAGMENTS, 0, n, nf.data());
      if (!rc) rc = stg_fetch_groups_array(ctx, d, STG_GA_FRAG_LEN, 0, n, fl.data());
      if (!rc) {
        int64_t P = 0, E = 0, X = 0;   // cursors: predictions, exons, entries of the exon-offset table (n_pred + 1 per bundle)
        for (int64_t i = 0; i < n; i++) {
          stg_bundle_result& r = results[i];
          const int64_t p0 = P, e0 = E, x0 = X;
          for (uint32_t q = noff[i]; q < noff[i + 1]; q++, P++, E++, X++) {
            o.p_start[P] = (int32_t)us[q]; o.p_end[P] = (int32_t)ue[q]; o.p_cov[P] = ncov[q]; o.p_tlen[P] = ntl[q]; o.p_geneno[P] = ngn[q];
            o.p_strand[P] = '.'; o.p_exon_off[X] = (uint32_t)(E - e0);
            o.e_start[E] = us[q]; o.e_end[E] = ue[q]; o.e_cov[E] = (float)ncov[q];
          }
          for (uint32_t q = poff[i]; q < poff[i + 1]; q++, P++, X++) {
            o.p_start[P] = (int32_t)gs[q]; o.p_end[P] = (int32_t)ge[q]; o.p_cov[P] = gcov[q]; o.p_tlen[P] = gtl[q]; o.p_geneno[P] = ggn[q];
            o.p_strand[P] = gst[q]; o.p_exon_off[X] = (uint32_t)(E - e0);
            for (uint32_t x = xo[q]; x < xo[q + 1]; x++, E++) { o.e_start[E] = es[x]; o.e_end[E] = ee[x]; o.e_cov[E] = ec[x]; }
          }
          o.p_exon_off[X++] = (uint32_t)(E - e0);
          if (!r.assemble) continue;
          r.n_pred = (int32_t)(P - p0); r.n_exon = (int32_t)(E - e0);
          r.num_fragments = nf[i]; r.frag_len = fl[i];
          r.p_start = o.p_start.data() + p0; r.p_end = o.p_end.data() + p0; r.p_cov = o.p_cov.data() + p0; r.p_tlen = o.p_tlen.data() + p0;
          r.p_geneno = o.p_geneno.data() + p0; r.p_strand = o.p_strand.data() + p0; r.p_exon_off = o.p_exon_off.data() + x0;
          r.e_start = o.e_start.data() + e0; r.e_end = o.e_end.data() + e0; r.e_cov = o.e_cov.data() + e0;
        }
      }
    }
  }
  stg_free_batch(ctx, d);
  return rc;
}

int stg_infer_transcripts(stg_ctx* ctx, const stg_bundle_view* view, stg_bundle_result* result) {
  if (!ctx || !view || !result) return fail(ctx, STG_EINVAL, "stg_infer_transcripts: null argument");
  return stg_infer_transcripts_batch(ctx, 1, view, result);
}

// ---- multi-GPU sharding (host only) -----------------------------------------------------------------------------
int stg_shard_assign(int64_t nb, const uint32_t* b_read_off, const int32_t* b_start, const int32_t* b_end, int world,
                     int32_t* rank_of, double* rank_cost) {
  if (nb < 0 || world < 1 || (nb && (!b_read_off || !b_start || !b_end || !rank_of)))
    return fail(nullptr, STG_EINVAL, "stg_shard_assign: bad argument");
  std::vector<double> cost((size_t)nb);
  std::vector<int64_t> order((size_t)nb);
  for (int64_t b = 0; b < nb; b++) {
    cost[b] = (double)(b_read_off[b + 1] - b_read_off[b]) + (double)((int64_t)b_end[b] - (int64_t)b_start[b] + 3) / 64.0;
    order[b] = b;
  }
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return cost[x] > cost[y]; });
  std::vector<double> load((size_t)world, 0.0);
  for (int64_t i = 0; i < nb; i++) {
    int best = 0;
    for (int r = 1; r < world; r++) if (load[r] < load[best]) best = r;
    rank_of[order[i]] = best;
    load[best] += cost[order[i]];
  }
  if (rank_cost) for (int r = 0; r < world; r++) rank_cost[r] = load[r];
  return STG_OK;
}

// ---- measurement hooks ------------------------------------------------------------------------------------------
int stg_set_profiling(stg_ctx* ctx, int enabled) {
  if (!ctx) return fail(nullptr, STG_EINVAL, "stg_set_profiling: null ctx");
  ctx->profiling = enabled != 0;
  return STG_OK;
}

int stg_kernel_count(const stg_ctx* ctx) { return ctx ? (int)ctx->timed.size() : 0; }

int stg_kernel_time(stg_ctx* ctx, int i, const char** name, float* ms) {
  if (!ctx || i < 0 || i >= (int)ctx->timed.size()) return fail(ctx, STG_EINVAL, "stg_kernel_time: bad index");
  if (name) *name = ctx->timed[i].name;
  if (ms) CK(ctx, cudaEventElapsedTime(ms, ctx->timed[i].e0, ctx->timed[i].e1));
  return STG_OK;
}

int64_t stg_launch_counter(const stg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int stg_batch_stats(stg_ctx* ctx, const stg_dbatch* d, int64_t* L, int64_t* S, int64_t* J) {
  if (!d) return fail(ctx, STG_EINVAL, "stg_batch_stats: null batch");
  if (L) *L = d->sum_len;
  if (S) *S = d->v.n_ivl;           // valid after stg_run
  if (J) *J = d->v.n_segs - d->v.n_reads;
  return STG_OK;
}

}  // extern "C"
