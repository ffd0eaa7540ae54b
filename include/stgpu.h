/* stgpu.h — C ABI of libstgpu, the B200 (sm_100a) engine that stands in for StringTie's per-bundle
 * assembly hot path.
 *
 * Drop-in boundary: the reference host calls
 *     int infer_transcripts(BundleData* bundle);          // rlink.h:380, call site stringtie.cpp:1477
 * one bundle at a time from a worker thread.  A host that links libstgpu flattens each BundleData
 * (bundle.h:314-424) into an `stg_bundle_view`, packs thousands of them into an `stg_batch`, and submits
 * the batch; the library returns what infer_transcripts() would have written back into BundleData.
 *
 * Everything here is plain C: pointers + sizes, no C++/torch types.  All functions return 0 on success or
 * a negative STG_E* code; stg_last_error() gives the message (the reference has no error returns: every
 * failure is GError()->exit(1), gclib/GBase.cpp:31-53 — the adaptor maps a non-zero return to GError).
 *
 * Stage coverage in this build (rows of SURVEY.md §8a):
 *   a1  boundary data model                      -> stg_bundle_view / stg_packed_batch
 *   a2  cov_edge_add / add_read_to_cov           -> stg_run(), kernels cov_expand_* / cov_tile
 *   a3  junction support loop                    -> stg_run(), kernel junc_support
 *   a4  clamped + BSIZE-blocked prefix scan      -> stg_run(), kernel cov_tile
 *   a5  get_cov / get_cov_sign                   -> stg_query_cov()
 *   a7  junction pass of build_graphs, good_junc -> stg_run(), kernel junc_pass; stg_fetch_junction_pass()
 *   a8  per-read junction repair, un-pairing     -> stg_build_groups(), kernel groups_walk
 *   a9  groups, colours, bundles / bundlenodes   -> stg_build_groups(), kernels groups_walk / groups_colour
 *   a10 create_graph (+ find_all_trims)          -> stg_assemble(), kernels asm_sweep / asm_finish; stg_find_trims()
 *   a11 read -> transfrag, update_abundance      -> stg_assemble(), kernels asm_frag_*, tf_*, nodecov_*
 *   a12 process_transfrags                       -> stg_assemble(), kernel asm_flow
 *   a13-a15 find_transcripts / push_max_flow /
 *       store_transcript                         -> stg_assemble(), kernel asm_flow; BundleData::pred via stg_fetch_asm_array()
 *   a19 (de novo part) neutral single-exon preds -> stg_neutral_predictions(), kernels neutral_* + find_trims
 *   x1  oversized bundles                        -> stg_build_groups() (thread-block clusters), stg_assemble() (CTA-wide ordered sums)
 *   b   int infer_transcripts(BundleData*)       -> stg_infer_transcripts(), stg_infer_transcripts_batch()
 * Refused with STG_EUNSUPPORTED (no CPU path behind them): guides / -e (a17), long reads / --mix (a6, a16), nascent (a18),
 * --viral, --ptf, splice-consensus filtering from the genome sequence; printResults (a20) stays on the host.
 */
#ifndef STGPU_H_
#define STGPU_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STG_ABI_VERSION 5 /* 2: groups; 3: stg_assemble, predictions; 4: copy-in / copy-out streams; 5: bpcov lane-major + packed, pool stats */
#define STG_BSIZE 10000 /* rlink.cpp:11 — block size of the blocked prefix sums in bpcov */

enum {
  STG_OK = 0,
  STG_EINVAL = -1,   /* malformed view / batch */
  STG_ECUDA = -2,    /* CUDA runtime error (message in stg_last_error) */
  STG_ENOMEM = -3,
  STG_ENODEV = -4,   /* no CUDA device: the library never falls back to the CPU */
  STG_EIO = -5,
  STG_ERANGE = -6,   /* batch exceeds a 32-bit index space: split it */
  STG_EUNSUPPORTED = -7
};

/* read flag bits (CReadAln bit-fields, bundle.h:176-177) */
#define STG_RF_UNITIG 1u
#define STG_RF_LONGREAD 2u

/* ---- options: snapshot of the globals rlink.cpp imports (rlink.cpp:16-55; defaults stringtie.cpp:122-190) */
typedef struct stg_params {
  uint32_t struct_size; /* = sizeof(stg_params), for forward compatibility */
  uint32_t junctionsupport; /* 10  anchor length (stringtie.cpp:140) */
  uint32_t sserror;         /* 25  (stringtie.cpp:141) */
  int32_t junctionthr;      /* 1   (stringtie.cpp:142) */
  uint32_t bundledist;      /* 50  (stringtie.cpp:145); -L/--mix force 0 (stringtie.cpp:1014-1023) */
  int32_t mintranscriptlen; /* 200 */
  int32_t allowed_nodes;    /* 1000 */
  float readthr;            /* 1 */
  float singlethr;          /* 4.75 */
  float isofrac;            /* 0.01 */
  float mcov;               /* 1 */
  uint8_t trim;             /* 1 */
  uint8_t eonly, nomulti, viral, mixedMode, isnascent, printNascent, guided;
  uint8_t isunitig;         /* 1 */
  uint8_t longreads, rawreads;
  uint8_t includesource;    /* 1 */
  uint8_t geneabundance, havePtFeatures;
  uint8_t have_gseq;        /* the host has the genomic sequence loaded (--ref / CRAM reference): the reference then filters
                               junctions by splice consensus (rlink.cpp:14456-14530); not built -> stg_init refuses */
  uint8_t reserved[17];
} stg_params;

/* Fill *p with the reference defaults. */
void stg_params_default(stg_params* p);

/* ---- one bundle, as the host sees it (mirror of BundleData's inputs, bundle.h:314-333) --------------
 * Coordinates are 1-based inclusive genomic positions exactly as in CReadAln::segs / CJunction.
 * Junction row 0 is the reference's null sentinel (rlink.cpp:821-824) when n_juncs > 0.              */
typedef struct stg_bundle_view {
  int32_t start, end;           /* BundleData::start / end */
  int32_t n_reads, n_juncs;
  /* reads, in readlist order (CReadAln, bundle.h:169-230) */
  const uint32_t* r_start;      /* GSeg::start */
  const uint32_t* r_end;        /* GSeg::end */
  const uint32_t* r_len;        /* CReadAln::len */
  const float* r_count;         /* CReadAln::read_count */
  const int16_t* r_nh;          /* CReadAln::nh */
  const int8_t* r_strand;       /* -1, 0, +1 */
  const uint8_t* r_flags;       /* STG_RF_* */
  const uint32_t* r_seg_off;    /* [n_reads+1] CSR into seg_* (CReadAln::segs) */
  const uint32_t* seg_start;
  const uint32_t* seg_end;
  const int32_t* rjunc_idx;     /* [n_segs - n_reads] row in j_* of CReadAln::juncs[i]; intron i of read n is
                                   entry r_seg_off[n] - n + i */
  const uint32_t* r_pair_off;   /* [n_reads+1] CSR into pair_* (CReadAln::pair_idx / pair_count) */
  const int32_t* pair_idx;      /* bundle-local read index of the mate (may be -1) */
  const float* pair_cnt;
  /* junction table, in BundleData::junction order (CJunction, bundle.h:141-166) */
  const uint32_t* j_start;
  const uint32_t* j_end;
  const int8_t* j_strand;
  const int8_t* j_guide_match;
  const double* j_nreads;
  const double* j_nm;
  const double* j_mm;
} stg_bundle_view;

/* ---- a batch of bundles, concatenated SoA (what the bundle batcher produces and the GPU consumes) ----
 * Same fields as stg_bundle_view with every array concatenated over bundles; offsets (r_seg_off, r_pair_off)
 * are GLOBAL into the concatenated arrays; pair_idx and rjunc_idx stay bundle-local.                   */
typedef struct stg_packed_batch {
  int64_t n_bundles, n_reads, n_segs, n_pairs, n_juncs;
  const int32_t* b_start;       /* [n_bundles] */
  const int32_t* b_end;         /* [n_bundles] */
  const uint32_t* b_read_off;   /* [n_bundles+1] */
  const uint32_t* b_junc_off;   /* [n_bundles+1] */
  const uint32_t* r_start;
  const uint32_t* r_end;
  const uint32_t* r_len;
  const float* r_count;
  const int16_t* r_nh;
  const int8_t* r_strand;
  const uint8_t* r_flags;
  const uint32_t* r_seg_off;    /* [n_reads+1] */
  const uint32_t* r_pair_off;   /* [n_reads+1] */
  const uint32_t* seg_start;
  const uint32_t* seg_end;
  const int32_t* rjunc_idx;     /* [n_segs - n_reads] */
  const int32_t* pair_idx;
  const float* pair_cnt;
  const uint32_t* j_start;
  const uint32_t* j_end;
  const int8_t* j_strand;
  const int8_t* j_guide_match;
  const double* j_nreads;
  const double* j_nm;
  const double* j_mm;
} stg_packed_batch;

typedef struct stg_ctx stg_ctx;         /* one per (host process, GPU) */
typedef struct stg_builder stg_builder; /* host-side bundle batcher (pinned staging) */
typedef struct stg_dbatch stg_dbatch;   /* a batch resident in HBM + its outputs */

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int stg_init(const stg_params* params, int device, stg_ctx** out);
int stg_shutdown(stg_ctx* ctx);
const char* stg_last_error(const stg_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
int stg_abi_version(void);

/* ---- bundle batcher: BundleData-shaped views -> one packed batch in pinned host memory ------------- */
int stg_builder_create(stg_ctx* ctx, stg_builder** out);
int stg_builder_add(stg_builder* b, const stg_bundle_view* view); /* copies; the view may be freed after */
int stg_builder_packed(stg_builder* b, stg_packed_batch* out);    /* borrow; valid until reset/destroy */
int stg_builder_reset(stg_builder* b);
int stg_builder_destroy(stg_builder* b);

/* ---- device batches --------------------------------------------------------------------------------
 * stg_upload: validate, plan the HBM layout (bpcov offsets, tiles), allocate, enqueue H2D copies on the
 *             context's copy-in stream.  `host` may be pageable or pinned; pinned makes the copies asynchronous, and
 *             then the arrays must stay valid until stg_upload_wait, stg_run or stg_free_batch of that batch returns.
 *             The copy-in stream is separate from the compute stream: batch i+1 may be uploaded (from another host
 *             thread, or before the stage calls of batch i) while batch i computes.
 * stg_run:    wait for the batch's upload, enqueue every kernel of the stages listed at the top of this file.  The call
 *             blocks the host twice for a 4-byte read-back that sizes its buffers (interval count, cell count: ~6 ms into a
 *             20 ms stage on a 10^8-alignment batch) and returns with the tile kernels and the junction pass still in flight;
 *             stg_build_groups / stg_assemble likewise read a few counts back between their kernels.
 * stg_sync:   wait for the context stream.                                                            */
int stg_upload(stg_ctx* ctx, const stg_packed_batch* host, stg_dbatch** out);
int stg_upload_wait(stg_ctx* ctx, stg_dbatch* batch);
int stg_run(stg_ctx* ctx, stg_dbatch* batch);
int stg_sync(stg_ctx* ctx);
int stg_free_batch(stg_ctx* ctx, stg_dbatch* batch);   /* device blocks go back to the context's pool for reuse */
/* Device memory of the batch back to the pool while the handle stays valid for stg_unpack_bpcov (a host that unpacks bpcov on its
 * worker threads while the next batch computes); stg_free_batch ends the handle. */
int stg_release_device(stg_ctx* ctx, stg_dbatch* batch);
int stg_trim(stg_ctx* ctx);                            /* release the pooled device memory (cudaFree) */

/* Outputs of the coverage / junction-support stage (count_good_junctions, rlink.cpp:16656-17136).
 * bpcov layout in HBM: for bundle b, lane s in {0:-, 1:all, 2:+}, index i in [0, len_b) with
 * len_b = end-start+3 (rlink.cpp:16875): element cov_off[b] + s*stride[b] + i  (float32).  Since ABI 5 the strand lanes
 * are batch-wide: stride[b] = total_floats / 3 for every bundle, i.e. a lane of all bundles is ONE contiguous block.   */
int stg_cov_layout(const stg_dbatch* batch, int64_t* cov_off /*[n_bundles]*/, int64_t* stride /*[n_bundles]*/,
                   int64_t* total_floats);
int stg_fetch_bpcov(stg_ctx* ctx, const stg_dbatch* batch, int64_t bundle, float* out /*[3*len_b], lane-major*/);
int stg_fetch_bpcov_all(stg_ctx* ctx, const stg_dbatch* batch, float* out /*[total_floats]*/);
/* The same copy started on the context's copy-out stream as soon as the last kernel that writes bpcov (inside stg_run)
 * has finished: it overlaps stg_build_groups / stg_neutral_predictions / stg_assemble of the same batch, none of which
 * writes bpcov.  `out` (pinned, or the copy is staged) is complete after stg_fetch_wait; stg_free_batch and a second
 * stg_run of the batch wait for it too. */
int stg_fetch_bpcov_all_begin(stg_ctx* ctx, stg_dbatch* batch, float* out /*[total_floats]*/);
/* One strand lane of every bundle (total_floats / 3 floats; bundle b at cov_off[b]) on the copy-out stream, like
 * stg_fetch_bpcov_all_begin; and one lane of one bundle, synchronously.  (The reference's printResults needs all three lanes:
 * lane 1 nearly everywhere, lanes 0 / 2 in the strand check of single-exon predictions, rlink.cpp:18453-18459, and in the gene
 * coverage pass, :20232-20236.) */
int stg_fetch_bpcov_lane_begin(stg_ctx* ctx, stg_dbatch* batch, int lane, float* out /*[total_floats / 3]*/);
int stg_fetch_bpcov_bundle_lane(stg_ctx* ctx, const stg_dbatch* batch, int64_t bundle, int lane, float* out /*[len_b]*/);
/* bpcov packed for the trip to the host.  A strand lane of a 512-position sub-tile that repeats one value (no event in it: the
 * gaps between loci, introns — four fifths of a short-read batch) travels as that value, the others as their 512 floats in one
 * dense buffer.  stg_pack_bpcov (after stg_run; synchronous: it returns the sizes) builds the packed form on the device:
 * n_tiles sub-tiles in the batch, n_raw lanes that travel in full.  stg_fetch_bpcov_packed_begin copies it on the copy-out
 * stream (complete after stg_fetch_wait): desc and val [3 * n_tiles], raw [n_raw * 512].  stg_unpack_bpcov (host only, any
 * thread) writes bpcov of the bundles [b0, b1) into a buffer laid out like stg_fetch_bpcov_all's, bit for bit what that
 * call returns for them: a fill or a copy per sub-tile lane. */
int stg_pack_bpcov(stg_ctx* ctx, stg_dbatch* batch, int64_t* n_tiles, int64_t* n_raw);
int stg_fetch_bpcov_packed_begin(stg_ctx* ctx, stg_dbatch* batch, uint32_t* desc, float* val, float* raw);
int stg_unpack_bpcov(const stg_dbatch* batch, int64_t b0, int64_t b1, const uint32_t* desc, const float* val, const float* raw,
                     float* out /*[total_floats]*/);
/* One bundle into its own arrays (lane-major, [3 * len_b]: BundleData::bpcov as stg_fetch_bpcov returns it): the call of a host
 * adaptor's worker thread; and the same for a worker's share [b0, b1) of a batch, bundle after bundle into ONE set of working
 * arrays (the reference reuses its BundleData from bundle to bundle, bundle.h:398-405), with a checksum of what was written. */
int stg_unpack_bpcov_bundle(const stg_dbatch* batch, int64_t bundle, const uint32_t* desc, const float* val, const float* raw,
                            float* out /*[3 * len_b]*/);
int stg_unpack_bpcov_stream(const stg_dbatch* batch, int64_t b0, int64_t b1, const uint32_t* desc, const float* val, const float* raw,
                            float* work, int64_t work_floats, double* check);
int stg_fetch_wait(stg_ctx* ctx, stg_dbatch* batch);
const float* stg_device_bpcov(const stg_dbatch* batch); /* device pointer (for consumers that stay on the GPU) */
/* CJunction::nreads_good / leftsupport / rightsupport after the stage; each [n_juncs] */
int stg_fetch_junction_support(stg_ctx* ctx, const stg_dbatch* batch, double* nreads_good, double* leftsupport,
                               double* rightsupport);
/* Row a7: the junction table at the end of the junction pass that opens build_graphs (rlink.cpp:14385-14809) —
 * CJunction::strand / nreads / nreads_good / leftsupport / rightsupport / mm (a negative nreads or nreads_good is the
 * negated row index of the better neighbour the junction was demoted to, rlink.cpp:14644, 14738) — the order of
 * `ejunction` (junction re-sorted by end, rlink.cpp:14387-14389) as bundle-local row indices, and the verdict of
 * good_junc (rlink.cpp:13967) on every row that still has a strand (-1 where it has none) together with the strand
 * good_junc would leave.  Each array [n_juncs]; any pointer may be NULL. */
int stg_fetch_junction_pass(stg_ctx* ctx, const stg_dbatch* batch, int8_t* strand, double* nreads, double* nreads_good,
                            double* leftsupport, double* rightsupport, double* mm, int32_t* ej, int8_t* good,
                            int8_t* good_strand);
/* per bundle: sum over the bundle of clamped coverage per lane, get_cov(s, 0, len_b-2) — [n_bundles*3] f64 */
int stg_fetch_cov_totals(stg_ctx* ctx, const stg_dbatch* batch, double* totals);

/* a5: batched get_cov (sign=0, rlink.cpp:1830) / get_cov_sign (sign=1, rlink.cpp:1842).
 * start/end are bundle-relative (pos - refstart), lane is the reference's `s` argument. Host arrays. */
int stg_query_cov(stg_ctx* ctx, const stg_dbatch* batch, int64_t n, const int32_t* bundle, const int8_t* lane,
                  const uint32_t* start, const uint32_t* end, int sign, double* out);

/* ---- rows a8 + a9: the read loop of build_graphs and what follows it up to the bundlenodes ------------------------
 * stg_build_groups runs, for every bundle of a batch that has been run (stg_run), the reference's
 *   per-read junction repair, un-pairing and un-stranding            rlink.cpp:14845-15092
 *   add_read_to_group / merge_read_to_group / merge_fwd_groups       rlink.cpp:1633-1727, 1281-1631, 1246-1278
 *   fragment counting (BundleData::num_fragments, frag_len)          rlink.cpp:15105-15116
 *   re-sort of the junction table after appended rows                rlink.cpp:15122-15125
 *   merging of groups closer than bundledist                         rlink.cpp:15134-15164
 *   colour pass over the three strand lanes (set_strandcol, neg_prop) rlink.cpp:15196-15299
 *   create_bundle / add_group_to_bundle -> CBundle, CBundlenode      rlink.cpp:15340-15427
 * on the device and keeps the results there for the stages that follow.  Short-read de novo path only: returns
 * STG_EUNSUPPORTED with longreads, mixedMode, viral or guided set.  The results are read with stg_fetch_groups_array:
 * every array is batch-wide; per-bundle slices are given by the *_OFF arrays (exclusive prefixes, [n_bundles+1]).
 *   reads after the loop (same CSR as the input): R_STRAND i8[n_reads], PAIR_IDX i32[n_pairs], SEG_START / SEG_END
 *     u32[n_segs], RJUNC i32[n_segs-n_reads] (row in the bundle's junction table BELOW)
 *   junction table after the loop, in the order the reference re-sorts it into; bundle b owns rows
 *     [b_junc_off[b] + APP_OFF[b], b_junc_off[b+1] + APP_OFF[b+1]): J_START, J_END u32, J_STRAND i8, J_NREADS,
 *     J_NREADS_GOOD, J_LEFT, J_RIGHT, J_NM, J_MM f64
 *   groups (CGroup; slot == CGroup::grid at creation), bundle b owns [GROUP_OFF[b], GROUP_OFF[b+1]): G_START, G_END
 *     u32, G_COLOR i32 (after the loop), G_NEXT i32 (next group of the strand lane, bundle-local, -1: none), G_MERGE
 *     i32 (slot it was merged into), G_COV, G_MULTI f32, G_ALIVE i8, G_COLOR2 i32 and G_NEG_PROP f32 (after the colour
 *     pass); STARTGROUP i32[3*n_bundles] first group of each lane
 *   colours, bundle b owns [COLOR_OFF[b], COLOR_OFF[b+1]): EQCOL (after the loop), EQCOL2, EQNEG, EQPOS i32
 *   readgroup: read n owns RG[RG_OFF[n] .. RG_OFF[n+1]) (i32 group slots)
 *   bundles / bundlenodes of strand lane s of bundle b: N_BUN / N_BN u32[3*n_bundles] entries from element
 *     3*GROUP_OFF[b] + s*(GROUP_OFF[b+1]-GROUP_OFF[b]) on: BUN_LEN, BUN_START, BUN_LAST i32, BUN_COV, BUN_MULTI f32
 *     (CBundle, bundle.h:260-268; start / last = bundlenode ids of the lane); BN_START, BN_END u32, BN_COV f32, BN_NEXT
 *     i32 (CBundlenode, rlink.h:56-62); G2B i32[3*groups]: group2bundle, same layout
 *   NUM_FRAGMENTS, FRAG_LEN f64[n_bundles] */
enum stg_groups_array {
  STG_GA_R_STRAND = 0, STG_GA_PAIR_IDX, STG_GA_SEG_START, STG_GA_SEG_END, STG_GA_RJUNC,
  STG_GA_GROUP_OFF, STG_GA_COLOR_OFF, STG_GA_APP_OFF, STG_GA_RG_OFF, STG_GA_STARTGROUP, STG_GA_NUM_FRAGMENTS, STG_GA_FRAG_LEN,
  STG_GA_G_START, STG_GA_G_END, STG_GA_G_COLOR, STG_GA_G_NEXT, STG_GA_G_MERGE, STG_GA_G_COV, STG_GA_G_MULTI, STG_GA_G_ALIVE,
  STG_GA_G_COLOR2, STG_GA_G_NEG_PROP, STG_GA_EQCOL, STG_GA_EQCOL2, STG_GA_EQNEG, STG_GA_EQPOS, STG_GA_G2B,
  STG_GA_BUN_LEN, STG_GA_BUN_START, STG_GA_BUN_LAST, STG_GA_BUN_COV, STG_GA_BUN_MULTI,
  STG_GA_BN_START, STG_GA_BN_END, STG_GA_BN_COV, STG_GA_BN_NEXT, STG_GA_N_BUN, STG_GA_N_BN, STG_GA_RG,
  STG_GA_J_START, STG_GA_J_END, STG_GA_J_STRAND, STG_GA_J_NREADS, STG_GA_J_NREADS_GOOD, STG_GA_J_LEFT, STG_GA_J_RIGHT,
  STG_GA_J_NM, STG_GA_J_MM,
  /* after stg_neutral_predictions: */
  STG_GA_NP_OFF, STG_GA_NP_START, STG_GA_NP_END, STG_GA_NP_COV, STG_GA_NP_TLEN, STG_GA_NP_GENENO,
  STG_GA_COUNT
};
int stg_build_groups(stg_ctx* ctx, stg_dbatch* batch);
/* element count and element size in bytes of array `which` (enum stg_groups_array) of a batch whose groups are built */
int stg_groups_array_info(const stg_dbatch* batch, int which, int64_t* count, int32_t* elem_bytes);
/* copies elements [first, first+count) of array `which` to the host buffer `out` */
int stg_fetch_groups_array(stg_ctx* ctx, const stg_dbatch* batch, int which, int64_t first, int64_t count, void* out);

/* ---- row a19 (de novo part): single-exon predictions of the neutral bundles ---------------------------------------
 * build_graphs, rlink.cpp:15500-15634, without guides / rawreads / long reads: a neutral bundle (strand lane 1) whose
 * coverage is not mostly multi-mapped (multi / cov <= mcov * 0.9) is cut, bundlenode by bundlenode, at the trim points
 * of find_all_trims; every piece longer than mintranscriptlen becomes a CPrediction (strand '.', one exon) with the mean
 * coverage of get_cov(1, ...).  Needs stg_build_groups.  Results through stg_fetch_groups_array: NP_OFF u32[n_bundles+1]
 * (bundle b owns predictions [NP_OFF[b], NP_OFF[b+1]), in the reference's order), NP_START / NP_END u32 (GSeg), NP_COV
 * f64 (CPrediction::cov; exoncov[0] is its float), NP_TLEN i32, NP_GENENO i32 (as stored: the bundle's gene counter). */
int stg_neutral_predictions(stg_ctx* ctx, stg_dbatch* batch);

/* ---- rows a10-a15: splice graphs, read -> transfrag mapping, flow decomposition, predictions --------------------------
 * stg_assemble runs, for every stranded CBundle of a batch whose groups are built and whose neutral predictions are made,
 *   create_graph (with trimnode_all, the coverage-drop links, traverse_dfs)          rlink.cpp:3579-4351, 2809, 2874
 *   get_fragment_pattern / get_read_pattern / update_abundance for every fragment   rlink.cpp:15796-15820, 4870, 4354, 4680
 *   process_transfrags (threshold, edge transfrags, trCmp sort, trf_real)            rlink.cpp:5768-6910
 *   find_transcripts / parse_trf / back_to_source_fast / fwd_to_sink_fast / push_max_flow / store_transcript
 *                                                                                    rlink.cpp:13633, 11992, 8651, 8483, 9075, 9688
 * on the device, i.e. the rest of build_graphs (rlink.cpp:15637-15935) on the short-read de novo path, and leaves the
 * CPrediction list of every bundle there.  Returns STG_EUNSUPPORTED with guides, -e, -L, --mix, -N / --nasc, --viral,
 * --ptf or rawreads set, or when the batch carries unitigs (super-reads); STG_ERANGE when a graph exceeds allowed_nodes
 * (prune_graph_nodes, rlink.cpp:3270, is not built) or a read spans more graph nodes than the kernels allow.
 * Results through stg_fetch_asm_array (batch-wide arrays):
 *   PRED_OFF u32[n_bundles+1]: bundle b owns predictions [PRED_OFF[b], PRED_OFF[b+1]) in the reference's order; they
 *     follow the bundle's neutral predictions (stg_neutral_predictions) in BundleData::pred
 *   P_START / P_END u32 (CPrediction::start / end), P_COV f64, P_TLEN i32, P_GENENO i32, P_STRAND i8 ('+' / '-'),
 *   P_EXON_OFF u32[n_pred+1] -> E_START / E_END u32 (CPrediction::exons), E_COV f32 (exoncov)
 *   stage views for tests: G_BUNDLE i32 / G_STRAND i8 / G_K i32 [n_graphs] (candidate graphs: input bundle, strand 0 '-' 1
 *   '+', CBundle index in its lane), G_NODE_OFF u32[n_graphs+1] (graphno = difference; 0: graph not built) -> N_START /
 *   N_END u32, N_COV f32 (CGraphnode incl. source and sink), G_NTF u32[n_graphs] (transfrags before process_transfrags) */
enum stg_asm_array {
  STG_AA_PRED_OFF = 0, STG_AA_P_START, STG_AA_P_END, STG_AA_P_COV, STG_AA_P_TLEN, STG_AA_P_GENENO, STG_AA_P_STRAND,
  STG_AA_P_EXON_OFF, STG_AA_E_START, STG_AA_E_END, STG_AA_E_COV,
  STG_AA_G_BUNDLE, STG_AA_G_STRAND, STG_AA_G_K, STG_AA_G_NODE_OFF, STG_AA_N_START, STG_AA_N_END, STG_AA_N_COV, STG_AA_G_NTF,
  STG_AA_COUNT
};
int stg_assemble(stg_ctx* ctx, stg_dbatch* batch);
/* Forget what stg_build_groups / stg_neutral_predictions / stg_assemble left for a resident batch (the device blocks go
 * back to the context's pool), so that these stages can run again on it (bench.py times them this way). */
int stg_rewind(stg_ctx* ctx, stg_dbatch* batch);
/* Sums over the batch of BundleData::num_fragments and frag_len (0 before stg_build_groups) and of the clamped lane-1
 * coverage, written to three f64 in DEVICE memory, asynchronously on the context stream: the values a multi-GPU host
 * all-reduces for the FPKM / TPM pass (stringtie.cpp:1490-1497, 915-917). */
int stg_reduce_totals(stg_ctx* ctx, stg_dbatch* batch, void* dev_out3);
int stg_asm_array_info(const stg_dbatch* batch, int which, int64_t* count, int32_t* elem_bytes);
int stg_fetch_asm_array(stg_ctx* ctx, const stg_dbatch* batch, int which, int64_t first, int64_t count, void* out);

/* find_all_trims (rlink.cpp:2118-2419, part of row a10; one of the bpcov range-query consumers): for n regions
 * (bundle[i], sno[i] in {0, 2} as at the call sites rlink.cpp:2813 / 15559, genomic start[i]..end[i]) of a batch that
 * has been run, the reference's `trimpoint` vector of each region: candidate source / sink trim positions with their
 * abundance; entries the reference's global re-validation rejected keep pos = 0.  Region i may return up to
 * stg_trims_capacity(start[i], end[i]) points; the caller passes cap_off = exclusive prefix of those capacities
 * ([n+1]) and arrays of cap_off[n] elements; count[i] points are written from element cap_off[i] on.  Host arrays.
 * Point features (--ptf, `tstartend`) are not supported. */
int stg_trims_capacity(uint32_t start, uint32_t end);
int stg_find_trims(stg_ctx* ctx, const stg_dbatch* batch, int64_t n, const int32_t* bundle, const int8_t* sno,
                   const uint32_t* start, const uint32_t* end, const uint32_t* cap_off, uint32_t* count,
                   uint32_t* pos, float* abundance, uint8_t* is_start);

/* Drop-in for ONE bundle, synchronous, the shape of the reference call `int infer_transcripts(BundleData*)` (rlink.h:380,
 * call site stringtie.cpp:1477): upload, every stage, fetch of what the reference writes back into BundleData.
 * With `assemble` == 0 only count_good_junctions runs (bpcov + junction support: the ABI-2 behaviour).  With `assemble`
 * != 0 the whole function runs — count_good_junctions, then build_graphs to its end — and the result carries
 * BundleData::pred (the neutral bundles' single-exon predictions, then the graphs' transcripts, in the reference's order),
 * BundleData::num_fragments and frag_len.  The prediction arrays are owned by the context and stay valid until the next
 * stg_infer_transcripts call on it. */
typedef struct stg_bundle_result {
  float* bpcov;            /* caller-allocated [3*(end-start+3)] or NULL */
  double* j_nreads_good;   /* caller-allocated [n_juncs] or NULL */
  double* j_leftsupport;
  double* j_rightsupport;
  /* ---- ABI 3 ---- */
  int32_t assemble;        /* in */
  int32_t n_pred, n_exon;  /* out */
  double num_fragments, frag_len;              /* out (rlink.cpp:15105-15116) */
  const int32_t *p_start, *p_end;              /* CPrediction::start / end (GSeg) */
  const double* p_cov;                         /* CPrediction::cov */
  const int32_t *p_tlen, *p_geneno;            /* CPrediction::tlen / geneno */
  const int8_t* p_strand;                      /* '+', '-' or '.' */
  const uint32_t* p_exon_off;                  /* [n_pred+1] */
  const uint32_t *e_start, *e_end;             /* CPrediction::exons */
  const float* e_cov;                          /* CPrediction::exoncov */
} stg_bundle_result;
int stg_infer_transcripts(stg_ctx* ctx, const stg_bundle_view* view, stg_bundle_result* result);
/* The same for n bundles in ONE device batch (what a host does that collects the bundles of its worker threads, or of its
 * bundle queue, before it calls: host/stgpu_adaptor.cpp).  results[i] describes bundle i; its p_exon_off is relative to its own e_start /
 * e_end / e_cov.  The arrays are owned by the context and stay valid until the next stg_infer_transcripts* call on it. */
int stg_infer_transcripts_batch(stg_ctx* ctx, int64_t n, const stg_bundle_view* views, stg_bundle_result* results);

/* ---- multi-GPU: bundles shard by genomic locus, one rank per GPU ---------------------------------
 * Greedy least-loaded assignment on cost = reads + span/64, heaviest bundles first (the reference has one queue
 * feeding N threads, stringtie.cpp:484-492; here the producer decides the GPU up front).  Host-only, no device
 * needed.  rank_of[b] in [0, world).  Deterministic: ties go to the lowest rank.                              */
int stg_shard_assign(int64_t n_bundles, const uint32_t* b_read_off /*[n+1]*/, const int32_t* b_start,
                     const int32_t* b_end, int world, int32_t* rank_of /*[n]*/, double* rank_cost /*[world] or NULL*/);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------
 * With profiling on, stg_run brackets every kernel with CUDA events on the context stream.            */
/* Self-test of the engine's ordered float sums: out[i] = (((s0[i] + x[off[i]]) + x[off[i]+1]) + ...) in float32 with
 * round-to-nearest at every step, i.e. what the reference's "+=" loops compute (cov_edge_add rlink.cpp:133-142,
 * update_abundance :4680), through the kernel code path the coverage and commit stages use for long sums (`plain` == 1:
 * through the bare chain of dependent adds instead; 2: the form in which a whole CTA shares one list).  kernel_ms (may be NULL) receives the kernel's duration. */
int stg_selftest_chain(stg_ctx* ctx, int64_t n, const uint32_t* off /*[n+1]*/, const float* x, const float* s0 /*[n]*/,
                       float* out /*[n]*/, int plain, float* kernel_ms);

/* Device-memory pool of the context: out[5] = {cudaMalloc calls, bytes they asked for, cudaFree calls of blocks the bounded pool
 * refused, times the whole pool was dropped to make room, bytes parked now}.  cudaMalloc / cudaFree synchronise the device:
 * a host that streams batches wants these counters to stand still. */
int stg_pool_stats(stg_ctx* ctx, int64_t* out /*[5]*/);
int stg_set_profiling(stg_ctx* ctx, int enabled);
int stg_kernel_count(const stg_ctx* ctx);                       /* kernels launched by the last stg_run */
int stg_kernel_time(stg_ctx* ctx, int i, const char** name, float* ms); /* after stg_sync */
int64_t stg_launch_counter(const stg_ctx* ctx);                 /* total kernel launches since init */
/* algorithmic work of the coverage stage for this batch:
 * L = sum of bundle lengths (end-start+3), S = mate-merged coverage intervals, J = junction-support updates */
int stg_batch_stats(stg_ctx* ctx, const stg_dbatch* batch, int64_t* L, int64_t* S, int64_t* J);
void* stg_stream(stg_ctx* ctx);                                 /* cudaStream_t of the context */

/* ---- packed-batch files (.stgb): named little-endian arrays; used by tests, bench and the oracle tools
 * file := "STGB" u32 version u32 n_arrays { char name[24]; u32 dtype; u32 pad; u64 count; bytes (padded to 8) } */
enum { STG_DT_I8 = 1, STG_DT_U8 = 2, STG_DT_I16 = 3, STG_DT_I32 = 4, STG_DT_U32 = 5, STG_DT_F32 = 6,
       STG_DT_F64 = 7, STG_DT_I64 = 8 };
#define STG_FILE_MAGIC 0x42475453u /* "STGB" */

#ifdef __cplusplus
}
#endif
#endif /* STGPU_H_ */
