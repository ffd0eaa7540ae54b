// stgpu_internal.h — shared between the C-ABI host code (stgpu_api.cu) and the kernels (cov_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/stgpu.h"

#define STG_TILE_W 512           // positions per sub-tile (one warp owns one sub-tile of all three strand lanes)
#define STG_TILE_SHIFT 9
#define STG_TILE_WARPS 32        // warps per CTA of the tile kernel (one CTA per SM; ~6.6 KB shared memory per warp)
#define STG_LANE_ALIGN 32        // every strand lane of every bundle starts on a 128-byte boundary
#define STG_CHUNK 512            // coverage intervals per binning chunk (one warp walks one chunk in order)
#define STG_READ_BLOCK 256       // reads per CTA of the per-read kernels (cta_bundle is sampled at this step)
#define STG_NO_TILE 0xffffffffu
#define STG_DEEP_EVENTS 2048u     // sub-tiles with at least this many events take the sort-by-position path

// Everything a warp needs to process one sub-tile, prepared by tile_plan_kernel (one 32-byte load).
struct __align__(32) TileRec {
  uint32_t gbase;     // global position of index 0 of its bundle
  uint32_t i0;        // in-bundle index of the first position of the sub-tile
  uint32_t e0;        // first event of the sub-tile in the binned event array
  uint32_t nev;       // number of events
  uint32_t stride;    // lane stride of the bundle
  uint32_t npos;      // valid positions (<= STG_TILE_W)
  uint32_t tile_id;   // bundle-major tile id (index of its carry / flag)
  uint32_t wait_tile; // nearest earlier sub-tile of the bundle that has events (STG_NO_TILE: none)
};

// Device view of one batch: inputs (copies of stg_packed_batch arrays), the planned layout, scratch and outputs.
struct DevBatchView {
  int64_t n_bundles, n_reads, n_segs, n_pairs, n_juncs;
  // inputs
  const int32_t *b_start, *b_end;
  const uint32_t *b_read_off, *b_junc_off;
  const uint32_t *r_start, *r_end, *r_len;
  const float* r_count;
  const int16_t* r_nh;
  const int8_t* r_strand;
  const uint8_t* r_flags;
  const uint32_t *r_seg_off, *r_pair_off, *seg_start, *seg_end;
  const int32_t* rjunc_idx;
  const int32_t* pair_idx;
  const float* pair_cnt;
  const uint32_t *j_start, *j_end;
  const int8_t *j_strand, *j_guide_match;
  const double *j_nreads, *j_nm, *j_mm;
  // layout (planned on the host in stg_upload)
  const uint32_t* b_len;       // [n_bundles]   end-start+3
  const uint32_t* b_stride;    // [n_bundles]   len rounded up to STG_LANE_ALIGN
  const uint32_t* b_gbase;     // [n_bundles+1] prefix sum of b_stride: "global position" of index 0 of bundle b
  const uint32_t* b_tile_off;  // [n_bundles+1] prefix sum of ceil(len/STG_TILE_W): bundle-major tile ids
  const uint32_t* cta_bundle;  // [ceil(n_reads/STG_READ_BLOCK)] bundle of read STG_READ_BLOCK*i
  int64_t n_rounds;            // rounds of the tile schedule: max(sub-tiles of the longest bundle, 2048)
  int64_t n_tiles;
  int64_t total_pos;           // b_gbase[n_bundles]
  // scratch
  uint32_t* err_flag;          // [1] STG_ERR_* bits raised by the kernels on malformed input
  unsigned long long* scan_state;  // [ceil((n_reads+1)/4096)] status words of the single-pass scan
  uint32_t* r_ivl_cnt;         // [n_reads+1] per-read interval count, then exclusive prefix (in place)
  int64_t n_ivl;               // filled after the count pass (host reads it back)
  uint4* ivl;                  // [n_ivl] {tile of the +w deposit, tile of the -w deposit,
                               //          pos(+w) | pos(-w) << 9 | lane << 18, w bits}, in reference order
  int64_t n_chunks;            // ceil(n_ivl / STG_CHUNK)
  uint32_t* ch_tlo;            // [n_chunks] lowest tile any event of the chunk lands in
  uint32_t* ch_thi;            // [n_chunks] highest
  uint32_t* ch_pmax_thi;       // [n_chunks] inclusive prefix max of ch_thi  (monotone: searchable)
  uint32_t* ch_smin_tlo;       // [n_chunks] inclusive suffix min of ch_tlo  (monotone: searchable)
  uint32_t* ch_row_off;        // [n_chunks+1] first cell of the chunk's row of per-tile event counts
  int64_t n_cells;             // ch_row_off[n_chunks] (host reads it back)
  uint32_t* cells;             // [n_cells] row c, column t-ch_tlo[c]: events of chunk c in tile t; after the column
                               //           scan: events of earlier chunks in tile t (the chunk's first slot)
  uint32_t* tile_evoff;        // [n_tiles+1] event count per tile, then exclusive prefix (in place)
  uint32_t* tile_lastne;       // [n_tiles]   (t+1 if tile t has events else 0), then inclusive prefix max
  uint2* events;               // [2*n_ivl]   {pos | lane << 9, signed weight bits}, tile-major, reference order
  uint2* events2;              // [2*n_ivl]   deep sub-tiles only: {lane, signed weight bits} sorted by position (stable)
  uint4* deep_list;            // [deep_cap]  {tile id, first event, events, -} of every deep sub-tile
  float* deep_d;               // [deep_cap * 3 * STG_TILE_W] their finished difference arrays
  uint4* deep_long;            // [long_cap]  {deep slot, position, first event in events2, events} of every long list
  uint32_t* deep_count;        // [4] deep sub-tiles, long lists, very long lists (stored from the far end of deep_long)
  uint32_t deep_long_cap;      // entries of deep_long
  uint32_t* scan_partials;     // scratch for the 3-phase scans
  uint8_t* junc_anchor;        // [n_juncs] anchor length each junction demands (10 or 25, rlink.cpp:16991-16997)
  uint32_t* round_off;         // [n_rounds+1] tiles with events per round, then exclusive prefix: first ticket of the
                               //              round; round_off[n_rounds] = n_main, the number of tiles with events
  uint32_t* round_cur;         // [n_rounds+1] tickets handed out inside the round; [n_rounds]: event-free tiles placed
  TileRec* tile_plan;          // [n_tiles] [0, n_main): tiles with events in ticket order; [n_main, n_tiles): the rest
  unsigned long long* tile_carry;  // [n_tiles*4] per strand lane (c | bs << 32) after the last position of a
                               //             sub-tile with events; all-ones until published (bundle-major tile id)
  uint32_t* tile_ticket;       // [6] work counters: cov_tile_kernel, deep_deposit_kernel, deep_long_kernel, scan blocks, very long lists
  // outputs
  float* bpcov;                // [3*total_pos], lane-major over the whole batch: bundle b lane s index i at s*cov_pitch + b_gbase[b] + i
  size_t cov_pitch;            // = total_pos: strand lane 1 of every bundle is ONE contiguous block (the part the host reads back)
  double *j_nreads_good, *j_leftsupport, *j_rightsupport;  // [n_juncs]
  double* b_cov_total;         // [n_bundles*3]
  // junction pass of build_graphs (row a7): CJunction fields after rlink.cpp:14385-14809, one per junction row
  int8_t* jp_strand;
  double *jp_nreads, *jp_nreads_good, *jp_left, *jp_right, *jp_mm;
  int32_t* jp_ej;              // bundle-local row of ejunction[k] (junction re-sorted by end)
  int8_t *jp_good, *jp_good_strand;  // good_junc verdict (-1: row has no strand) and the strand it leaves
};

// Rows a8 + a9 (groups_device.cuh, group_kernels.cu).  "sparse" arrays are laid out by CAPACITY: read n owns the slots
// [rcap_off[n], rcap_off[n+1]) of rg, bundle b (first read r0) owns the group slots from rcap_off[r0] and the colour
// slots from rcap_off[r0] + 3*r0; "dense" arrays are compacted by the real counts once the walk has finished.
struct GroupsView {
  uint32_t bundledist;
  const uint32_t* order;       // [n_bundles] bundles by decreasing read count (the longest walks start first)
  const uint32_t* rcap_off;    // [n_reads+1]
  // mutable copies of the reads (after the stage: the repaired reads)
  int8_t* strand; uint32_t *seg_start, *seg_end; int32_t *pair, *rj;
  // junction table state: copies of jp_strand / jp_mm, rows appended by the repair (bundle b: intron base of b)
  int8_t* jt_strand; double* jt_mm;
  uint32_t *ap_start, *ap_end; int8_t *ap_strand, *ap_mm;
  int32_t *jt_perm, *jt_inv;   // [n_juncs + n_introns] bundle b at b_junc_off[b] + intron base of b
  // very deep bundles (order[0 .. n_vdeep)): log of the ordered part of every simple read; read r of the i-th at vlog_off[i] + r
  const uint32_t* vlog_off; int32_t* vlog_grp; float2 *vlog_val, *vlog_sorted; double *vlog_fl, *vlog_nf;
  uint32_t* cl_ticket;   // [1 + clusters]: next oversized bundle, and the one every resident cluster is walking
  void* cl_ctl; struct FastRec* cl_rec; int32_t* cl_claim;   // oversized bundles (order[0 .. n_cluster)): control block, window records, claims
  // sparse groups / colours / readgroup
  uint32_t *g_start, *g_end; int32_t *g_color, *g_next, *g_merge, *g_grid; float *g_cov, *g_multi; int8_t* g_alive;
  int32_t* eqcol;
  int32_t* rg; uint32_t* rg_cnt;
  // per bundle
  uint32_t *b_ng, *b_ncol, *b_napp;   // [n_bundles+1] counts, then exclusive prefixes (in place, after the walk)
  int32_t* b_startgroup;              // [3*n_bundles]
  double *b_numfrag, *b_fraglen;      // [n_bundles]
  uint32_t* rg_off;                   // [n_reads+1] rg_cnt, then its exclusive prefix
  int64_t n_groups, n_colors, n_rg, n_app;   // totals (host fills them after the scans)
  // dense results
  uint32_t *d_g_start, *d_g_end; int32_t *d_g_color, *d_g_next, *d_g_merge; float *d_g_cov, *d_g_multi; int8_t* d_g_alive;
  int32_t* d_color2; float* d_negp;                    // [n_groups] colour / neg_prop after the colour pass
  int32_t *d_eqcol, *d_eq2, *d_eqneg, *d_eqpos, *d_bundlecol;   // [n_colors]
  int32_t* d_g2b;                                      // [3*n_groups] bundle b: 3*ng_off[b] + lane*ng + group
  int32_t *bun_len, *bun_start, *bun_last; float *bun_cov, *bun_multi;   // [3*n_groups] lane s of bundle b at 3*ng_off[b] + s*ng
  uint32_t *bn_start, *bn_end; float* bn_cov; int32_t* bn_next;          // same layout
  uint32_t *b_nbun, *b_nbn;                            // [3*n_bundles]
  int32_t* d_rg;                                       // [n_rg]
  // junction table after the stage, re-sorted as the reference does: bundle b at b_junc_off[b] + b_napp[b] (prefix)
  uint32_t *oj_start, *oj_end; int8_t* oj_strand; double *oj_nreads, *oj_nreads_good, *oj_left, *oj_right, *oj_nm, *oj_mm;
};

// Row a19, de novo part (neutral_kernels.cu): single-exon predictions of the neutral bundles
struct NeutralView {
  float mcov; int32_t mintranscriptlen;
  int64_t n_regions, n_preds;
  uint32_t* reg_off;           // [n_bundles+1] first region slot of bundle b (bound: its neutral bundlenodes)
  uint32_t* reg_cnt;           // [n_bundles]   regions in use
  int32_t* reg_bundle; int8_t* reg_sno; uint32_t *reg_start, *reg_end; uint8_t* reg_first;   // [n_regions]
  uint32_t* cap_off;           // [n_regions+1] trim point capacity, then exclusive prefix
  uint32_t *trim_cnt, *trim_pos; float* trim_ab; uint8_t* trim_st;    // find_trims results
  uint32_t *p_start, *p_end; double* p_cov; int32_t *p_tlen, *p_geneno;   // [cap_off[n_regions] + n_regions] uncompacted
  uint32_t* pred_off;          // [n_bundles+1] counts, then exclusive prefix
  uint32_t *o_start, *o_end; double* o_cov; int32_t *o_tlen, *o_geneno;   // [n_preds]
  int32_t* b_geneno;           // [n_bundles] the gene counter after the neutral bundles (where the graphs' genes start)
};

// Rows a10-a15 (asm_device.cuh, frag_device.cuh, asm_kernels.cu): splice graphs, read -> transfrag mapping, flow,
// predictions.  A "graph" is a candidate (strand lane, CBundle) pair; graph ids run bundle by bundle, '-' lane first —
// the order in which the reference builds and parses them (rlink.cpp:15660-15752, 15838-15926).
struct Graph;
struct AsmView {
  int64_t n_graphs, n_junc_rows;        // candidate graphs; rows of the junction table after rows a8 + a9 (n_juncs + n_app)
  const uint32_t* gr_first;             // [2*n_bundles+1] first graph of (bundle b, strand s) at 2b+s
  const uint32_t* jrow_off;             // [n_bundles+1]   first row of bundle b in GroupsView::oj_*
  int32_t *g_bundle, *g_k; int8_t* g_s; // [G] input bundle, CBundle index inside the lane, strand (0: '-', 1: '+')
  // strand-filtered junction lists: bundle b, strand s at 2*jrow_off[b] + s*(rows of b); jl_cnt[2b+s] entries in use
  uint32_t *js_start, *js_end, *je_end; int32_t *je_k, *reg_node, *reg_graph; uint32_t* jl_cnt;
  // sweep: bound-sized temporaries (cap_*: capacities per graph, then exclusive prefixes)
  uint32_t *cap_n, *cap_e, *cap_f, *cap_t;
  uint32_t *t_nstart, *t_nend; int32_t *t_evu, *t_evv, *t_ft1, *t_ft2; float* t_ftab; uint32_t* t_trpos; float* t_trab; uint8_t* t_trst;
  int32_t *sw_nn, *sw_ne, *sw_nf, *sw_edgeno;       // [G] what the sweep produced (nn == 0: graph not built)
  int32_t *bn_first, *bn_cnt, *bn_graph;            // [3*n_groups] bundle2graph, laid out like GroupsView::bn_start
  // exact-size graph storage (x_*: sizes per graph, then exclusive prefixes)
  uint32_t *x_node, *x_edge, *x_tf0, *x_reach;
  Graph* graphs;                                    // [G]
  uint32_t *n_start, *n_end; float* n_cov; int32_t *ch_off, *ch_cnt, *ch, *pa_off, *pa_cnt, *pa; uint8_t* n_sinkpar;
  unsigned long long* reach; int32_t *tf_n1, *tf_n2; float* tf_ab;
  int32_t* fin_stack; double* fin_ncs; uint8_t* fin_visit;   // scratch of graph_finish
  int64_t n_nodes, n_edges, n_tf0cap, n_reach;
  // a11: per-read counts -> prefixes; records; transfrag tables
  uint32_t *r_ncov, *r_ntf, *r_nkey;                // [n_reads+1]
  uint32_t* tf0_off;                                // [G+1] transfrags create_graph left, then exclusive prefix
  int32_t* g_node_off;                              // [G] = x_node (as int32, for FragWrite)
  int64_t n_covrec, n_rec, n_key, n_tf0;
  uint32_t* cov_node; float* cov_val;               // [n_covrec]
  int32_t* rec_graph; float* rec_ab; uint32_t *rec_koff, *rec_klen; unsigned long long* rec_hash; int32_t* rec_tf;   // [n_tf0 + n_rec]
  uint32_t* key;                                    // [2*n_tf0 + n_key + extra]: a11's keys, then the edge transfrags' of a12
  uint32_t* tab_off;                                // [n_bundles+1] table slots per bundle (powers of two), then exclusive prefix
  unsigned long long* slot_hash; int32_t* slot_id;  // [tab_off[n_bundles]]
  int32_t* tf_rep; float* tf_abund; int32_t* tf_graph;   // [n_tf0 + n_rec] bundle b from tfslot0(b) = tf0_off[first graph] + r_ntf[first read]
  int32_t* b_ntf;                                   // [n_bundles] distinct transfrags of the bundle
  uint32_t* slot_minrec; int32_t* slot_tf0;         // [slots] first record of the key; last transfrag of create_graph with this key (-1)
  uint32_t *rec_slot, *rec_first;                   // [n_rec] slot of the record's key; [n_rec+1] first occurrence flags -> exclusive prefix
  float *ab_sorted, *cov_sorted;                    // [n_rec], [n_covrec] values in key order (deep bundles)
  uint32_t* pool; unsigned long long* pool_ctr; unsigned long long pool_cap;   // scratch of the deep bundles' counting sorts
  uint32_t* deep_cnt; uint32_t deep_cap;            // [2] deep bundles on the transfrag / node-coverage list
  uint32_t *deep_tf_bundle, *deep_cov_bundle; unsigned long long *deep_tf_off, *deep_cov_off;   // [deep_cap]
  uint8_t* b_deep;                                  // [n_bundles] bit 0: transfrag sums on the CTA path, bit 1: node coverage
  // a12-a15: per-graph sizes -> prefixes
  uint32_t *g_nt, *g_ngap, *g_reachsum, *g_fill;    // [G] transfrags after a11, with an unflagged pair, sum of (last-first+1); gather cursor
  uint32_t *y_tf, *y_pat, *y_path, *y_ntrf, *y_istr;   // [G+1] capacities, then prefixes
  int64_t n_ytf, n_ypat, n_ypath, n_yntrf, n_yistr;
  uint32_t xkey_base;                               // first entry of `key` past a11's keys; graph g's edge transfrags at + 2*x_edge[g] + 2g
  uint32_t *f_koff, *f_klen, *f_wkoff, *f_wklen; float *f_ab, *f_wab; uint8_t* f_real; int32_t *f_usepath, *f_path_off, *f_path_cnt, *f_order;
  unsigned long long *f_pat, *f_skey; int32_t *f_path_node, *f_path_cont; float* f_path_ab; int32_t *f_ntrf_off, *f_ntrf_cnt, *f_ntrf;
  uint8_t* f_ecov;
  // flow scratch: per node arrays at x_node[g] + 2g (gno + 2 entries per graph), per edge arrays at x_edge[g]
  float *s_nodecov, *s_capl, *s_capr, *s_suml, *s_sumr, *s_flux, *s_exc; int32_t *s_path, *s_succ, *s_n2p; uint32_t *s_exs, *s_exe;
  unsigned long long *s_pathpat, *s_istr; uint8_t *s_used, *s_prevn, *s_preve;
  // predictions: pool (any order), then final (graph order)
  uint32_t* p_counters; uint32_t cap_pred, cap_exon;
  int32_t *pp_graph, *pp_seq, *pp_tlen; uint32_t *pp_start, *pp_end, *pp_exon0, *pp_nex; double* pp_cov; uint32_t *pe_start, *pe_end; float* pe_cov;
  uint32_t* g_npred;                                // [G+1] predictions per graph, then exclusive prefix
  int32_t* g_geneno;                                // [G]
  int64_t n_pred, n_exon;
  uint32_t* o_pred_off;                             // [n_bundles+1] bundle b owns predictions [o_pred_off[b], o_pred_off[b+1])
  uint32_t *o_start, *o_end, *o_exon_off; double* o_cov; int32_t *o_tlen, *o_geneno, *o_src; int8_t* o_strand;   // [n_pred] (+1 for exon_off)
  uint32_t *oe_start, *oe_end; float* oe_cov;       // [n_exon]
};

struct JuncParams {
  uint32_t junctionsupport, sserror;
  int32_t junctionthr;
  uint32_t eonly;
};

struct KernelParams {
  uint32_t junctionsupport;
  int32_t junctionthr;
  uint32_t longreads;
};

// kernel launchers (cov_kernels.cu). Each enqueues on `st` and returns the CUDA error of the launch.
cudaError_t launch_expand_count(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_scan_exclusive_add(uint32_t* data, int64_t n_plus_1, unsigned long long* state, uint32_t* ticket,
                                      cudaStream_t st);
cudaError_t launch_expand_write(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_chunk_hull(const DevBatchView& v, cudaStream_t st);      // + the three chunk-level scans
cudaError_t launch_bin_count(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_bin_columns(const DevBatchView& v, cudaStream_t st);     // + the two tile-level scans
cudaError_t launch_bin_scatter(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_junc_anchor(const DevBatchView& v, const KernelParams& p, cudaStream_t st);
cudaError_t launch_tile_plan(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_cov_pack_flag(const DevBatchView& v, uint32_t* flag, float* val, cudaStream_t st);
cudaError_t launch_cov_pack_gather(const DevBatchView& v, const uint32_t* slot, uint32_t* desc, float* raw, cudaStream_t st);
cudaError_t launch_chain_selftest(int64_t n, const uint32_t* off, const float* x, const float* s0, float* out, int plain, cudaStream_t st);
cudaError_t launch_deep_deposit(const DevBatchView& v, int num_sms, cudaStream_t st);
cudaError_t launch_cov_tiles(const DevBatchView& v, int num_sms, cudaStream_t st);
cudaError_t launch_cov_fill(const DevBatchView& v, int num_sms, cudaStream_t st);
cudaError_t launch_cov_totals(const DevBatchView& v, cudaStream_t st);
cudaError_t launch_query_cov(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* lane,
                             const uint32_t* start, const uint32_t* end, int sign, double* out, cudaStream_t st);
cudaError_t launch_junc_pass(const DevBatchView& v, const JuncParams& p, cudaStream_t st);
cudaError_t launch_find_trims(const DevBatchView& v, int64_t n, const int32_t* bundle, const int8_t* sno, const uint32_t* start,
                              const uint32_t* end, const uint32_t* cap_off, uint32_t* o_cnt, uint32_t* o_pos, float* o_ab,
                              uint8_t* o_st, cudaStream_t st);
cudaError_t launch_group_caps(const DevBatchView& v, uint32_t* rcap, unsigned long long* total, int32_t* rj, cudaStream_t st);
cudaError_t launch_groups_walk(const DevBatchView& v, const GroupsView& g, const JuncParams& p, int64_t n_cluster, int64_t n_deep,
                               cudaStream_t st, cudaStream_t aux, cudaStream_t aux3, cudaEvent_t fork, cudaEvent_t join, cudaEvent_t join3);
void groups_set_debug(int on);
int64_t groups_cluster_reads();   // bundles with at least this many reads are walked by a cluster of CTAs (row x1)
size_t groups_cluster_ctl_bytes(); size_t groups_cluster_rec_bytes();
int64_t groups_deep_reads();   // bundles with at least this many reads are walked by the wide kernel
cudaError_t launch_groups_finish(const DevBatchView& v, const GroupsView& g, cudaStream_t st);
cudaError_t launch_neutral_count(const DevBatchView& v, const GroupsView& g, uint32_t* reg_off, cudaStream_t st);
cudaError_t launch_neutral_regions(const DevBatchView& v, const GroupsView& g, const NeutralView& nv, cudaStream_t st);
cudaError_t launch_neutral_preds(const DevBatchView& v, const GroupsView& g, const NeutralView& nv, cudaStream_t st);
cudaError_t launch_neutral_compact(const DevBatchView& v, const NeutralView& nv, cudaStream_t st);
struct AsmParams;
cudaError_t launch_asm_enum(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_junc_lists(const DevBatchView& v, const GroupsView& g, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_bounds(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, cudaStream_t st);
cudaError_t launch_asm_sweep(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, cudaStream_t st);
cudaError_t launch_asm_sizes(const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_finish(const DevBatchView& v, const GroupsView& g, const AsmView& A, const AsmParams& P, cudaStream_t st);
cudaError_t launch_asm_frag_count(const DevBatchView& v, const GroupsView& g, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_frag_write(const DevBatchView& v, const GroupsView& g, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_table_caps(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_tf_commit(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_tf_ids(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_tf_sums(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_nodecov_commit(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_tf_stats(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_gather(const DevBatchView& v, const AsmView& A, cudaStream_t st);
cudaError_t launch_asm_flow(const DevBatchView& v, const AsmView& A, const AsmParams& P, uint32_t* bins, uint32_t* order, int num_sms,
                            cudaStream_t st);
size_t asm_flow_bins_words();
void asm_set_debug(int on);
cudaError_t launch_asm_geneno(const DevBatchView& v, const AsmView& A, const int32_t* b_geneno, cudaStream_t st);
cudaError_t launch_asm_pred_scatter(const AsmView& A, uint32_t npool, cudaStream_t st);
cudaError_t launch_asm_exon_scatter(const AsmView& A, uint32_t npool, cudaStream_t st);
size_t asm_graph_struct_bytes();
cudaError_t launch_asm_totals(int64_t nb, const double* numfrag, const double* fraglen, const double* covtot, double* out, cudaStream_t st);
int64_t scan_partials_needed(int64_t n);
int cov_tile_smem_bytes();
