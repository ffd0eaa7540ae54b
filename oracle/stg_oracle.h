/* oracle/stg_oracle.h — TEST INFRASTRUCTURE ONLY.
 * Plain-C, single-threaded-per-bundle restatement of the reference's coverage / junction-support stage
 * (count_good_junctions, rlink.cpp:16656-17136, short-read path) over the packed-batch schema of
 * include/stgpu.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * Pinned against the reference itself: tests/test_oracle_vs_reference.py compares it bit-for-bit with the
 * stage dumps of oracle/_ref/ref_tool and oracle/_ref/ref_hot, and with tests/golden/ fixtures. */
#ifndef STG_ORACLE_H_
#define STG_ORACLE_H_
#include "../include/stgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  uint32_t junctionsupport; /* 10 */
  int32_t junctionthr;      /* 1 */
  uint8_t longreads;        /* 0 */
} orc_params;

/* bpcov length of bundle b: end-start+3 (rlink.cpp:16875) */
int64_t orc_bundle_len(const stg_packed_batch* B, int64_t b);

/* One bundle.  cov: [3*len] lane-major, overwritten.  good/left/right: this bundle's junction rows (n = junc
 * count of the bundle), ACCUMULATED into (the reference accumulates into CJunction fields that start at 0). */
void orc_bundle(const stg_packed_batch* B, const orc_params* P, int64_t b, float* cov, double* good, double* left,
                double* right);

/* Whole batch, OpenMP over bundles (each bundle is sequential, exactly in reference order).
 * cov_off[b] = float offset of bundle b in cov (lane stride = its len).  cov may be NULL (junction support only). */
void orc_batch_run(const stg_packed_batch* B, const orc_params* P, const int64_t* cov_off, float* cov, double* good,
                   double* left, double* right, int nthreads);

/* get_cov (rlink.cpp:1830) / get_cov_sign (rlink.cpp:1842) on one bundle's lane-major bpcov */
double orc_get_cov(const float* cov, int64_t len, int s, uint32_t start, uint32_t end);
double orc_get_cov_sign(const float* cov, int64_t len, int s, uint32_t start, uint32_t end);

/* SURVEY.md §8 row a7 — the junction pass that opens build_graphs (rlink.cpp:14385-14809) plus good_junc
 * (rlink.cpp:13967) on every stranded junction, for one bundle (oracle/stg_oracle_junc.c).
 * cov: the bundle's finished bpcov [3*len]; good/left/right_in: junction counters after the coverage stage (this
 * bundle's rows).  Outputs, one per junction row of the bundle: CJunction fields after the pass; o_ej[k] = row of
 * ejunction[k]; o_good = good_junc verdict on a copy (-1 where strand == 0), o_good_strand = the copy's strand. */
void orc_junction_pass(const stg_packed_batch* B, const orc_params* P, int64_t b, const float* cov,
                       const double* good_in, const double* left_in, const double* right_in, int8_t* o_strand,
                       double* o_nreads, double* o_nreads_good, double* o_left, double* o_right, double* o_mm,
                       int32_t* o_ej, int8_t* o_good, int8_t* o_good_strand);

/* SURVEY.md §8 rows a8 + a9 (first half) — the read loop of build_graphs (rlink.cpp:14845-15127: per-read junction
 * repair, un-pairing, add_read_to_group, fragment counting) and the merging of neighbouring groups (:15134-15164),
 * for one bundle (oracle/stg_oracle_groups.c).  Inputs as for orc_junction_pass.  Everything in the result is
 * malloc'ed by the callee; release with orc_groups_free. */
typedef struct orc_groups_out {
  int32_t n_reads, n_segs, n_pairs, n_juncs, n_groups, n_colors, color;
  int32_t startgroup[3];          /* first group of every strand lane (-1: none) */
  double num_fragments, frag_len; /* BundleData::num_fragments / frag_len */
  /* reads after the loop: strand, pair links (CSR of the input), segments (CSR of the input), junction rows */
  int8_t* r_strand; int32_t* pair_idx; uint32_t *seg_start, *seg_end; int32_t* rjunc;
  /* junction table after the loop (appended rows included, re-sorted as the reference does) */
  uint32_t *j_start, *j_end; int8_t* j_strand; double *j_nreads, *j_nreads_good, *j_left, *j_right, *j_nm, *j_mm;
  /* groups: slot g == CGroup::grid; dead slots (merged away) have g_alive == 0 */
  uint32_t *g_start, *g_end; int32_t *g_color, *g_next; float *g_cov_sum, *g_multi; int8_t* g_alive;
  int32_t* merge;                 /* [n_groups] */
  int32_t* eqcol;                 /* [n_colors] */
  uint32_t* rg_off; int32_t* rg;  /* readgroup[n] = rg[rg_off[n] .. rg_off[n+1]) */
  /* second half of a9 — colour pass (rlink.cpp:15196-15299) and bundle creation (:15340-15427) */
  int32_t* g_color2;              /* [n_groups] CGroup::color after both passes */
  float* g_neg_prop;              /* [n_groups] */
  int32_t *eqcol2, *eqnegcol, *eqposcol;   /* [n_colors] equalcolor after the passes, eqnegcol, eqposcol */
  int32_t bun_off[4], bn_off[4];  /* bundles / bundlenodes of strand lane s: [off[s], off[s+1]) */
  int32_t *bun_len, *bun_start, *bun_last; float *bun_cov, *bun_multi;        /* CBundle, bundle.h:260-268 */
  uint32_t *bn_start, *bn_end; float* bn_cov; int32_t* bn_next;                /* CBundlenode, rlink.h:56-62 */
  int32_t* group2bundle;          /* [3 * n_groups] */
} orc_groups_out;
void orc_groups_bundle(const stg_packed_batch* B, const orc_params* P, int64_t b, const float* cov,
                       const double* good_in, const double* left_in, const double* right_in, uint32_t bundledist,
                       orc_groups_out* out);
void orc_groups_free(orc_groups_out* out);

/* find_all_trims (rlink.cpp:2118-2419, part of SURVEY.md §8 row a10) on one region of one bundle (no point features):
 * cov = the bundle's finished bpcov [3*blen], sno in {0, 2} as at the call sites (rlink.cpp:2813, 15559), start / end
 * genomic.  Writes the reference's `trimpoint` vector (rejected points keep pos = 0) and returns its length, at most
 * orc_trims_capacity(start, end). */
int orc_trims_capacity(uint32_t start, uint32_t end);
int orc_find_all_trims(const float* cov, int64_t blen, int64_t refstart, int sno, uint32_t start, uint32_t end,
                       uint32_t* o_pos, float* o_abund, uint8_t* o_start);

/* SURVEY.md §8 row a19 (de novo part) — single-exon predictions of the neutral bundles (rlink.cpp:15500-15634), for one
 * bundle: G = its groups / bundles (orc_groups_bundle).  Writes up to `cap` predictions (genomic start / end, mean
 * coverage, length, geneno as stored: the reference's counter minus one) and returns how many there are. */
int orc_neutral_preds(const float* cov, int64_t blen, int64_t refstart, const orc_groups_out* G, int mintranscriptlen,
                      float mcov, int cap, uint32_t* o_start, uint32_t* o_end, double* o_cov, int32_t* o_tlen,
                      int32_t* o_geneno);

/* work counters for the roofline: S = mate-merged coverage intervals emitted, J = junction-support updates */
void orc_batch_stats(const stg_packed_batch* B, int64_t* L, int64_t* S, int64_t* J);

#ifdef __cplusplus
}
#endif
#endif
