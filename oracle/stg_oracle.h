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

/* work counters for the roofline: S = mate-merged coverage intervals emitted, J = junction-support updates */
void orc_batch_stats(const stg_packed_batch* B, int64_t* L, int64_t* S, int64_t* J);

#ifdef __cplusplus
}
#endif
#endif
