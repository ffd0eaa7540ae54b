/* oracle/stg_oracle_internal.h — TEST INFRASTRUCTURE ONLY: shared between the oracle's translation units. */
#ifndef STG_ORACLE_INTERNAL_H_
#define STG_ORACLE_INTERNAL_H_
#include "stg_oracle.h"

typedef struct {   /* CJunction, bundle.h:141-166 */
  uint32_t start, end;
  int8_t strand, guide_match, consleft, consright;
  double nreads, nreads_good, leftsupport, rightsupport, nm, mm;
} junc_t;

/* junction table of bundle b after the junction pass of build_graphs (rlink.cpp:14385-14809) */
void orc_junction_state(const stg_packed_batch* B, const orc_params* P, int64_t b, const float* cov,
                        const double* good_in, const double* left_in, const double* right_in, int extra,
                        junc_t** J_out, int32_t** ej_out, int* n_out);
/* good_junc (rlink.cpp:13967-14070) on a copy of *jd; returns the verdict, *strand_out = the strand it leaves */
int orc_good_junc(const junc_t* jd, int64_t refstart, const float* cov, int64_t len, int junctionthr, int8_t* strand_out);
/* the reference's quicksort (gclib/GVec.hh:896-920: Hoare partition, middle pivot, not stable) on an index list */
void orc_qsort_ref(int32_t* list, const junc_t* J, int L, int R, int (*cmp)(const junc_t*, const junc_t*));
#endif
