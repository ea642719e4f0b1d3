// Lockstep driver of the AU test (do_bptest + do_bootrep of src/consel.c:1286-1303,
// 1441-1538): every item's vcalpval / invtpval state machine (aufit.h) advances one step
// per round, and each round costs ONE count query over all items.  The query is a
// callback so the same driver serves the command line (cgpu_count_query on one GPU), the
// Python pipeline and the sharded multi-GPU run (local query + all-reduce).
#pragma once
#include <vector>

#include "../../include/consel_gpu.h"
#include "aufit.h"

namespace audriver {

// same shape as cgpu_count_query(const cgpu_reps*, int nq, const double *thr, cgpu_rowsum *out)
typedef int (*QueryFn)(void *ctx, int nq, const double *thr, cgpu_rowsum *out);
// same shape as cgpu_count_query_rows: rows [row0, row0+nrows) only (thr nrows x nq, out nrows x S x nq)
typedef int (*QueryRowsFn)(void *ctx, int row0, int nrows, int nq, const double *thr, cgpu_rowsum *out);

// Sharded runs (one process per GPU): `in` describes THIS rank's items only and every query is a
// collective that all ranks enter together -- same shape as cgpu_count_query_dist(reps, 0, ...):
// thr = cm x nq thresholds of the rank's items, out = cm x kk x nq summaries over ALL replicates,
// `active` = this rank still iterates, *any_active = some rank does.  nq is the same on every
// rank in every round (1 in the vcalpval phase, min(2*nalpha, 16) per chunk in the invtpval phase),
// and a rank whose items have converged keeps entering rounds until nobody is active.
typedef int (*DistQueryFn)(void *ctx, int nq, const double *thr, cgpu_rowsum *out, int active, int *any_active);

struct Input {
    int cm = 0, kk = 0;
    const double *rr = nullptr;   // kk scales
    const int *bb = nullptr;      // kk replicate counts (whole rows, all ranks)
    int i0 = 0;                   // scale nearest r = 1
    double threshold = 0.0;
    const double *median = nullptr;      // cm: stat[i0][i][bb[i0]/2]
    const cgpu_rowsum *r0 = nullptr;     // cm x kk summaries at `threshold` (gives min/max of the i0 rows)
    int nalpha = 0;
    const double *alpha = nullptr;
    bool confidence = true;
};

struct Output {
    std::vector<double> pv, se, pv0, se0, rss, df, th, pf, beta;  // cm each (beta: 3*cm)
    std::vector<int> ret;                                          // vcalpval return code per item
    std::vector<double> ci, cs, ei, ci0, cs0, ei0;                 // cm x nalpha each
    std::vector<int> ci_ret, ci_warn;                              // cm x nalpha x 2: return code, 1=degenerate 2=no convergence
    std::vector<double> ci_warnval;
    double t_fit_sec = 0;
};

// returns 0, or the first non-zero code of the query callback
// With `query_rows` the items are split into two halves that alternate between the device
// (count query) and the host (fits); results do not depend on the schedule.
int run(const Input &in, const aufit::Options &opt, QueryFn query, void *ctx, Output *out, QueryRowsFn query_rows = nullptr,
        DistQueryFn dist_query = nullptr);

}  // namespace audriver
