// makermt -> consel on device-resident replicate blocks, single GPU or one process per GPU
// (see pipeline.cpp).  Shared by libconselhost.so (chost_consel_run, the Python mirror) and
// the `consel` program's multi-rank mode.
#pragma once
#include "../../include/consel_gpu.h"
#include "aufit.h"

namespace pipeline {
// outputs as documented at chost_consel_run; q0_out (optional): cm x kk row summaries at `threshold`
int consel_run(cgpu_reps *reps, int K, const double *datvec, int kk, const double *rr, const int *bb, int cm,
               const int *ass_ptr, const int *ass_idx, double threshold, double varadd, const aufit::Options &opt, int do_mc,
               int do_au, int nalpha, const double *alpha, double *pv, double *se, double *cnt, double *ax7, double *ci6,
               int *ret, double *timing, cgpu_rowsum *q0_out);
}  // namespace pipeline
