// Host side of the multiscale-bootstrap (AU) p-value: the tiny per-item curve fit that
// BASELINE.json leaves on the CPU.  Re-entrant restatement of the arithmetic of
// src/consel.c:2164-2354 (WLS + MLE fit), :2744-2808 (vcalpval threshold iteration) and
// :2816-2879 (invtpval bisection), reorganised as STATE MACHINES so that all items
// advance in lockstep: one device count query per round serves every item
// (SURVEY 3.3), and the fits of a round run on all host threads.
#pragma once
#include <vector>

namespace aufit {

enum FitMode { FIT_WLS = 0, FIT_MLE = 1 };

struct Options {
    int fitmode = FIT_MLE;  // src/consel.c:228 default
    double kappa = 1.0;     // :156
    double rrmin = 0.0, rrmax = 100.0;  // :2699-2700
    int dfmin = 0;          // :2743
    double mleeps = 1e-10;  // :2236
    int mleloopmax = 30;    // :2235
    int vcloopmax = 30;     // :2739
    double vcscale = 0.5, vceps1 = 0.01, vceps2 = 0.1;  // :2740-2742
    double cieps = 0.1;     // :2815
    int ciloopmax = 30;     // :2814
};

struct Fit {
    double pv = 0, se = 0, pv0 = 0, se0 = 0, rss = 0, df = 0;
    double beta[3] = {0, 0, 0};
    int rc = 0;  // 0 ok, non-zero degenerate
};

// calcpval (src/consel.c:2677-2697): cnts/rr/bb are kk-vectors
Fit calcpval(const double *cnts, const double *rr, const double *bb, int kk, double rmin, double rmax, double kappa,
             const Options &opt);

// vcalpval as a resumable iteration.  Usage:
//   VcIter it(x_start, threshold);  while (!it.done) { cnts = counts at it.x ; it.step(cnts, rr, bb, kk, opt); }
struct VcIter {
    double x;          // threshold to query next
    bool done = false;
    int ret = 0;       // vcalpval's return value once done (0 ok, 1 degenerate, 2 no convergence)
    // accepted results (the *pvp ... of the reference)
    double pv = 0, se = 0, pv0 = 0, se0 = 0, rss = 0, df = 0, th_out = 0;
    double beta[3] = {0, 0, 0};
    VcIter(double x_start, double threshold) : x(x_start), xn(x_start), th(threshold), threshold(threshold) {}
    void step(const double *cnts, const double *rr, const double *bb, int kk, const Options &opt);

  private:
    double xn, th, threshold;
    double z = 0, ze = 0, z0 = 0, ze0 = 0;
    int idf0 = -2, j = 0;
};

// invtpval as a resumable bisection for one (alpha, kappa)
struct CiIter {
    double x = 0;      // threshold to query next (valid after construction)
    bool done = false;
    int ret = 0;
    double cival = 0, seval = 0, eival = 0;
    bool degenerate_warning = false, noconv_warning = false;
    double warn_value = 0;
    CiIter(double xmin, double xmax, double alpha, double kappa);
    void step(const double *cnts, const double *rr, const double *bb, int kk, const Options &opt);

  private:
    double x1, x2, x0 = 0, alpha, kappa, pv = 0, se = 1.0, e = 0;
    int j = 0;
};

}  // namespace aufit
