#include "aufit.h"

#include <cmath>

#include "refio.h"
#include "specfun.h"

namespace aufit {
using specfun::dnorm;
using specfun::pnorm;
using specfun::qnorm;

// ---- inverse of a small matrix by Crout LU with implicit row scaling, the scheme of the
// routine the reference uses (src/opt.c:22-111): same pivot rule (last maximum wins),
// same 1e-20 guard, same substitution order, so the 2x2 covariance inverse is bit-equal.
static void small_inverse(double a[3][3], double inv[3][3], int n) {
    int index[3];
    double scale[3];
    for (int i = 0; i < n; i++) {
        double big = 0.0;
        for (int j = 0; j < n; j++)
            if (fabs(a[i][j]) > big) big = fabs(a[i][j]);
        if (big == 0.0) refio::die("luinverse: singular matrix");
        scale[i] = 1.0 / big;
    }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) {
            double sum = a[i][j];
            for (int k = 0; k < i; k++) sum -= a[i][k] * a[k][j];
            a[i][j] = sum;
        }
        double big = 0.0;
        int imax = 0;
        for (int i = j; i < n; i++) {
            double sum = a[i][j];
            for (int k = 0; k < j; k++) sum -= a[i][k] * a[k][j];
            a[i][j] = sum;
            double t = scale[i] * fabs(sum);
            if (t >= big) {
                big = t;
                imax = i;
            }
        }
        if (j != imax) {
            for (int k = 0; k < n; k++) {
                double t = a[imax][k];
                a[imax][k] = a[j][k];
                a[j][k] = t;
            }
            scale[imax] = scale[j];
        }
        index[j] = imax;
        if (a[j][j] == 0.0) a[j][j] = 1.0e-20;
        if (j != n - 1) {
            double t = 1.0 / a[j][j];
            for (int i = j + 1; i < n; i++) a[i][j] *= t;
        }
    }
    for (int col = 0; col < n; col++) {
        double w[3] = {0, 0, 0};
        w[col] = 1.0;
        int first = -1;
        for (int i = 0; i < n; i++) {
            int ip = index[i];
            double sum = w[ip];
            w[ip] = w[i];
            if (first != -1) {
                for (int j = first; j < i; j++) sum -= a[i][j] * w[j];
            } else if (sum != 0.0) {
                first = i;
            }
            w[i] = sum;
        }
        for (int i = n - 1; i >= 0; i--) {
            double sum = w[i];
            for (int j = i + 1; j < n; j++) sum -= a[i][j] * w[j];
            w[i] = sum / a[i][i];
        }
        for (int i = 0; i < n; i++) inv[i][col] = w[i];
    }
}

// weighted least squares with 2 predictors (lsfit, src/opt.c:146-195)
static void wls2(const double *X0, const double *X1, const double *Y, const double *W, int n, double beta[2],
                 double *rss, double vmat[2][2]) {
    const double *X[2] = {X0, X1};
    double cov[3][3] = {{0}}, inv[3][3] = {{0}}, xy[2];
    for (int i = 0; i < 2; i++) {
        double x = 0.0;
        for (int k = 0; k < n; k++) x += X[i][k] * Y[k] * W[k];
        xy[i] = x;
        for (int j = 0; j <= i; j++) {
            x = 0.0;
            for (int k = 0; k < n; k++) x += X[i][k] * X[j][k] * W[k];
            cov[i][j] = cov[j][i] = x;
        }
    }
    small_inverse(cov, inv, 2);
    // symmetrise (sym_mat, src/opt.c:117-130)
    double asym = fabs(inv[1][0] - inv[0][1]);
    double s = inv[1][0] + inv[0][1];
    inv[1][0] = inv[0][1] = 0.5 * s;
    if (asym > 1e-5) refio::warn("lsfit: covmat singularity %g", asym);
    if (asym > 1e-3) refio::warn("lsfit: COVARIANCE MATRIX IS SINGULAR");
    for (int i = 0; i < 2; i++) {
        double x = 0.0;
        for (int j = 0; j < 2; j++) x += inv[i][j] * xy[j];
        beta[i] = x;
    }
    double x = 0.0;
    for (int k = 0; k < n; k++) {
        double y = 0.0;
        for (int i = 0; i < 2; i++) y += X[i][k] * beta[i];
        x += W[k] * (Y[k] - y) * (Y[k] - y);
    }
    *rss = x;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) vmat[i][j] = inv[i][j];
}

// wlscalcpval (src/consel.c:2164-2225)
static Fit fit_wls(const double *cnts, const double *rr, const double *bb, int kk, double rmin, double rmax,
                   double kappa, double vmat[2][2]) {
    Fit f;
    // per-thread scratch: this function runs ~10^4 times per consel call on every pool thread
    static thread_local std::vector<double> scratch;
    if ((int)scratch.size() < 5 * kk) scratch.resize(5 * (size_t)kk);
    double *X0 = scratch.data(), *X1 = X0 + kk, *npv = X1 + kk, *zval = npv + kk, *wt = zval + kk;
    int m = 0;
    for (int i = 0; i < kk; i++) {
        double s = sqrt(1.0 / rr[i]);
        X0[i] = 1.0 / s;
        X1[i] = s;
        npv[i] = cnts[i] / bb[i];
        if (rr[i] < rmin || rr[i] > rmax) {
            zval[i] = 0.0;
            wt[i] = 0.0;
        } else if (npv[i] >= 1.0 || npv[i] <= 0.0) {
            zval[i] = 0.0;
            wt[i] = 0.0;
        } else {
            zval[i] = -qnorm(npv[i]);
            double x = dnorm(zval[i]);
            wt[i] = bb[i] * x * x / ((1.0 - npv[i]) * npv[i]);
            m++;
        }
    }
    f.df = (double)(m - 2);
    if (m < 2) {
        double x = 0.0;
        for (int i = 0; i < kk; i++) x += zval[i];
        f.pv = x >= 0.0 ? 0.0 : 1.0;
        f.se = 0.0;
        f.rss = 0.0;
        f.beta[0] = f.beta[1] = 0.0;
        f.pv0 = f.pv;
        f.se0 = f.se;
        f.rc = 1;
        refio::dprint(1, "\n# error in wls");
    } else {
        double beta[2];
        wls2(X0, X1, zval, wt, kk, beta, &f.rss, vmat);
        f.beta[0] = beta[0];
        f.beta[1] = beta[1];
        f.pv = pnorm(-(beta[0] - kappa * beta[1]));
        double x = dnorm(beta[0] - kappa * beta[1]);
        f.se = sqrt(x * x * (vmat[0][0] + kappa * kappa * vmat[1][1] - kappa * vmat[0][1] - kappa * vmat[1][0]));
        f.pv0 = pnorm(-(beta[0] + beta[1]));
        x = dnorm(beta[0] + beta[1]);
        f.se0 = sqrt(x * x * (vmat[0][0] + vmat[1][1] + vmat[0][1] + vmat[1][0]));
        f.rc = 0;
    }
    f.beta[2] = 0.0;
    return f;
}

// -log(1-x) with a series near 0 (src/consel.c:2222-2233)
static double log3(double x) {
    if (fabs(x) > 1.0e-3) return -log(1.0 - x);
    double z1 = x, z2 = z1 * x, z3 = z2 * x, z4 = z3 * x, z5 = z4 * x;
    return ((((z5 / 5.0) + z4 / 4.0) + z3 / 3.0) + z2 / 2.0) + z1;
}

// mlecoef (src/consel.c:2237-2330): Newton-Raphson on the binomial likelihood
static int fit_mle_coef(const double *cnts, const double *rr, const double *bb, int kk, double coef0[2],
                        double vmat[2][2], double *lrt, double *df, double rmin, double rmax, const Options &opt) {
    static thread_local std::vector<double> scratch;
    if ((int)scratch.size() < 9 * kk) scratch.resize(9 * (size_t)kk);
    double *s = scratch.data(), *r = s + kk, *c = r + kk, *b = c + kk, *z = b + kk, *p = z + kk, *d = p + kk, *g = d + kk,
           *h = g + kk;
    int m = 0;
    for (int i = 0; i < kk; i++)
        if (rr[i] >= rmin && rr[i] <= rmax) {
            r[m] = rr[i];
            s[m] = sqrt(rr[i]);
            c[m] = cnts[i];
            b[m] = bb[i];
            m++;
        }
    if (m < 2) return 1;
    double coef[2] = {coef0[0], coef0[1]}, update[2];
    double v11 = 0, v12 = 0, v22 = 0;
    refio::dprint(3, "\n### mlecoef: deg=%d, coef0=(%g,%g)", m - 2, coef[0], coef[1]);
    int loop;
    for (loop = 0; loop < opt.mleloopmax; loop++) {
        double d1f = 0, d2f = 0, d11f = 0, d12f = 0, d22f = 0;
        for (int i = 0; i < m; i++) {
            z[i] = coef[0] * s[i] + coef[1] / s[i];
            p[i] = pnorm(-z[i]);
            d[i] = dnorm(z[i]);
            if (p[i] > 0.0 && p[i] < 1.0) {
                g[i] = d[i] * (d[i] * (-c[i] + 2.0 * c[i] * p[i] - b[i] * p[i] * p[i]) /
                                   (p[i] * p[i] * (1.0 - p[i]) * (1.0 - p[i])) +
                               z[i] * (c[i] - b[i] * p[i]) / (p[i] * (1.0 - p[i])));
                h[i] = d[i] * (c[i] - b[i] * p[i]) / (p[i] * (1.0 - p[i]));
            } else {
                g[i] = h[i] = 0.0;
            }
            d1f += -h[i] * s[i];
            d2f += -h[i] / s[i];
            d11f += g[i] * r[i];
            d12f += g[i];
            d22f += g[i] / r[i];
            refio::dprint(3, "\n### mlecoef: %d %g %g %g %g %g", i, z[i], p[i], d[i], g[i], h[i]);
        }
        double a = d11f * d22f - d12f * d12f;
        if (a == 0.0) {
            refio::dprint(1, "\n# singular matrix in mle");
            return 2;
        }
        v11 = -d22f / a;
        v12 = d12f / a;
        v22 = -d11f / a;
        update[0] = v11 * d1f + v12 * d2f;
        update[1] = v12 * d1f + v22 * d2f;
        coef[0] += update[0];
        coef[1] += update[1];
        double e = -d11f * update[0] * update[0] - 2.0 * d12f * update[0] * update[1] - d22f * update[1] * update[1];
        refio::dprint(2, "\n## mlecoef: %d %g %g %g %g %g", loop, coef[0], coef[1], update[0], update[1], e);
        if (e < opt.mleeps) break;
    }
    *lrt = 0.0;
    *df = 0.0;
    for (int i = 0; i < m; i++) {
        if (p[i] > 0.0 && p[i] < 1.0) {
            *df += 1.0;
            double a;
            if (c[i] > 0.0) a = c[i] * log(c[i] / b[i] / p[i]); else a = 0.0;
            if (c[i] < b[i]) a += (b[i] - c[i]) * (log3(p[i]) - log3(c[i] / b[i]));
            *lrt += a;
        }
    }
    *lrt *= 2.0;
    *df -= 2.0;
    coef0[0] = coef[0];
    coef0[1] = coef[1];
    vmat[0][0] = v11;
    vmat[0][1] = vmat[1][0] = v12;
    vmat[1][1] = v22;
    return (loop == opt.mleloopmax || *df < -0.01) ? 1 : 0;
}

Fit calcpval(const double *cnts, const double *rr, const double *bb, int kk, double rmin, double rmax, double kappa,
             const Options &opt) {
    double vmat[2][2] = {{0, 0}, {0, 0}};
    Fit f = fit_wls(cnts, rr, bb, kk, rmin, rmax, kappa, vmat);
    if (opt.fitmode == FIT_WLS || f.rc) return f;
    // mlecalcpval (src/consel.c:2333-2365): start from the WLS coefficients
    double coef[2] = {f.beta[0], f.beta[1]};
    double vm[2][2], lrt = f.rss, df = f.df;
    int rc = fit_mle_coef(cnts, rr, bb, kk, coef, vm, &lrt, &df, rmin, rmax, opt);
    if (rc) {
        // "error in mle, use wls instead" (src/consel.c:2352): the WLS p-values and
        // coefficients stand; rss/df are whatever mlecoef left behind its pointers (the
        // WLS values if it stopped on a singular Hessian, its own lrt/df if it ran out of
        // iterations or fitted with negative df).
        f.rss = lrt;
        f.df = df;
        refio::dprint(0, "\n# error in mle, use wls instead");
        f.rc = 0;
        return f;
    }
    Fit g;
    g.rss = lrt;
    g.df = df;
    g.pv = pnorm(-(coef[0] - kappa * coef[1]));
    double x = dnorm(coef[0] - kappa * coef[1]);
    g.se = sqrt(x * x * (vm[0][0] + kappa * kappa * vm[1][1] - kappa * vm[0][1] - kappa * vm[1][0]));
    g.pv0 = pnorm(-(coef[0] + coef[1]));
    x = dnorm(coef[0] + coef[1]);
    g.se0 = sqrt(x * x * (vm[0][0] + vm[1][1] + vm[0][1] + vm[1][0]));
    g.beta[0] = coef[0];
    g.beta[1] = coef[1];
    g.beta[2] = 0.0;
    g.rc = 0;
    return g;
}

// ---------------------------------------------------------------- vcalpval (src/consel.c:2744-2808)
void VcIter::step(const double *cnts, const double *rr, const double *bb, int kk, const Options &opt) {
    if (done) return;
    const double xq = x;  // the threshold the counts were taken at
    Fit f = calcpval(cnts, rr, bb, kk, opt.rrmin, opt.rrmax, opt.kappa, opt);
    if (!f.rc) {
        z = -f.pv;
        ze = f.se;
    }
    const int idf = (int)f.df;
    refio::dprint(2, "\n## vcalpval: %3d %6.3f %6.3f %9.6f %9.6f %3.0f %6.3f %6.3f %6.3f %6.3f %g", j, th, xq, f.pv,
                  f.se, f.df, f.beta[0], f.beta[1], z, ze, f.rss);
    auto next_round = [&]() {
        j++;
        if (j >= opt.vcloopmax) {
            refio::dprint(1, "\n# no-convergence");
            done = true;
            ret = 2;
        }
        x = xn;
    };
    if (idf < opt.dfmin && idf0 < opt.dfmin) {
        done = true;
        ret = 1;
        return;
    }
    if ((idf < opt.dfmin) || (idf0 >= opt.dfmin && (z - z0) * (xq - th_out) > 0.0 && fabs(z - z0) > opt.vceps2 * ze0)) {
        refio::dprint(1, "\n# non-monotone");
        th = xq;
        xn = opt.vcscale * xq + (1.0 - opt.vcscale) * th_out;
        next_round();
        return;
    }
    if (idf0 >= opt.dfmin && (fabs(z - z0) < opt.vceps1 * ze0)) {
        if (fabs(th - threshold) < 1e-10) xn = th; else th = xq;
    } else {
        xn = opt.vcscale * th + (1.0 - opt.vcscale) * xq;
    }
    pv = f.pv; se = f.se; rss = f.rss; df = f.df; th_out = xq;
    pv0 = f.pv0; se0 = f.se0;
    beta[0] = f.beta[0]; beta[1] = f.beta[1]; beta[2] = f.beta[2];
    z0 = z; ze0 = ze; idf0 = idf;
    if (fabs(xq - th) < 1e-10) {
        done = true;
        ret = 0;
        return;
    }
    next_round();
}

// ---------------------------------------------------------------- invtpval (src/consel.c:2816-2879)
CiIter::CiIter(double xmin, double xmax, double a, double kap) : x1(xmin), x2(xmax), alpha(a), kappa(kap) {
    if (alpha < 0.0) alpha = 0.0; else if (alpha > 1.0) alpha = 1.0;
    x0 = 0.0;
    x = 0.5 * (x1 + x2);
}

void CiIter::step(const double *cnts, const double *rr, const double *bb, int kk, const Options &opt) {
    if (done) return;
    const double se0 = se, pv0 = pv;
    Fit f = calcpval(cnts, rr, bb, kk, opt.rrmin, opt.rrmax, kappa, opt);
    pv = f.pv;
    se = f.se;
    e = pv - alpha;
    (void)se0;
    if ((pv0 - alpha) * (pv - pv0) > 0.0) refio::dprint(1, "non-monotone in ci (%g %g)->(%g %g)", x0, pv0, x, pv);
    if (f.rc) {
        degenerate_warning = true;
        warn_value = f.df;
        cival = x0;
        seval = 0.0;
        eival = 0.0;
        done = true;
        ret = 1;
        return;
    }
    auto finish = [&](double pv_prev) {
        cival = x;
        seval = se * (x - x0) / (pv - pv_prev);
        eival = e / se;
    };
    if (j > 0 && fabs(e) < se * opt.cieps) {
        finish(pv0);
        done = true;
        ret = 0;
        return;
    }
    if (e > 0.0) x2 = x; else x1 = x;
    j++;
    if (j >= opt.ciloopmax) {
        finish(pv0);
        noconv_warning = true;
        warn_value = e / se;
        done = true;
        ret = 1;
        return;
    }
    x0 = x;
    x = 0.5 * (x1 + x2);
}

}  // namespace aufit
