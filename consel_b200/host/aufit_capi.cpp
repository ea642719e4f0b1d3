// C handles over the AU-fit state machines (aufit.h) so the host fit can be unit-tested
// from Python on a CPU-only box (tests/test_host_fit.py feeds it counts taken from the
// reference's golden .rep rows).  Not used by the programs themselves.
#include "aufit.h"
#include "specfun.h"
#include "refio.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

struct aufit_opts_c {
    int fitmode;
    double kappa, rrmin, rrmax;
    int dfmin;
};
static aufit::Options make(const aufit_opts_c *o) {
    aufit::Options opt;
    if (o) {
        opt.fitmode = o->fitmode;
        opt.kappa = o->kappa;
        opt.rrmin = o->rrmin;
        opt.rrmax = o->rrmax;
        opt.dfmin = o->dfmin;
    }
    return opt;
}

void *aufit_vc_new(double x_start, double threshold) { return new aufit::VcIter(x_start, threshold); }
void aufit_vc_free(void *h) { delete static_cast<aufit::VcIter *>(h); }
double aufit_vc_x(void *h) { return static_cast<aufit::VcIter *>(h)->x; }
int aufit_vc_done(void *h) { return static_cast<aufit::VcIter *>(h)->done; }
void aufit_vc_step(void *h, const double *cnts, const double *rr, const double *bb, int kk, const aufit_opts_c *o) {
    static_cast<aufit::VcIter *>(h)->step(cnts, rr, bb, kk, make(o));
}
// out: pv se pv0 se0 rss df th beta0 beta1 beta2 ret
void aufit_vc_get(void *h, double *out) {
    auto *it = static_cast<aufit::VcIter *>(h);
    double v[11] = {it->pv, it->se, it->pv0, it->se0, it->rss, it->df, it->th_out, it->beta[0], it->beta[1],
                    it->beta[2], (double)(it->done ? it->ret : 2)};
    for (int i = 0; i < 11; i++) out[i] = v[i];
}

void *aufit_ci_new(double xmin, double xmax, double alpha, double kappa) {
    return new aufit::CiIter(xmin, xmax, alpha, kappa);
}
void aufit_ci_free(void *h) { delete static_cast<aufit::CiIter *>(h); }
double aufit_ci_x(void *h) { return static_cast<aufit::CiIter *>(h)->x; }
int aufit_ci_done(void *h) { return static_cast<aufit::CiIter *>(h)->done; }
void aufit_ci_step(void *h, const double *cnts, const double *rr, const double *bb, int kk, const aufit_opts_c *o) {
    static_cast<aufit::CiIter *>(h)->step(cnts, rr, bb, kk, make(o));
}
// out: cival seval eival ret
void aufit_ci_get(void *h, double *out) {
    auto *it = static_cast<aufit::CiIter *>(h);
    out[0] = it->cival; out[1] = it->seval; out[2] = it->eival; out[3] = it->ret;
}

double aufit_pochisq(double x, int df) { return specfun::pochisq(x, df); }
double aufit_pnorm(double x) { return specfun::pnorm(x); }
double aufit_qnorm(double p) { return specfun::qnorm(p); }
}

// test seam for the matrix text reader (tests/test_host_fit.py): reads a .mt file, returns
// the shape, copies up to `cap` values, and reports the token that follows the matrix
extern "C" int refio_test_read_mat(const char *path, int *m, int *n, double *out, long long cap, char *next_word, int next_cap) {
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    int mm = 0, nn = 0;
    std::vector<double> a = refio::read_mat(fp, &mm, &nn);
    *m = mm;
    *n = nn;
    for (long long i = 0; i < cap && i < (long long)a.size(); i++) out[i] = a[i];
    std::string w = refio::read_word(fp);
    snprintf(next_word, next_cap, "%s", w.c_str());
    fclose(fp);
    return 0;
}

extern "C" {
// qnorm's shortcut against the plain 40-step bisection it replaces: number of arguments (out of n) on which the
// two differ in any bit.  Arguments: uniform, both tails, exact pnorm values, the neighbourhoods of z = 0 and
// of the series / continued-fraction switch (where the computed pnorm jumps).
long long aufit_qnorm_mismatches(long long n, unsigned long long seed) {
    unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto u01 = [&]() {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        return (double)(st >> 11) * (1.0 / 9007199254740992.0);
    };
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        const double r = u01();
        double p;
        switch ((int)(u01() * 6)) {
            case 0: p = r; break;
            case 1: p = r * 1e-3; break;
            case 2: p = 1.0 - r * 1e-3; break;
            case 3: p = specfun::pnorm(-8.0 + 16.0 * r); break;
            case 4: p = 0.5 + (r - 0.5) * 1e-6; break;
            default: p = specfun::pnorm((r < 0.5 ? -1.0 : 1.0) * (1.7320508 + (u01() - 0.5) * 1e-4)); break;
        }
        specfun::qnorm_set_fast(0);
        const double a = specfun::qnorm(p);
        specfun::qnorm_set_fast(1);
        const double b = specfun::qnorm(p);
        if (memcmp(&a, &b, sizeof(double)) != 0) bad++;
    }
    return bad;
}
int aufit_qnorm_grid_monotone(void) { return specfun::qnorm_grid_is_monotone() ? 1 : 0; }

}
