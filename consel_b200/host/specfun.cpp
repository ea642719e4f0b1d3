#include "specfun.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "refio.h"

namespace specfun {

double dnorm(double x) { return 0.398942280401432677939946059934 * exp(-0.5 * x * x); }

// ln Gamma(x), Lanczos g=5 n=6 (NR 6.1)
static double lngamma(double xx) {
    static const double cof[6] = {76.18009172947146,   -86.50532032941677,   24.01409824083091,
                                  -1.231739572450155,  0.1208650973866179e-2, -0.5395239384953e-5};
    double x = xx, y = xx;
    double tmp = x + 5.5;
    tmp -= (x + 0.5) * log(tmp);
    double ser = 1.000000000190015;
    for (int j = 0; j <= 5; j++) ser += cof[j] / ++y;
    return -tmp + log(2.5066282746310005 * ser / x);
}

// every call of the incomplete gamma function on this path has a = 0.5 (erfc): ln Gamma(0.5) is
// evaluated once by the same routine (same bits as evaluating it per call, a third of pnorm's cost)
static double lngamma_cached(double a) {
    static const double half = lngamma(0.5);
    return a == 0.5 ? half : lngamma(a);
}

static const int kItMax = 1000;
static const double kEps = 3.0e-7, kFpMin = 1.0e-30;

// P(a,x) by its series (NR 6.2 gser)
static double gamma_series(double a, double x, int *iters = nullptr) {
    const double gln = lngamma_cached(a);
    if (x <= 0.0) {
        if (x < 0.0) refio::die("x less than 0 in routine gser");
        if (iters) *iters = 0;
        return 0.0;
    }
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 1; n <= kItMax; n++) {
        ++ap;
        del *= x / ap;
        sum += del;
        if (fabs(del) < fabs(sum) * kEps) {
            if (iters) *iters = n;
            return sum * exp(-x + a * log(x) - gln);
        }
    }
    refio::die("a too large, ITMAX too small in routine gser");
}

// Q(a,x) by the modified Lentz continued fraction (NR 6.2 gcf)
static double gamma_cfrac(double a, double x, int *iters = nullptr) {
    const double gln = lngamma_cached(a);
    double b = x + 1.0 - a, c = 1.0 / kFpMin, d = 1.0 / b, h = d;
    int i;
    for (i = 1; i <= kItMax; i++) {
        double an = -i * (i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < kFpMin) d = kFpMin;
        c = b + an / c;
        if (fabs(c) < kFpMin) c = kFpMin;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < kEps) break;
    }
    if (i > kItMax) refio::die("a too large, ITMAX too small in gcf");
    if (iters) *iters = i;
    return exp(-x + a * log(x) - gln) * h;
}

static double gammp(double a, double x) {
    if (x < 0.0 || a <= 0.0) refio::die("Invalid arguments in routine gammp");
    return x < a + 1.0 ? gamma_series(a, x) : 1.0 - gamma_cfrac(a, x);
}
static double gammq(double a, double x) {
    if (x < 0.0 || a <= 0.0) refio::die("Invalid arguments in routine gammq");
    return x < a + 1.0 ? 1.0 - gamma_series(a, x) : gamma_cfrac(a, x);
}
static double erfc_nr(double x) { return x < 0.0 ? 1.0 + gammp(0.5, x * x) : gammq(0.5, x * x); }

double pnorm(double x) { return 0.5 * erfc_nr(-x * 0.707106781186547524400844362105); }

// pnorm together with a tag that names the path the evaluation took: which side of erfc, series or
// continued fraction, and after how many terms the 3e-7 stopping rule fired.  Between two arguments
// with the same tag the computed function is one fixed smooth expression (identical operation
// sequence), hence monotone at any resolution far above one ulp; the tag changes exactly where the
// computed function jumps (by up to ~3e-7 relative).
static double pnorm_tagged(double z, uint32_t *tag) {
    const double x = -z * 0.707106781186547524400844362105;
    const double x2 = x * x;
    int it = 0;
    if (x < 0.0) {  // erfc = 1 + P(1/2, x^2)
        if (x2 < 0.0) refio::die("Invalid arguments in routine gammp");
        double g;
        if (x2 < 1.5) {
            g = gamma_series(0.5, x2, &it);
            *tag = (0u << 16) | (uint32_t)it;
        } else {
            g = 1.0 - gamma_cfrac(0.5, x2, &it);
            *tag = (1u << 16) | (uint32_t)it;
        }
        return 0.5 * (1.0 + g);
    }
    double g;  // erfc = Q(1/2, x^2)
    if (x2 < 1.5) {
        g = 1.0 - gamma_series(0.5, x2, &it);
        *tag = (2u << 16) | (uint32_t)it;
    } else {
        g = gamma_cfrac(0.5, x2, &it);
        *tag = (3u << 16) | (uint32_t)it;
    }
    return 0.5 * g;
}

// qnorm is a plain bisection on pnorm over [-37, 37] to 1e-10 (src/rand.c:465-493): 40 halvings, the
// answer is the midpoint of the cell [g_j, g_j + h] of the grid h = 74 * 2^-40 with
// pnorm(g_j) <= p < pnorm(g_j + h) (all grid arithmetic is exact in fp64).  To keep the printed
// digits of the reference, that cell must be found for the COMPUTED pnorm, whose 3e-7 stopping
// rule makes it jump wherever the number of series / continued-fraction terms changes.
//
// Levels 1-20.  On the coarse grid G_i = -37 + i * 74 * 2^-20 the computed pnorm is non-decreasing
// (its increment per cell, >= 5e-5 relative, dwarfs any jump; tests/test_host_fit.py checks all
// 2^20 + 1 values), so the state of the bisection after 20 halvings is THE cell with
// pnorm(G_i) <= p < pnorm(G_{i+1}), whichever way it is found: a closed-form estimate of the quantile
// picks the cell, the memoised end values (lazy table, 1M entries; benign race: every thread writes
// the same values) confirm or correct it.
//
// Levels 21-40 all lie inside that cell of width 7e-5.  When both ends carry the SAME tag the
// computed pnorm is one smooth monotone expression on the whole cell, the final cell is the unique
// crossing, and it is found by interpolation + secant steps + verification (3-4 evaluations instead
// of 20; every evaluation must show the same tag).  Otherwise (a tag boundary inside: the function
// jumps there) the plain bisection runs over the remaining 20 levels.
// The result is the bisection's, bit for bit (tests/test_host_fit.py compares the two on millions
// of arguments: tails, exact grid values, the neighbourhoods of z = 0 and of the series /
// continued-fraction switch).
struct QnormGrid {
    std::atomic<uint64_t> *pv;
    std::atomic<uint32_t> *tag;
};
static constexpr int kGridBits = 20;
static constexpr long long kGridCells = 1ll << kGridBits;
static QnormGrid qnorm_grid() {
    static QnormGrid t = [] {
        QnormGrid m;
        m.pv = new std::atomic<uint64_t>[kGridCells + 1];
        m.tag = new std::atomic<uint32_t>[kGridCells + 1];
        for (long long i = 0; i <= kGridCells; i++) {
            m.pv[i].store(~0ull, std::memory_order_relaxed);  // a NaN pattern = empty
            m.tag[i].store(0u, std::memory_order_relaxed);
        }
        return m;
    }();
    return t;
}
static inline double grid_point(long long i) { return -37.0 + (double)i * (74.0 / 1048576.0); }  // exact
static inline double grid_value(const QnormGrid &g, long long i, uint32_t *tag) {
    uint64_t bits = g.pv[i].load(std::memory_order_acquire);
    double pv;
    if (bits == ~0ull) {
        pv = pnorm_tagged(grid_point(i), tag);
        memcpy(&bits, &pv, 8);
        g.tag[i].store(*tag, std::memory_order_relaxed);
        g.pv[i].store(bits, std::memory_order_release);
    } else {
        memcpy(&pv, &bits, 8);
        *tag = g.tag[i].load(std::memory_order_relaxed);
    }
    return pv;
}

// rational approximation of the normal quantile (P. J. Acklam), relative error ~1e-9: only used to
// pick the coarse cell to look at first
static double quantile_estimate(double p) {
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01,  -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    if (p < 0.02425) {
        const double q = sqrt(-2.0 * log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    if (p > 1.0 - 0.02425) {
        const double q = sqrt(-2.0 * log(1.0 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    const double q = p - 0.5, r = q * q;
    return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
           (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
}

static std::atomic<int> g_qnorm_fast{1};
void qnorm_set_fast(int on) { g_qnorm_fast.store(on); }

// the reference's loop, as published
static double qnorm_bisect(double p, double lo, double hi) {
    double z = (hi + lo) * 0.5;
    while (hi - lo > 1e-10) {
        if (pnorm(z) > p) hi = z; else lo = z;
        z = (hi + lo) * 0.5;
    }
    return z;
}

bool qnorm_grid_is_monotone() {  // tests: the property the coarse-cell shortcut rests on
    const QnormGrid g = qnorm_grid();
    uint32_t tag;
    double prev = grid_value(g, 0, &tag);
    for (long long i = 1; i <= kGridCells; i++) {
        const double v = grid_value(g, i, &tag);
        if (v < prev) return false;
        prev = v;
    }
    return true;
}

double qnorm(double p) {
    if (p <= 0.0 || p >= 1.0) return 0.0;
    if (!g_qnorm_fast.load(std::memory_order_relaxed)) return qnorm_bisect(p, -37.0, 37.0);
    const QnormGrid g = qnorm_grid();
    // ---- levels 1-20: the coarse cell [G_i, G_{i+1}] with pnorm(G_i) <= p < pnorm(G_{i+1})
    double zq = quantile_estimate(p);
    if (!(zq > -37.0)) zq = -37.0;
    if (zq > 37.0) zq = 37.0;
    long long i = (long long)floor((zq + 37.0) * (1048576.0 / 74.0));
    if (i < 0) i = 0;
    if (i > kGridCells - 1) i = kGridCells - 1;
    uint32_t tlo, thi;
    double plo = grid_value(g, i, &tlo), phi = grid_value(g, i + 1, &thi);
    // (the bisection never evaluates the ends -37 and 37: it takes pnorm(-37) <= p < pnorm(37) for granted)
    while (i > 0 && plo > p) {
        i--;
        phi = plo;
        thi = tlo;
        plo = grid_value(g, i, &tlo);
    }
    while (i < kGridCells - 1 && phi <= p) {
        i++;
        plo = phi;
        tlo = thi;
        phi = grid_value(g, i + 1, &thi);
    }
    const double A = grid_point(i), B = grid_point(i + 1);
    // ---- levels 21-40 inside the cell
    if (tlo == thi && i > 0 && i < kGridCells - 1) {
        const double h = 74.0 / 1099511627776.0;  // 74 * 2^-40: every grid point A + j*h is exact
        const double slope = (phi - plo) / 1048576.0;  // pnorm per fine cell, good to ~1e-4 relative over 7e-5
        bool ok = slope > 0.0;
        // first probe: the chord estimate t of the position inside the cell, corrected for the curvature of
        // the normal CDF over the cell (P'' = -z P': w = t + b t (1 - t) with b = -z W / 2; the chord alone is
        // off by up to ~10 |z| fine cells, the corrected estimate by less than one).  Only the number of
        // probes depends on it: the answer is pinned by the invariant below.
        long long j = 0;
        if (ok) {
            const double t = (p - plo) / (phi - plo);
            const double b = -0.5 * A * (74.0 / 1048576.0);
            j = (long long)((t + b * t * (1.0 - t)) * 1048576.0);
        }
        // invariant: P(A + jl*h) <= p < P(A + jh*h), jl < jh
        long long jl = 0, jh = 1048576;
        for (int step = 0; ok && step < 12 && jh - jl > 1; step++) {
            if (j <= jl) j = jl + 1;
            if (j >= jh) j = jh - 1;
            uint32_t t2;
            const double pj = pnorm_tagged(A + (double)j * h, &t2);
            if (t2 != tlo) { ok = false; break; }
            long long dj = (long long)floor((p - pj) / slope);
            if (pj > p) {
                jh = j;
                if (dj > -1) dj = -1;
            } else {
                jl = j;
                if (dj < 1) dj = 1;
            }
            j += dj;
        }
        if (ok && jh - jl == 1) return (A + (double)jl * h + (A + (double)jh * h)) * 0.5;
        // not settled (a tag boundary inside the cell, or a flat stretch): the plain bisection decides
    }
    return qnorm_bisect(p, A, B);
}

double poz(double z) {
    double x;
    if (z == 0.0) {
        x = 0.0;
    } else {
        double y = 0.5 * fabs(z);
        if (y >= 3.0) {
            x = 1.0;
        } else if (y < 1.0) {
            double w = y * y;
            x = ((((((((0.000124818987 * w - 0.001075204047) * w + 0.005198775019) * w - 0.019198292004) * w +
                     0.059054035642) * w - 0.151968751364) * w + 0.319152932694) * w - 0.531923007300) * w +
                 0.797884560593) * y * 2.0;
        } else {
            y -= 2.0;
            x = (((((((((((((-0.000045255659 * y + 0.000152529290) * y - 0.000019538132) * y - 0.000676904986) * y +
                          0.001390604284) * y - 0.000794620820) * y - 0.002034254874) * y + 0.006549791214) * y -
                      0.010557625006) * y + 0.011630447319) * y - 0.009279453341) * y + 0.005353579108) * y -
                  0.002141268741) * y + 0.000535310849) * y + 0.999936657524;
        }
    }
    return z > 0.0 ? (x + 1.0) * 0.5 : (1.0 - x) * 0.5;
}

static inline double ex20(double x) { return x < -20.0 ? 0.0 : exp(x); }

double pochisq(double x, int df) {
    if (x <= 0.0 || df < 1) return 1.0;
    const double a = 0.5 * x;
    const bool even = (2 * (df / 2)) == df;
    double y = 0.0;
    if (df > 1) y = ex20(-a);
    double s = even ? y : 2.0 * poz(-sqrt(x));
    if (df <= 2) return s;
    const double xm = 0.5 * (df - 1.0);
    double z = even ? 1.0 : 0.5;
    if (a > 20.0) {
        double e = even ? 0.0 : 0.5723649429247000870717135;
        const double c = log(a);
        while (z <= xm) {
            e = log(z) + e;
            s += ex20(c * z - a - e);
            z += 1.0;
        }
        return s;
    }
    double e = even ? 1.0 : 0.5641895835477562869480795 / sqrt(a);
    double c = 0.0;
    while (z <= xm) {
        e = e * (a / z);
        c = c + e;
        z += 1.0;
    }
    return c * y + s;
}

}  // namespace specfun
