#include "audriver.h"

#include <chrono>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <future>
#include <mutex>
#include <thread>

#include "refio.h"
#include "specfun.h"

namespace audriver {

// The fits of one lockstep round are independent (one per item, or one per (item, alpha,
// au|bp) in the confidence-limit rounds) and each costs tens of microseconds (ten qnorm
// bisections): spread them one at a time over a small persistent pool of host threads.
namespace {
class Pool {
  public:
    explicit Pool(int nthreads) {
        for (int t = 0; t < nthreads; t++) workers.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_all();
        for (auto &w : workers) w.join();
    }
    int size() const { return (int)workers.size(); }
    void run(int n, const std::function<void(int)> &body) {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = &body;
            total = n;
            next.store(0);
            pending = (int)workers.size();
            epoch++;
        }
        cv.notify_all();
        work(body, n);  // the caller helps
        std::unique_lock<std::mutex> lk(mu);
        done_cv.wait(lk, [this] { return pending == 0; });
        fn = nullptr;
    }

  private:
    void work(const std::function<void(int)> &body, int n) {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            body(i);
        }
    }
    void loop() {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int)> *f;
            int n;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                f = fn;
                n = total;
            }
            work(*f, n);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) done_cv.notify_all();
            }
        }
    }
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv, done_cv;
    const std::function<void(int)> *fn = nullptr;
    std::atomic<int> next{0};
    int total = 0, pending = 0;
    unsigned epoch = 0;
    bool quit = false;
};
}  // namespace

static void parallel_for(int n, const std::function<void(int)> &body) {
    if (refio::debugmode > 0 || n < 2) {  // keep debug traces in item order
        for (int i = 0; i < n; i++) body(i);
        return;
    }
    static Pool pool([] {
        // one process per GPU shares the host: torchrun exports LOCAL_WORLD_SIZE
        unsigned hw = std::thread::hardware_concurrency();
        unsigned share = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = (unsigned)std::max(1, atoi(e));
        if (const char *e = getenv("CHOST_THREADS")) return std::max(1, atoi(e));
        hw = std::max(1u, hw / share);
        return (int)std::max(1u, hw > 1 ? hw - 1 : 1u);
    }());
    pool.run(n, body);
}

static double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Two-half lockstep: the items are cut into two row ranges that take turns on the device, so
// the host fits of one half run while the other half's count query is in flight (the query
// is issued from a helper thread; the fits never call into the device library).  `prep(h)`
// fills half h's thresholds and returns how many per row (0 = nothing left to do), `ask(h, nq)`
// runs the query, `fit(h, nq)` consumes the answers.  With one half everything is sequential.
template <class Prep, class Ask, class Fit>
static int lockstep(int nhalf, Prep prep, Ask ask, Fit fit, double *tq, double *tf, int *nround) {
    int nq[2] = {0, 0};       // thresholds per row of the query being prepared / answered
    bool pending[2] = {false, false};
    for (int h = 0; h < nhalf; h++) nq[h] = prep(h);
    for (;;) {
        bool any = false;
        for (int h = 0; h < nhalf; h++) {
            const int o = nhalf == 2 ? 1 - h : h;  // the half whose answers are fitted meanwhile
            std::future<int> fut;
            const bool asked = nq[h] > 0;
            const double w0 = wall_now();
            if (asked) {
                if (nhalf == 2) fut = std::async(std::launch::async, [&, h] { return ask(h, nq[h]); });
                else if (int rc = ask(h, nq[h])) return rc;
                (*nround)++;
            }
            const double w1 = wall_now();
            if (nhalf == 2 && pending[o]) {
                fit(o, nq[o]);
                pending[o] = false;
                nq[o] = prep(o);
            }
            const double w2 = wall_now();
            if (asked && nhalf == 2)
                if (int rc = fut.get()) return rc;
            const double w3 = wall_now();
            if (asked) {
                if (nhalf == 2) {
                    pending[h] = true;
                } else {
                    fit(h, nq[h]);
                    nq[h] = prep(h);
                }
            }
            *tf += nhalf == 2 ? w2 - w1 : wall_now() - w3;
            *tq += nhalf == 2 ? w3 - w2 : w1 - w0;  // overlapped: only the part of the query the fits did not hide
            any = any || asked || pending[0] || pending[1];
        }
        if (!any) break;
    }
    return 0;
}

// Sharded runs: one half, every round is a collective; the loop ends when NO rank is active.
template <class Prep, class Ask, class Fit>
static int lockstep_dist(Prep prep, Ask ask, Fit fit, double *tq, double *tf, int *nround) {
    for (;;) {
        const int nq = prep(0);  // 0: this rank's items are done (it still serves the round)
        int any = 0;
        const double w0 = wall_now();
        if (int rc = ask(nq, &any)) return rc;
        const double w1 = wall_now();
        (*nround)++;
        *tq += w1 - w0;
        if (!any) break;
        if (nq > 0) fit(0, nq);
        *tf += wall_now() - w1;
    }
    return 0;
}

int run(const Input &in, const aufit::Options &opt, QueryFn query, void *ctx, Output *out, QueryRowsFn query_rows,
        DistQueryFn dist_query) {
    const int cm = in.cm, kk = in.kk;
    static const bool trace = getenv("CHOST_TIMING") != nullptr;  // per-phase wall times on stderr
    static const bool no_overlap = getenv("CHOST_NO_OVERLAP") != nullptr;
    double tq = 0, tf = 0;
    int nround = 0;
    std::vector<double> bbd(kk);
    for (int k = 0; k < kk; k++) bbd[k] = (double)in.bb[k];
    const double t0 = refio::cpu_seconds();
    // row ranges of the halves
    const int nhalf = (!dist_query && query_rows && cm >= 16 && refio::debugmode <= 0 && !no_overlap) ? 2 : 1;
    const int lo[2] = {0, nhalf == 2 ? cm / 2 : cm}, hi[2] = {nhalf == 2 ? cm / 2 : cm, cm};
    auto ask_rows = [&](int h, int nq, const double *thr, cgpu_rowsum *rs) {
        if (nhalf == 1) return query(ctx, nq, thr, rs);
        return query_rows(ctx, lo[h], hi[h] - lo[h], nq, thr, rs);
    };

    // ---- vcalpval for every item (src/consel.c:2744-2808), one query per round and half
    std::vector<aufit::VcIter> vc;
    vc.reserve(cm);
    for (int i = 0; i < cm; i++) vc.emplace_back(in.median[i], in.threshold);
    std::vector<double> thr(std::max(cm, 1), 1.0e300);
    std::vector<cgpu_rowsum> rs((size_t)std::max(cm, 1) * kk);
    int vround[2] = {0, 0};
    {
        auto prep = [&](int h) {
            if (vround[h] >= opt.vcloopmax) return 0;
            bool any = false;
            for (int i = lo[h]; i < hi[h]; i++) {
                thr[i] = vc[i].x;
                any = any || !vc[i].done;
            }
            if (!any) return 0;
            vround[h]++;
            return 1;
        };
        auto ask = [&](int h, int) { return ask_rows(h, 1, thr.data() + lo[h], rs.data() + (size_t)lo[h] * kk); };
        auto fit = [&](int h, int) {
            parallel_for(hi[h] - lo[h], [&](int e) {
                const int i = lo[h] + e;
                if (vc[i].done) return;
                std::vector<double> cnts(kk);
                for (int k = 0; k < kk; k++) cnts[k] = cgpu_smooth_count(&rs[(size_t)i * kk + k], in.bb[k], thr[i]);
                vc[i].step(cnts.data(), in.rr, bbd.data(), kk, opt);
            });
        };
        if (dist_query) {
            auto dask = [&](int nq1, int *any) { return dist_query(ctx, 1, thr.data(), rs.data(), nq1 > 0, any); };
            if (int rc = lockstep_dist(prep, dask, fit, &tq, &tf, &nround)) return rc;
        } else if (int rc = lockstep(nhalf, prep, ask, fit, &tq, &tf, &nround)) return rc;
    }
    if (trace) fprintf(stderr, "[audriver] vcalpval: %d queries, exposed query %.2f ms, fits %.2f ms\n", nround, tq * 1e3, tf * 1e3);
    tq = tf = 0;
    nround = 0;
    out->pv.assign(cm, 0); out->se.assign(cm, 0); out->pv0.assign(cm, 0); out->se0.assign(cm, 0);
    out->rss.assign(cm, 0); out->df.assign(cm, 0); out->th.assign(cm, 0); out->pf.assign(cm, 0);
    out->beta.assign(3 * (size_t)cm, 0); out->ret.assign(cm, 0);
    for (int i = 0; i < cm; i++) {
        const aufit::VcIter &it = vc[i];
        const int ret = it.done ? it.ret : 2;
        out->ret[i] = ret;
        out->pv[i] = it.pv; out->se[i] = it.se; out->pv0[i] = it.pv0; out->se0[i] = it.se0;
        out->rss[i] = it.rss; out->df[i] = it.df; out->th[i] = it.th_out;
        if (!ret) {
            out->pf[i] = specfun::pochisq(it.rss, (int)it.df);
            for (int c = 0; c < 3; c++) out->beta[3 * (size_t)i + c] = it.beta[c];
        }
    }
    out->t_fit_sec = refio::cpu_seconds() - t0;
    if (!in.confidence || in.nalpha <= 0) return 0;

    // ---- invtpval for every (item, alpha, au|bp) (src/consel.c:2816-2879)
    const int nalpha = in.nalpha, nq = 2 * nalpha;
    std::vector<aufit::CiIter> ci;
    ci.reserve((size_t)cm * nq);
    for (int j = 0; j < cm; j++) {
        const cgpu_rowsum &e = in.r0[(size_t)j * kk + in.i0];
        for (int a = 0; a < nalpha; a++) {
            ci.emplace_back(e.min1, e.max1, in.alpha[a], opt.kappa);
            ci.emplace_back(e.min1, e.max1, in.alpha[a], -1.0);
        }
    }
    const int qchunk = 16;  // thresholds per item and query
    std::vector<double> cthr;
    std::vector<cgpu_rowsum> crs;
    // Each round queries only the searches that have not converged yet: per item the active
    // thresholds are packed to the front (padded with a value above every statistic), so the
    // device pass gets narrower as the round count grows.
    std::vector<int> act((size_t)cm * nq), nact(cm);
    int cround[2] = {0, 0};
    {
        auto prep = [&](int h) {
            if (cround[h] >= opt.ciloopmax) return 0;
            int nmax = 0;
            for (int j = lo[h]; j < hi[h]; j++) {
                int n = 0;
                for (int q = 0; q < nq; q++)
                    if (!ci[(size_t)j * nq + q].done) act[(size_t)j * nq + n++] = q;
                nact[j] = n;
                nmax = std::max(nmax, n);
            }
            if (nmax == 0) return 0;
            cround[h]++;
            return nmax;
        };
        // thresholds and answers of half h live at row offset lo[h] of buffers sized for nq per row
        cthr.assign((size_t)std::max(cm, 1) * nq, 1.0e300);
        crs.resize((size_t)std::max(cm, 1) * kk * nq);
        auto ask = [&](int h, int nmax) {
            const int rows = hi[h] - lo[h];
            double *t = cthr.data() + (size_t)lo[h] * nq;
            for (int e = 0; e < rows; e++) {
                const int j = lo[h] + e;
                for (int a = 0; a < nmax; a++)
                    t[(size_t)e * nmax + a] = a < nact[j] ? ci[(size_t)j * nq + act[(size_t)j * nq + a]].x : 1.0e300;
            }
            cgpu_rowsum *o = crs.data() + (size_t)lo[h] * kk * nq;
            for (int a0 = 0; a0 < nmax; a0 += qchunk) {  // more than 16 thresholds per row: several queries
                if (a0 == 0 && nmax <= qchunk) return ask_rows(h, nmax, t, o);
                // (never taken with the reference's five alphas; kept simple: one query per chunk)
                const int nqc = std::min(qchunk, nmax - a0);
                std::vector<double> tt((size_t)rows * nqc);
                std::vector<cgpu_rowsum> oo((size_t)rows * kk * nqc);
                for (int e = 0; e < rows; e++)
                    for (int a = 0; a < nqc; a++) tt[(size_t)e * nqc + a] = t[(size_t)e * nmax + a0 + a];
                if (int rc = ask_rows(h, nqc, tt.data(), oo.data())) return rc;
                for (int e = 0; e < rows; e++)
                    for (int k = 0; k < kk; k++)
                        for (int a = 0; a < nqc; a++)
                            o[((size_t)e * kk + k) * nmax + a0 + a] = oo[((size_t)e * kk + k) * nqc + a];
            }
            return 0;
        };
        auto fit = [&](int h, int nmax) {
            const int rows = hi[h] - lo[h];
            const double *t = cthr.data() + (size_t)lo[h] * nq;
            const cgpu_rowsum *o = crs.data() + (size_t)lo[h] * kk * nq;
            parallel_for(rows * nmax, [&](int e2) {
                const int e = e2 / nmax, a = e2 % nmax, j = lo[h] + e;
                if (a >= nact[j]) return;
                aufit::CiIter &c = ci[(size_t)j * nq + act[(size_t)j * nq + a]];
                std::vector<double> cnts(kk);
                const double tv = t[(size_t)e * nmax + a];
                for (int k = 0; k < kk; k++) cnts[k] = cgpu_smooth_count(&o[((size_t)e * kk + k) * nmax + a], in.bb[k], tv);
                c.step(cnts.data(), in.rr, bbd.data(), kk, opt);
            });
        };
        if (dist_query) {
            // fixed width on every rank: all nq searches of an item, qchunk at a time (converged
            // searches and finished ranks send a threshold above every statistic)
            const int nfix = std::min(nq, qchunk);
            auto dprep = [&](int h) { return prep(h) > 0 ? nq : 0; };
            auto dask = [&](int nloc, int *any) {
                const int rows = cm;
                double *t = cthr.data();
                for (int e = 0; e < rows; e++)
                    for (int a = 0; a < nq; a++)
                        t[(size_t)e * nq + a] = (nloc > 0 && a < nact[e]) ? ci[(size_t)e * nq + act[(size_t)e * nq + a]].x : 1.0e300;
                cgpu_rowsum *o = crs.data();
                *any = 0;
                if (nq <= qchunk) return dist_query(ctx, nq, t, o, nloc > 0, any);
                for (int a0 = 0; a0 < nq; a0 += nfix) {  // more than 16 searches per item: fixed-size chunks
                    std::vector<double> tt((size_t)std::max(rows, 1) * nfix, 1.0e300);
                    std::vector<cgpu_rowsum> oo((size_t)std::max(rows, 1) * kk * nfix);
                    const int nqc = std::min(nfix, nq - a0);
                    for (int e = 0; e < rows; e++)
                        for (int a = 0; a < nqc; a++) tt[(size_t)e * nfix + a] = t[(size_t)e * nq + a0 + a];
                    int any1 = 0;
                    if (int rc = dist_query(ctx, nfix, tt.data(), oo.data(), nloc > 0, &any1)) return rc;
                    *any |= any1;
                    for (int e = 0; e < rows; e++)
                        for (int k = 0; k < kk; k++)
                            for (int a = 0; a < nqc; a++)
                                o[((size_t)e * kk + k) * nq + a0 + a] = oo[((size_t)e * kk + k) * nfix + a];
                }
                return 0;
            };
            if (int rc = lockstep_dist(dprep, dask, fit, &tq, &tf, &nround)) return rc;
        } else if (int rc = lockstep(nhalf, prep, ask, fit, &tq, &tf, &nround)) return rc;
    }
    if (trace) fprintf(stderr, "[audriver] invtpval: %d queries, exposed query %.2f ms, fits %.2f ms\n", nround, tq * 1e3, tf * 1e3);
    const size_t n2 = (size_t)cm * nalpha;
    out->ci.assign(n2, 0); out->cs.assign(n2, 0); out->ei.assign(n2, 0);
    out->ci0.assign(n2, 0); out->cs0.assign(n2, 0); out->ei0.assign(n2, 0);
    out->ci_ret.assign(2 * n2, 0); out->ci_warn.assign(2 * n2, 0); out->ci_warnval.assign(2 * n2, 0);
    for (int j = 0; j < cm; j++)
        for (int a = 0; a < nalpha; a++)
            for (int which = 0; which < 2; which++) {
                const aufit::CiIter &c = ci[(size_t)j * nq + 2 * a + which];
                const size_t e = (size_t)j * nalpha + a;
                (which ? out->ci0 : out->ci)[e] = c.cival;
                (which ? out->cs0 : out->cs)[e] = c.seval;
                (which ? out->ei0 : out->ei)[e] = c.eival;
                out->ci_ret[2 * e + which] = c.ret;
                out->ci_warn[2 * e + which] = c.degenerate_warning ? 1 : (c.noconv_warning ? 2 : 0);
                out->ci_warnval[2 * e + which] = c.warn_value;
            }
    return 0;
}

}  // namespace audriver

// ---- C export for the Python pipeline (consel_b200/pipeline.py)
// optional row-range query for the next chost_au_run call (same ctx): enables the overlapped schedule
static audriver::QueryRowsFn g_query_rows = nullptr;
extern "C" void chost_set_query_rows(audriver::QueryRowsFn fn) { g_query_rows = fn; }

extern "C" int chost_au_run(int cm, int kk, const double *rr, const int *bb, int i0, double threshold,
                            const double *median, const cgpu_rowsum *r0, int nalpha, const double *alpha,
                            int fitmode, double kappa, audriver::QueryFn query, void *ctx,
                            double *pv2 /*cm x 2: au, np*/, double *se2 /*cm x 2*/, double *ax7 /*cm x 7*/,
                            double *ci6 /*6 x cm x nalpha: CI SE EI CI0 SE0 EI0*/, int *ret /*cm*/) {
    audriver::Input in;
    in.cm = cm; in.kk = kk; in.rr = rr; in.bb = bb; in.i0 = i0; in.threshold = threshold;
    in.median = median; in.r0 = r0; in.nalpha = nalpha; in.alpha = alpha; in.confidence = ci6 != nullptr;
    aufit::Options opt;
    opt.fitmode = fitmode;
    opt.kappa = kappa;
    audriver::Output o;
    int rc = audriver::run(in, opt, query, ctx, &o, g_query_rows);
    g_query_rows = nullptr;
    if (rc) return rc;
    for (int i = 0; i < cm; i++) {
        pv2[2 * i] = o.pv[i]; pv2[2 * i + 1] = o.pv0[i];
        se2[2 * i] = o.se[i]; se2[2 * i + 1] = o.se0[i];
        double *ax = ax7 + 7 * (size_t)i;
        ax[0] = o.pf[i]; ax[1] = o.rss[i]; ax[2] = o.df[i]; ax[3] = o.beta[3 * (size_t)i];
        ax[4] = o.beta[3 * (size_t)i + 1]; ax[5] = o.th[i]; ax[6] = o.beta[3 * (size_t)i + 2];
        if (ret) ret[i] = o.ret[i];
    }
    if (ci6) {
        const size_t n2 = (size_t)cm * nalpha;
        const std::vector<double> *m[6] = {&o.ci, &o.cs, &o.ei, &o.ci0, &o.cs0, &o.ei0};
        for (int t = 0; t < 6; t++) std::copy(m[t]->begin(), m[t]->end(), ci6 + t * n2);
    }
    return 0;
}
