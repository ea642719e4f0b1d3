#include "audriver.h"

#include <chrono>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
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
        unsigned hw = std::thread::hardware_concurrency();
        return (int)std::max(1u, hw ? hw - 1 : 1u);
    }());
    pool.run(n, body);
}

static double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int run(const Input &in, const aufit::Options &opt, QueryFn query, void *ctx, Output *out) {
    const int cm = in.cm, kk = in.kk;
    static const bool trace = getenv("CHOST_TIMING") != nullptr;  // per-phase wall times on stderr
    double tq = 0, tf = 0;
    int nround = 0;
    std::vector<double> bbd(kk);
    for (int k = 0; k < kk; k++) bbd[k] = (double)in.bb[k];
    const double t0 = refio::cpu_seconds();

    // ---- vcalpval for every item (src/consel.c:2744-2808), one query per round
    std::vector<aufit::VcIter> vc;
    vc.reserve(cm);
    for (int i = 0; i < cm; i++) vc.emplace_back(in.median[i], in.threshold);
    std::vector<double> thr(cm);
    std::vector<cgpu_rowsum> rs((size_t)cm * kk);
    for (int round = 0; round < opt.vcloopmax; round++) {
        bool any = false;
        for (int i = 0; i < cm; i++) {
            thr[i] = vc[i].x;
            any = any || !vc[i].done;
        }
        if (!any) break;
        double w0 = wall_now();
        if (int rc = query(ctx, 1, thr.data(), rs.data())) return rc;
        double w1 = wall_now();
        tq += w1 - w0;
        nround++;
        parallel_for(cm, [&](int i) {
            if (vc[i].done) return;
            std::vector<double> cnts(kk);
            for (int k = 0; k < kk; k++) cnts[k] = cgpu_smooth_count(&rs[(size_t)i * kk + k], in.bb[k], thr[i]);
            vc[i].step(cnts.data(), in.rr, bbd.data(), kk, opt);
        });
        tf += wall_now() - w1;
    }
    if (trace) fprintf(stderr, "[audriver] vcalpval: %d rounds, query %.2f ms, fits %.2f ms\n", nround, tq * 1e3, tf * 1e3);
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
    for (int round = 0; round < opt.ciloopmax; round++) {
        int nmax = 0;
        for (int j = 0; j < cm; j++) {
            int n = 0;
            for (int q = 0; q < nq; q++)
                if (!ci[(size_t)j * nq + q].done) act[(size_t)j * nq + n++] = q;
            nact[j] = n;
            nmax = std::max(nmax, n);
        }
        if (nmax == 0) break;
        for (int a0 = 0; a0 < nmax; a0 += qchunk) {
            const int nqc = std::min(qchunk, nmax - a0);
            cthr.assign((size_t)cm * nqc, 1.0e300);
            for (int j = 0; j < cm; j++)
                for (int a = 0; a < nqc && a0 + a < nact[j]; a++)
                    cthr[(size_t)j * nqc + a] = ci[(size_t)j * nq + act[(size_t)j * nq + a0 + a]].x;
            crs.resize((size_t)cm * kk * nqc);
            double w0 = wall_now();
            if (int rc = query(ctx, nqc, cthr.data(), crs.data())) return rc;
            double w1 = wall_now();
            tq += w1 - w0;
            nround++;
            parallel_for(cm * nqc, [&](int e) {
                const int j = e / nqc, a = e % nqc;
                if (a0 + a >= nact[j]) return;
                aufit::CiIter &c = ci[(size_t)j * nq + act[(size_t)j * nq + a0 + a]];
                std::vector<double> cnts(kk);
                const double t = cthr[(size_t)j * nqc + a];
                for (int k = 0; k < kk; k++)
                    cnts[k] = cgpu_smooth_count(&crs[((size_t)j * kk + k) * nqc + a], in.bb[k], t);
                c.step(cnts.data(), in.rr, bbd.data(), kk, opt);
            });
            tf += wall_now() - w1;
        }
    }
    if (trace) fprintf(stderr, "[audriver] invtpval: %d rounds, query %.2f ms, fits %.2f ms\n", nround, tq * 1e3, tf * 1e3);
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
    int rc = audriver::run(in, opt, query, ctx, &o);
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
