// consel — drop-in command line of the reference's consel (src/consel.c), rmt mode
// (do_rmtmode, :800-883).  The replicate-statistics loops run on the GPU through the C
// ABI: cgpu_mctest (do_mctest :1352-1439), cgpu_stats_build (genrep/repminmaxs :751-787,
// 1939-1959), cgpu_count_query + cgpu_order_stat (every cntdist / sorted-row access of
// do_bptest :1286-1303, do_bootrep :1441-1538, vcalpval :2744-2808, invtpval :2816-2879).
// The per-item AU curve fit stays on the host (aufit.cpp) and is driven in LOCKSTEP over
// all items so that one device query per round serves every item; the fits of a round
// run on all host threads.
//
// Rep mode (-R, do_repmode :1195-1242) feeds statistic rows that already exist (a `.rep`
// file) to the same count queries through cgpu_stats_upload: BP and AU only, like the
// reference.
//
// The rescaling approximation (-f [-p scales.pa], genrep :757-771) derives the statistics of
// every target scale from the replicates of the scale nearest r=1 on the device
// (cgpu_stats_build_rescaled).
//
// Cnt mode (-C, do_cntmode :1244-1283) is host arithmetic only in the reference too: the AU
// fit from the counts of a `.cnt` file, no replicates involved.
//
// Multi-gene mode (-g, do_rmtmultimode :1056-1192): per gene the statistics and one count
// query on the device, then congcnt and the count-matrix fit on the host.
//
// Not carried by this build (error at the command line): the --chi* spherical fit (off the
// hot path, SURVEY 8f "next").
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/consel_gpu.h"
#include "assoc.h"
#include "audriver.h"
#include "pipeline.h"
#include "aufit.h"
#include "refio.h"
#include "specfun.h"

using namespace refio;

static void byebye() { die("error in command line"); }
static void dots(int n) {
    for (int i = 0; i < n; i++) putchar('.');
    fflush(stdout);
}
// The "small variance" warnings of do_mctest (src/consel.c:1403-1415): the pairwise sums
// x_ij = sum_b (xc_ib - xc_jb)^2 over the centred r~1 block are formed on the device (pair_weight_kernel), which
// reports no per-pair diagnostics.  For blocks small enough (mm^2 * nb <= 2e9: every case the reference finishes
// in seconds) the sums are replayed here on the host threads in the reference's order of operations, and the
// warnings are printed where the reference prints them, between the progress dots of the rows; larger blocks
// print the dots alone.
static void small_variance_warnings(const double *rep0, int mm, int nb, double varadd) {
    const double kWarn = 100.0;  // VARADDWARN, src/consel.c:154
    std::vector<std::vector<std::pair<int, double>>> small(mm);
    if (rep0 && (double)mm * mm * nb <= 2e9) {
        std::vector<double> xc(rep0, rep0 + (size_t)mm * nb);
        for (int j = 0; j < mm; j++) {  // centring, :1374-1382
            double *row = &xc[(size_t)j * nb];
            double x = 0.0;
            for (int i = 0; i < nb; i++) x += row[i];
            x = x / nb;
            for (int i = 0; i < nb; i++) row[i] -= x;
        }
        std::atomic<int> next{0};
        auto work = [&]() {
            for (int i = next.fetch_add(1); i < mm; i = next.fetch_add(1)) {
                const double *xp = &xc[(size_t)(mm - 1 - i) * nb];  // long rows first
                const int ii = mm - 1 - i;
                for (int j = 0; j < ii; j++) {
                    const double *yp = &xc[(size_t)j * nb];
                    double x = 0.0;
                    for (int ib = 0; ib < nb; ib++) x += (xp[ib] - yp[ib]) * (xp[ib] - yp[ib]);
                    if (x < varadd * kWarn) small[ii].push_back({j, x});
                }
            }
        };
        unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        if ((double)mm * mm * nb < 4e6) nt = 1;
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    for (int i = 0; i < mm; i++) {
        for (const auto &w : small[i]) warn("small variance [%d,%d]: %g<%g*%g", i + 1, w.first + 1, w.second, varadd, kWarn);
        putchar('.');
    }
    fflush(stdout);
}

#define CG(call)                                           \
    do {                                                   \
        if ((call) != CGPU_OK) die("%s", cgpu_last_error()); \
    } while (0)



// ---------------------------------------------------------------- count-matrix modes
// State shared by cnt mode (-C) and multi-gene mode (-g): per-item scale, replicate-count and
// count matrices (cm x kk each) and what do_bootcnt / write_pv derive from them.
struct CntWork {
    int cm = 0, kk = 0;
    std::vector<int> orderv;
    std::vector<double> obsvec, rrmat, bbmat, cntmat;
    std::vector<double> npvec, nsvec, bapvec, pvvec, sevec, pv0vec, se0vec, rssvec, dfvec, pfvec, thvec, beta;
};

// BP from the counts of the scale nearest r=1 (src/consel.c:1268-1275, :1168-1175)
static void cnt_bptest(CntWork &w) {
    w.npvec.assign(w.cm, 0.0);
    w.nsvec.assign(w.cm, 0.0);
    for (int j = 0; j < w.cm; j++) {
        const double *rr = &w.rrmat[(size_t)j * w.kk];
        int i0 = 0;
        for (int i = 1; i < w.kk; i++)
            if (fabs(rr[i] - 1.0) < fabs(rr[i0] - 1.0)) i0 = i;
        const double b = w.bbmat[(size_t)j * w.kk + i0];
        w.npvec[j] = w.cntmat[(size_t)j * w.kk + i0] / b;
        w.nsvec[j] = sqrt(w.npvec[j] * (1.0 - w.npvec[j]) / b);
    }
}

// do_bootcnt (src/consel.c:1540-1592).  rcalpval (:2703-2737) fits on the scales from the
// smallest r >= rmin that is below 1 upwards, drops the smallest scale and fits again while
// the fit's p-value is below pfmin (--pmin; the default 0 accepts the first fit), and reports
// "degenerated" without fitting when there is no scale in [rmin, 1).
static void cnt_bootcnt(CntWork &w, double threshold, const aufit::Options &opt, double pfmin) {
    const int cm = w.cm, kk = w.kk;
    for (auto *v : {&w.pvvec, &w.sevec, &w.pv0vec, &w.se0vec, &w.rssvec, &w.dfvec, &w.pfvec, &w.thvec}) v->assign(cm, 0.0);
    w.beta.assign(3 * (size_t)cm, 0.0);
    printf("\n# AU-TEST STARTS");
    printf("\n# calculating approximately unbiased p-values by ");
    printf(opt.fitmode == aufit::FIT_WLS ? "WLS (very fast)" : "MLE (fast)");
    printf(" fitting");
    fflush(stdout);
    const double t0 = cpu_seconds();
    for (int i = 0; i < cm; i++) {
        const double *rr = &w.rrmat[(size_t)i * kk];
        std::vector<double> rs(rr, rr + kk);
        std::sort(rs.begin(), rs.end());
        int i1 = 0;
        while (i1 < kk && rs[i1] < opt.rrmin) i1++;
        int i2 = i1;
        while (i2 < kk && rs[i2] < 1.0) i2++;
        aufit::Fit f;
        int j = 1;
        for (int t = i1; t < i2; t++) {  // raise rmin until the fit is acceptable (--pmin; default 0 = first fit)
            f = aufit::calcpval(&w.cntmat[(size_t)i * kk], rr, &w.bbmat[(size_t)i * kk], kk, rs[t], opt.rrmax, opt.kappa, opt);
            j = f.rc;
            if (j) break;
            if (specfun::pochisq(f.rss, (int)f.df) >= pfmin) break;
        }
        w.pvvec[i] = f.pv; w.sevec[i] = f.se; w.pv0vec[i] = f.pv0;
        if (cm > 1) w.se0vec[1] = f.se0;  // the reference passes `se0vec+1` for every item (:1567)
        w.rssvec[i] = f.rss; w.dfvec[i] = f.df;
        w.thvec[i] = threshold;  // unused in the count-matrix modes
        dprint(1, "\n# rank=%d item=%d", i + 1, w.orderv[i] + 1);
        dprint(1, "\n# ret=%d", j);
        if (j) {
            w.pfvec[i] = 0.0;
            warn("regression degenerated: df[%d]=%g", i + 1, w.dfvec[i]);
        } else {
            w.pfvec[i] = specfun::pochisq(w.rssvec[i], (int)w.dfvec[i]);
            if (w.pfvec[i] < 0.01) warn("theory does not fit well: pfit[%d]=%.4f", i + 1, w.pfvec[i]);
            for (int c = 0; c < 3; c++) w.beta[3 * (size_t)i + c] = f.beta[c];
        }
        putchar('.');
        fflush(stdout);
    }
    printf("\n# time elapsed for AU test is t=%g sec", cpu_seconds() - t0);
    printf("\n# AU-TEST DONE");
}

// write_pv(sw_bp, sw_ba, 0, sw_au) (src/consel.c:1611-1683)
static void cnt_write_pv(const CntWork &w, const char *fname_out, bool dobp, bool doba, bool doau) {
    std::string used;
    FILE *fp = stdout;
    if (fname_out) {
        fp = open_file(fname_out, ".pv", "w", &used);
        printf("\n# writing %s", used.c_str());
    } else {
        printf("\n# writing to stdout\n");
    }
    const int cm = w.cm;
    const int outbit = (dobp ? 1 : 0) + (doba ? 2 : 0) + (doau ? 8 : 0);
    const int pvnum = (dobp ? 1 : 0) + (doba ? 1 : 0) + (doau ? 2 : 0), auxnum = doau ? 7 : 0;
    fprintf(fp, "\n# ID:\n%d\n", 1);
    fprintf(fp, "\n# ITEM:\n"); write_ivec(fp, w.orderv.data(), cm);
    fprintf(fp, "\n# STAT:\n"); write_vec(fp, w.obsvec.data(), cm);
    fprintf(fp, "\n# BIT:\n%d\n", outbit);
    std::vector<double> pvmat((size_t)cm * pvnum), semat((size_t)cm * pvnum), auxmat((size_t)cm * auxnum);
    for (int i = 0; i < cm; i++) {
        int j = 0;
        auto put = [&](double p, double se) {
            pvmat[(size_t)i * pvnum + j] = p;
            semat[(size_t)i * pvnum + j] = se;
            j++;
        };
        if (dobp) put(w.npvec[i], w.nsvec[i]);
        if (doba) put(w.bapvec[i], 0.0);
        if (doau) {
            put(w.pvvec[i], w.sevec[i]);
            put(w.pv0vec[i], w.se0vec[i]);
            double *ax = &auxmat[(size_t)i * auxnum];
            ax[0] = w.pfvec[i]; ax[1] = w.rssvec[i]; ax[2] = w.dfvec[i];
            ax[3] = w.beta[3 * (size_t)i]; ax[4] = w.beta[3 * (size_t)i + 1];
            ax[5] = w.thvec[i]; ax[6] = w.beta[3 * (size_t)i + 2];
        }
    }
    fprintf(fp, "\n# PV:\n"); write_mat(fp, pvmat.data(), cm, pvnum);
    fprintf(fp, "\n# SE:\n"); write_mat(fp, semat.data(), cm, pvnum);
    fprintf(fp, "\n# AX:\n"); write_mat(fp, auxnum ? auxmat.data() : pvmat.data(), cm, auxnum);
    if (fname_out) fclose(fp);
}

// cnt mode (consel -C): do_cntmode (src/consel.c:1244-1283).  No replicates, no device: the
// reference's path is the same host arithmetic.
static void run_cntmode(const char *fname_in, const char *fname_out, bool dobp, bool doau, double threshold,
                        const aufit::Options &opt, double pfmin) {
    std::string used;
    FILE *fp = stdin;
    if (fname_in) {
        fp = open_file(fname_in, ".cnt", "r", &used);
        printf("\n# reading %s", used.c_str());
    } else {
        printf("\n# reading from stdin");
    }
    CntWork w;
    w.orderv = read_ivec(fp);
    w.cm = (int)w.orderv.size();
    w.obsvec = read_vec(fp, w.cm);
    w.rrmat = read_mat(fp, &w.cm, &w.kk);
    w.bbmat = read_mat(fp, &w.cm, &w.kk);
    w.cntmat = read_mat(fp, &w.cm, &w.kk);
    if (fname_in) fclose(fp);
    printf("\n# CM:%d", w.cm);
    printf("\n# K:%d", w.kk);
    if (dobp) cnt_bptest(w);
    if (doau) cnt_bootcnt(w, threshold, opt, pfmin);
    if (dobp || doau) cnt_write_pv(w, fname_out, dobp, false, doau);
}

// ---------------------------------------------------------------- multi-gene mode (consel -g)
// do_rmtmultimode (src/consel.c:1056-1192): one .rmt per gene, the items' counts of every gene
// (gencnt :892-951, here cgpu_stats_build[_rescaled] + one count query per gene on the device),
// combined into effective counts / replicate numbers by congcnt (:961-1043), then the
// count-matrix BP and AU fit above.  The reference marks the mode experimental; its quirks are
// kept (the congruence row sums the first MF posterior entries, :1146).
struct MultiArgs {
    const char *fname_in, *fname_out, *fname_ass, *fname_vt, *fname_pa;
    bool dobp, doba, doau, outcnt, smoothcnt, nosort, fastrep;
    int cm_arg, gpu;
    double threshold, bapcoef, congbpadd, pfmin;
};

static void run_multimode(const MultiArgs &A, const aufit::Options &opt) {
    std::string used;
    FILE *fp = A.fname_in ? open_file(A.fname_in, ".svt", "r", &used) : stdin;
    const std::vector<std::string> infiles = read_svec(fp);
    if (A.fname_in) fclose(fp);
    const int mf = (int)infiles.size();
    printf("\n# MF:%d", mf);

    int mm = 0, kk = 0;
    std::vector<std::vector<double>> datvecm(mf), rrm(mf);
    std::vector<std::vector<int>> bbm(mf);
    for (int f = 0; f < mf; f++) {  // :1080-1093
        fp = open_file(infiles[f].c_str(), ".rmt", "rb", &used);
        printf("\n# checking %s", used.c_str());
        datvecm[f] = read_bvec(fp, mm);
        mm = (int)datvecm[f].size();
        rrm[f] = read_bvec(fp, kk);
        kk = (int)rrm[f].size();
        bbm[f] = read_bivec(fp, kk);
        printf("\n# R:");
        for (int i = 0; i < kk; i++) printf("%6.2f ", rrm[f][i]);
        printf("\n# B:");
        for (int i = 0; i < kk; i++) printf("%6d ", bbm[f][i]);
        for (int i = 0; i < kk; i++)
            if (fabs(rrm[f][i] - rrm[0][i]) > rrm[0][i] * 0.1) die("rr mismatch");
        if (read_bint(fp) != kk) die("wrong size in rmt");
        fclose(fp);
    }
    printf("\n# K:%d", kk);
    printf("\n# M:%d", mm);
    if (kk < 1 || mf < 1) return;

    std::vector<double> datvec0(mm);
    for (int i = 0; i < mm; i++) {  // total log-likelihood, :1101-1105
        double x = 0.0;
        for (int f = 0; f < mf; f++) x += datvecm[f][i];
        datvec0[i] = x;
    }
    Assoc ass;
    CntWork w;
    int cm = read_assoc_and_order(A.fname_ass, datvec0, A.cm_arg, A.nosort, &ass, &w.obsvec, &w.orderv);

    bool cong = true;  // sw_cong, :230; dropped when the associations overlap (:1112-1119)
    {
        std::vector<int> hits(mm, 0);
        for (int i = 0; i < cm; i++)
            for (int q = 0; q < ass.len(i); q++) hits[ass.at(i)[q]]++;
        for (int i = 0; i < mm; i++)
            if (hits[i] > 1) {
                printf("\n# non-disjoint");
                cong = false;
                break;
            }
    }
    const int cm2 = cm + (cong ? 1 : 0);
    w.obsvec.resize(cm2);
    w.orderv.resize(cm2);
    if (cong) {
        w.obsvec[cm] = 0.0;
        w.orderv[cm] = cm;
    }

    if (A.fname_vt) {  // read_alpha: the alphas are read (and announced) but unused in this mode
        fp = open_file(A.fname_vt, ".vt", "r", &used);
        printf("\n# reading %s", used.c_str());
        read_vec(fp);
        fclose(fp);
    }
    std::vector<double> rr1 = {0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4};
    if (A.fastrep) {  // read_scale, :562-582
        if (A.fname_pa) {
            fp = open_file(A.fname_pa, ".pa", "r", &used);
            printf("\n# reading %s", used.c_str());
            rr1 = read_vec(fp);
            fclose(fp);
        }
        printf("\n# RESCALING");
        printf("\n# K:%d", (int)rr1.size());
        printf("\n# R:");
        for (double r : rr1) printf("%g ", r);
    }

    if (A.doba) {  // per-gene posteriors multiplied (:1133-1151)
        w.bapvec.assign(cm2, 0.0);
        std::vector<double> prod(cm, 1.0);
        for (int f = 0; f < mf; f++) {
            const std::vector<double> &dv = datvecm[f];
            double x = -kHuge;
            for (int i = 0; i < mm; i++)
                if (dv[i] > x) x = dv[i];
            std::vector<double> p(mm);
            double y = 0.0;
            for (int i = 0; i < mm; i++) y += p[i] = exp(A.bapcoef * (dv[i] - x));
            for (int i = 0; i < mm; i++) p[i] = p[i] / y;
            for (int i = 0; i < cm; i++) {
                double t = 0.0;
                for (int q = 0; q < ass.len(i); q++) t += p[ass.at(i)[q]];
                prod[i] *= t;
            }
        }
        for (int i = 0; i < cm; i++) w.bapvec[i] = prod[i];
        if (cong) {
            double x = 0.0;
            for (int i = 0; i < mf && i < cm2; i++) x += w.bapvec[i];  // `i<mf` as in the reference (:1146)
            w.bapvec[cm] = x;
        }
    }

    if (A.doau || A.dobp || A.outcnt) {
        CG(cgpu_init(A.gpu));
        std::vector<std::vector<double>> cntm(mf);  // cm x kk per gene
        for (int f = 0; f < mf; f++) {  // gencnt
            fp = open_file(infiles[f].c_str(), ".rmt", "rb", &used);
            printf("\n# reading %s", used.c_str());
            std::vector<double> dv = read_bvec(fp, mm);
            std::vector<double> rr = read_bvec(fp);
            int k0 = (int)rr.size();
            std::vector<int> bb = read_bivec(fp, k0);
            if (read_bint(fp) != k0) die("wrong size in rmt");
            std::vector<std::vector<double>> repmats(k0);
            int i0 = 0;
            for (int i = 1; i < k0; i++)
                if (fabs(rr[i] - 1.0) < fabs(rr[i0] - 1.0)) i0 = i;
            const int nread = A.fastrep ? i0 + 1 : k0;  // the rescaling branch stops reading at scale i0 (:908-912)
            for (int i = 0; i < nread; i++) {
                int m = mm, b = bb[i];
                repmats[i] = read_bmat(fp, &m, &b);
            }
            fclose(fp);
            cgpu_reps *reps = nullptr, *stats = nullptr;
            if (A.fastrep) {
                std::vector<const double *> ptrs{repmats[i0].data()};
                CG(cgpu_reps_upload(mm, 1, &bb[i0], ptrs.data(), &reps));
                const int kk1 = (int)rr1.size();
                std::vector<double> rs(kk1);
                for (int i = 0; i < kk1; i++) rs[i] = sqrt(rr[i0] / rr1[i]);
                CG(cgpu_stats_build_rescaled(reps, 0, kk1, rs.data(), cm, ass.ptr.data(), ass.idx.data(), &stats));
                bb.assign(kk1, bb[i0]);
                rr = rr1;
            } else {
                std::vector<const double *> ptrs(k0);
                for (int i = 0; i < k0; i++) ptrs[i] = repmats[i].data();
                CG(cgpu_reps_upload(mm, k0, bb.data(), ptrs.data(), &reps));
                CG(cgpu_stats_build(reps, cm, ass.ptr.data(), ass.idx.data()));
            }
            const int kf = (int)rr.size();
            if (f > 0 && kf != kk) die("size mismatch in vec");
            kk = kf;
            dots(kk);
            std::vector<double> thr(cm, A.threshold);
            std::vector<cgpu_rowsum> q((size_t)cm * kk);
            CG(cgpu_count_query(stats ? stats : reps, 1, thr.data(), q.data()));
            cntm[f].resize((size_t)cm * kk);
            for (int j = 0; j < cm; j++)
                for (int i = 0; i < kk; i++) {
                    const cgpu_rowsum &e = q[(size_t)j * kk + i];
                    cntm[f][(size_t)j * kk + i] = A.smoothcnt ? cgpu_smooth_count(&e, bb[i], A.threshold) : e.cnt;
                }
            if (stats) cgpu_reps_free(stats);
            cgpu_reps_free(reps);
            bbm[f] = bb;
            rrm[f] = rr;
        }
        w.cm = cm2;
        w.kk = kk;
        w.rrmat.assign((size_t)cm2 * kk, 0.0);
        for (int i = 0; i < kk; i++) {  // averaged scales, :1161-1165
            double x = 0.0;
            for (int f = 0; f < mf; f++) x += rrm[f][i];
            x = x / mf;
            for (int j = 0; j < cm2; j++) w.rrmat[(size_t)j * kk + i] = x;
        }

        // ---- congcnt (:961-1043)
        auto at = [kk](std::vector<double> &m, int i, int j) -> double & { return m[(size_t)i * kk + j]; };
        std::vector<std::vector<double>> pm(mf, std::vector<double>((size_t)cm * kk));
        std::vector<double> pv((size_t)cm2 * kk, 0.0), va((size_t)cm2 * kk, 0.0);
        w.cntmat.assign((size_t)cm2 * kk, 0.0);
        w.bbmat.assign((size_t)cm2 * kk, 0.0);
        auto combine = [&]() {  // product over genes, congruence row = sum over items
            for (int j = 0; j < kk; j++)
                for (int i = 0; i < cm; i++) {
                    double x = 1.0;
                    for (int f = 0; f < mf; f++) x *= at(pm[f], i, j);
                    at(pv, i, j) = x;
                }
            if (cong)
                for (int j = 0; j < kk; j++) {
                    double x = 0.0;
                    for (int i = 0; i < cm; i++) x += at(pv, i, j);
                    at(pv, cm, j) = x;
                }
        };
        for (int f = 0; f < mf; f++)  // bootstrap probabilities pulled towards 1/2 by congbpadd counts
            for (int j = 0; j < kk; j++)
                for (int i = 0; i < cm; i++) {
                    const double x = at(cntm[f], i, j), y = bbm[f][j];
                    const double z = y * 0.5 + ((y - 2.0 * A.congbpadd) / y) * (x - y * 0.5);
                    at(pm[f], i, j) = z / y;
                }
        combine();
        for (int j = 0; j < kk; j++) {  // delta-method variance of the product
            double z = 0.0, y = 0.0;
            for (int f = 0; f < mf; f++) y += 1.0 / bbm[f][j];
            for (int i = 0; i < cm; i++) {
                double x = 0.0;
                if (at(pv, i, j) > 0.0)
                    for (int f = 0; f < mf; f++) x += 1.0 / (bbm[f][j] * at(pm[f], i, j));
                x *= at(pv, i, j) * at(pv, i, j);
                z += x;
                at(va, i, j) = x - y * (at(pv, i, j) * at(pv, i, j));
            }
            if (cong) at(va, cm, j) = z - y * (at(pv, cm, j) * at(pv, cm, j));
        }
        for (int j = 0; j < kk; j++)  // effective number of replicates
            for (int i = 0; i < cm2; i++) {
                const double x = at(pv, i, j);
                at(w.bbmat, i, j) = x * (1.0 - x) / at(va, i, j);
            }
        for (int f = 0; f < mf; f++)  // uncorrected probabilities
            for (int j = 0; j < kk; j++)
                for (int i = 0; i < cm; i++) at(pm[f], i, j) = at(cntm[f], i, j) / bbm[f][j];
        combine();
        for (int j = 0; j < kk; j++)
            for (int i = 0; i < cm2; i++) at(w.cntmat, i, j) = at(w.bbmat, i, j) * at(pv, i, j);
    }
    w.cm = cm2;
    if (A.dobp) cnt_bptest(w);
    if (A.doau) cnt_bootcnt(w, A.threshold, opt, A.pfmin);
    if (A.dobp || A.doau || A.doba) cnt_write_pv(w, A.fname_out, A.dobp, A.doba, A.doau);
    if (A.outcnt) {  // write_cnt2
        fp = stdout;
        if (A.fname_out) {
            fp = open_file(A.fname_out, ".cnt", "w", &used);
            printf("\n# writing %s", used.c_str());
        } else {
            printf("\n# writing to stdout\n");
        }
        fprintf(fp, "\n# ITEM:\n"); write_ivec(fp, w.orderv.data(), w.cm);
        fprintf(fp, "\n# STAT:\n"); write_vec(fp, w.obsvec.data(), w.cm);
        fprintf(fp, "\n# R:\n"); write_mat(fp, w.rrmat.data(), w.cm, w.kk);
        fprintf(fp, "\n# B:\n"); write_mat(fp, w.bbmat.data(), w.cm, w.kk);
        fprintf(fp, "\n# CNT:\n"); write_mat(fp, w.cntmat.data(), w.cm, w.kk);
        if (A.fname_out) fclose(fp);
    }
    if (A.doau || A.dobp || A.outcnt) cgpu_shutdown();
}

int main(int argc, char **argv) {
    if (getenv("RANK") && atoi(getenv("RANK")) > 0 && getenv("WORLD_SIZE") && atoi(getenv("WORLD_SIZE")) > 1)
        if (!freopen("/dev/null", "w", stdout)) return 1;  // rank 0 speaks for the job
    const char *fname_in = nullptr, *fname_out_arg = nullptr, *fname_ass = nullptr, *fname_vt = nullptr, *fname_pa = nullptr;
    bool fastrep = false;
    bool outrep = false, outcnt = false, smoothcnt = false, nosort = false, inrep = false, incnt = false, multi = false, quickmc = false;
    double congbpadd = 1.0;  // src/consel.c:960
    const char *fname_pdm = nullptr;  // --pdm: parameters per item for the AIC correction (:584-744)
    int aictype = 2;                  // :239
    double bapn = 1.0, pfmin = 0.0;   // :182, :2700
    bool dobp = true, doba = true, domc = true, doau = true;
    int cm_arg = 0, gpu = -1;
    double threshold = 0.0, varadd = 1.0, bapcoef = 1.0;
    aufit::Options opt;

    printf("# consel_b200 consel (sm_100a); command line of consel.c,v 1.20");

    for (int i = 1, j = 1; i < argc; i++) {  // src/consel.c:268-431
        const std::string a = argv[i];
        auto need = [&]() { if (i + 1 >= argc) byebye(); };
        auto dbl = [&](double *x) { need(); if (sscanf(argv[++i], "%lf", x) != 1) byebye(); };
        auto integer = [&](int *x) { need(); if (sscanf(argv[++i], "%d", x) != 1) byebye(); };
        if (argv[i][0] != '-') {
            if (j == 1) fname_in = argv[i];
            else if (j == 2) fname_out_arg = argv[i];
            else byebye();
            j++;
        } else if (a == "-r") outrep = true;
        else if (a == "-R") inrep = true;
        else if (a == "-C") incnt = true;
        else if (a == "-g") multi = true;
        else if (a == "--congbp") dbl(&congbpadd);
        else if (a == "-f") fastrep = true;
        else if (a == "-p") { need(); fname_pa = argv[++i]; }
        else if (a == "-c") outcnt = true;
        else if (a == "-a") { need(); fname_ass = argv[++i]; }
        else if (a == "-v") { need(); fname_vt = argv[++i]; }
        else if (a == "-d") integer(&debugmode);
        else if (a == "-t") integer(&cm_arg);
        else if (a == "--no_sort") nosort = true;
        else if (a == "--no_bp") dobp = false;
        else if (a == "--no_pp") doba = false;
        else if (a == "--no_sh") domc = false;
        else if (a == "--no_au") doau = false;
        else if (a == "--smooth_cnt") smoothcnt = true;
        else if (a == "--quick_mctest") quickmc = true;
        else if (a == "--th") dbl(&threshold);
        else if (a == "--kappa") dbl(&opt.kappa);
        else if (a == "--ppcoef") dbl(&bapcoef);
        else if (a == "--cieps") dbl(&opt.cieps);
        else if (a == "--vceps2") dbl(&opt.vceps2);
        else if (a == "--rmin") dbl(&opt.rrmin);
        else if (a == "--rmax") dbl(&opt.rrmax);
        else if (a == "--dmin") integer(&opt.dfmin);
        else if (a == "--mleeps") dbl(&opt.mleeps);
        else if (a == "--mle") opt.fitmode = aufit::FIT_MLE;
        else if (a == "--wls") opt.fitmode = aufit::FIT_WLS;
        else if (a == "--gpu") integer(&gpu);
        else if (a == "--pdm") { need(); fname_pdm = argv[++i]; }
        else if (a == "--aictype") integer(&aictype);
        else if (a == "--ppn") dbl(&bapn);
        else if (a == "--pmin") dbl(&pfmin);
        else if (a.rfind("--chi", 0) == 0)
            die("switch %s is not supported by the B200 build of consel (the spherical fit is not carried)", a.c_str());
        else byebye();
    }

    std::string outbase;
    if (fname_in && !fname_out_arg) fname_out_arg = fname_in;
    if (fname_out_arg) outbase = strip_extension(fname_out_arg);
    const char *fname_out = fname_out_arg ? outbase.c_str() : nullptr;

    if (incnt && inrep) die("only one of -C and -R can be specified");
    if (incnt) {
        run_cntmode(fname_in, fname_out, dobp, doau, threshold, opt, pfmin);
        printf("\n# exit normally\n");
        return 0;
    }

    if (multi && !inrep) {  // rmt mode only (src/consel.c:439-440)
        const MultiArgs A{fname_in, fname_out, fname_ass, fname_vt, fname_pa, dobp, doba, doau, outcnt, smoothcnt, nosort,
                          fastrep, cm_arg, gpu, threshold, bapcoef, congbpadd, pfmin};
        run_multimode(A, opt);
        printf("\n# exit normally\n");
        return 0;
    }

    std::string used;
    std::vector<double> datvec, rr;
    std::vector<int> bb;
    std::vector<std::vector<double>> repmats;
    int mm = 0, kk = 0, cm = 0;
    Assoc ass;
    std::vector<double> obsvec, datvec00, pdmvec;
    std::vector<int> orderv;
    if (inrep) {
        // ---------------- read .rep (do_repmode, src/consel.c:1195-1226): the statistic rows
        // land in `repmats`, cm x B_s each
        doba = domc = false;
        std::vector<int> cnt;
        if (fname_in) {
            FILE *fp = open_file(fname_in, ".rep", "rb", &used);
            printf("\n# reading %s", used.c_str());
            orderv = read_bivec(fp);
            obsvec = read_bvec(fp);
            cm = (int)orderv.size();
            if ((int)obsvec.size() != cm) die("size mismatch in vec");
            rr = read_bvec(fp);
            kk = (int)rr.size();
            bb = read_bivec(fp, kk);
            if (read_bint(fp) != kk) die("wrong number of matrices");
            repmats.resize(kk);
            for (int i = 0; i < kk; i++) {
                int m = cm, b = bb[i];
                repmats[i] = read_bmat(fp, &m, &b);
            }
            fclose(fp);
        } else {
            printf("\n# reading from stdin");
            orderv = read_ivec(stdin);
            obsvec = read_vec(stdin);
            cm = (int)orderv.size();
            if ((int)obsvec.size() != cm) die("size mismatch in vec");
            rr = read_vec(stdin);
            kk = (int)rr.size();
            bb = read_ivec(stdin, kk);
            if (read_int(stdin) != kk) die("wrong number of matrices");
            repmats.resize(kk);
            for (int i = 0; i < kk; i++) {
                int m = cm, b = bb[i];
                repmats[i] = read_mat(stdin, &m, &b);
                putchar('.');
                fflush(stdout);
            }
        }
        printf("\n# K:%d", kk);
        printf("\n# R:");
        for (int i = 0; i < kk; i++) printf("%g ", rr[i]);
        printf("\n# B:");
        for (int i = 0; i < kk; i++) printf("%d ", bb[i]);
        printf("\n# CM:%d", cm);
        if (kk < 1 || cm < 1) { printf("\n# exit normally\n"); return 0; }
    } else {
    // ---------------- read .rmt (src/consel.c:806-834)
    if (fname_in) {
        FILE *fp = open_file(fname_in, ".rmt", "rb", &used);
        printf("\n# reading %s", used.c_str());
        datvec = read_bvec(fp);
        mm = (int)datvec.size();
        rr = read_bvec(fp);
        kk = (int)rr.size();
        bb = read_bivec(fp, kk);
        if (read_bint(fp) != kk) die("wrong size in rmt");
        repmats.resize(kk);
        for (int i = 0; i < kk; i++) {
            int m = mm, b = bb[i];
            repmats[i] = read_bmat(fp, &m, &b);
            putchar('.');
            fflush(stdout);
        }
        fclose(fp);
    } else {
        printf("\n# reading from stdin");
        datvec = read_vec(stdin);
        mm = (int)datvec.size();
        rr = read_vec(stdin);
        kk = (int)rr.size();
        bb = read_ivec(stdin, kk);
        if (read_int(stdin) != kk) die("wrong size in rmt");
        repmats.resize(kk);
        for (int i = 0; i < kk; i++) {
            int m = mm, b = bb[i];
            repmats[i] = read_mat(stdin, &m, &b);
            putchar('.');
            fflush(stdout);
        }
    }
    printf("\n# K:%d", kk);
    printf("\n# R:");
    for (int i = 0; i < kk; i++) printf("%g ", rr[i]);
    printf("\n# B:");
    for (int i = 0; i < kk; i++) printf("%d ", bb[i]);
    printf("\n# M:%d", mm);
    if (kk < 1) { printf("\n# exit normally\n"); return 0; }
    if (mm < 2) {
        warn("M should be at least two");
        printf("\n# exit normally\n");
        return 0;
    }

    // ---------------- AIC correction (read_pdm / do_pdm, src/consel.c:675-744): a constant per tree
    // added to the observed log-likelihoods and to every replicate, on the host before the
    // upload (one rounded add per element, as in the reference)
    datvec00 = datvec;
    if (fname_pdm) {
        FILE *fp = open_file(fname_pdm, ".vt", "r", &used);
        printf("\n# reading %s", used.c_str());
        pdmvec = read_vec(fp);
        fclose(fp);
        printf("\n# pdm:");
        for (double v : pdmvec) printf("%g ", v);
        printf("\n# aictype:%d", aictype);
    }
    if (!pdmvec.empty()) {
        if ((int)pdmvec.size() != mm) die("size of pdm should be M");
        if (aictype != 0) {
            for (int i = 0; i < mm; i++) {
                double x = 0.0;  // an aictype outside 1..4 leaves x as it was in the reference (uninitialised); 0 here
                if (aictype == 1 || aictype == 4) x = -1.0 * pdmvec[i];
                else if (aictype == 2 || aictype == 3) x = -0.5 * pdmvec[i];
                datvec[i] += x;
            }
            for (int k = 0; k < kk; k++)
                for (int i = 0; i < mm; i++) {
                    double x = 0.0;
                    if (aictype == 1) x = -1.0 * pdmvec[i] / rr[k];
                    else if (aictype == 2) x = -0.5 * pdmvec[i];
                    else if (aictype == 3) x = -0.5 * pdmvec[i] * (1.0 - 1.0 / rr[k]);
                    else if (aictype == 4) x = -1.0 * pdmvec[i];
                    double *row = &repmats[k][(size_t)i * bb[k]];
                    for (int j = 0; j < bb[k]; j++) row[j] += x;
                }
        }
    }

    cm = read_assoc_and_order(fname_ass, datvec, cm_arg, nosort, &ass, &obsvec, &orderv);
    }  // rmt mode

    // ---------------- alphas (read_alpha, src/consel.c:543-560)
    std::vector<double> alphavec = {0.05, 0.10, 0.5, 0.90, 0.95};
    if (fname_vt) {
        FILE *fp = open_file(fname_vt, ".vt", "r", &used);
        printf("\n# reading %s", used.c_str());
        alphavec = read_vec(fp);
        fclose(fp);
    }
    const int nalpha = (int)alphavec.size();
    std::vector<double> rr1 = {0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4};  // rr0, src/consel.c:162-163
    if (fastrep && !inrep) {  // read_scale, src/consel.c:562-582
        if (fname_pa) {
            FILE *fp = open_file(fname_pa, ".pa", "r", &used);
            printf("\n# reading %s", used.c_str());
            rr1 = read_vec(fp);
            fclose(fp);
        }
        printf("\n# RESCALING");
        printf("\n# K:%d", (int)rr1.size());
        printf("\n# R:");
        for (double r : rr1) printf("%g ", r);
    }

    // ---------------- Bayesian posterior (do_batest, src/consel.c:1311-1340)
    std::vector<double> bapvec(cm, 0.0);
    if (doba) {
        std::vector<double> w(mm);
        double x = -kHuge, y = 0.0;
        if ((int)pdmvec.size() != mm) {
            for (int i = 0; i < mm; i++)
                if (datvec[i] > x) x = datvec[i];
            for (int i = 0; i < mm; i++) y += w[i] = exp(bapcoef * (datvec[i] - x));
        } else {  // BIC correction: P = const*exp{bapcoef*logL - 0.5*pdm*log(n)} on the uncorrected logL (:1325-1332)
            printf("\n# ppn:%g", bapn);
            for (int i = 0; i < mm; i++) w[i] = bapcoef * datvec00[i] - 0.5 * pdmvec[i] * log(bapn);
            for (int i = 0; i < mm; i++)
                if (w[i] > x) x = w[i];
            for (int i = 0; i < mm; i++) y += w[i] = exp(w[i] - x);
        }
        for (int i = 0; i < mm; i++) w[i] = w[i] / y;
        for (int i = 0; i < cm; i++) {
            double s = 0.0;
            for (int q = 0; q < ass.len(i); q++) s += w[ass.at(i)[q]];
            bapvec[i] = s;
        }
    }

    // ---------------- device: upload the replicate blocks
    CG(cgpu_init(gpu));
    // one process per GPU (RANK / WORLD_SIZE): every rank uploads its share of the replicate columns of every
    // scale; KH/SH tests and AU fits are split over items, count queries are NCCL collectives (pipeline.cpp);
    // rank 0 writes the files.  The modes that need whole rows on one device stay single-GPU.
    CG(cgpu_comm_init_env());
    int g_rank = 0, g_world = 1;
    cgpu_comm_info(&g_rank, &g_world);
    if (g_world > 1 && (inrep || fastrep || outrep || quickmc || smoothcnt || debugmode > 0))
        die("-R, -f, -r, --quick_mctest, smoothed -c and -d need a single GPU (WORLD_SIZE=%d)", g_world);
    cgpu_reps *reps = nullptr;
    {
        std::vector<const double *> ptrs(kk);
        for (int i = 0; i < kk; i++) ptrs[i] = repmats[i].data();
        if (inrep) CG(cgpu_stats_upload(cm, kk, bb.data(), ptrs.data(), &reps));
        else CG(cgpu_reps_upload(mm, kk, bb.data(), ptrs.data(), &reps));
    }
    int i0 = 0;  // scale closest to r = 1 (find_i0, src/consel.c:1778-1789)
    for (int i = 1; i < kk; i++)
        if (fabs(rr[i] - 1.0) < fabs(rr[i0] - 1.0)) i0 = i;
    const int nb0 = bb[i0];
    auto seval = [](double p, double nb) { return sqrt(p * (1.0 - p) / nb); };  // getseval :2151-2158

    // ---------------- KH / SH tests (do_mctest)
    std::vector<double> khp(cm), mcp(cm), khwp(cm), mcwp(cm);
    std::vector<double> d_pv, d_se, d_ax, d_ci;  // multi-rank results (pipeline::consel_run)
    std::vector<int> d_ret;
    std::vector<cgpu_rowsum> d_q0;
    if (g_world > 1) {
        d_pv.assign((size_t)cm * 8, 0.0); d_se.assign((size_t)cm * 8, 0.0); d_ax.assign((size_t)cm * 7, 0.0);
        d_ci.assign((size_t)6 * cm * std::max(nalpha, 1), 0.0); d_ret.assign(cm, 0); d_q0.resize((size_t)cm * kk);
        if (kk < 2) doau = false;
        printf("\n# MC/BP/AU tests on %d GPUs", g_world);
        fflush(stdout);
        if (pipeline::consel_run(reps, mm, datvec.data(), kk, rr.data(), bb.data(), cm, ass.ptr.data(), ass.idx.data(), threshold,
                                 varadd, opt, domc ? 1 : 0, doau ? 1 : 0, nalpha, alphavec.data(), d_pv.data(), d_se.data(), nullptr,
                                 d_ax.data(), doau ? d_ci.data() : nullptr, d_ret.data(), nullptr, d_q0.data()) != 0)
            die("%s", cgpu_last_error());
        for (int i = 0; i < cm; i++) {
            khp[i] = d_pv[8 * (size_t)i + 2]; mcp[i] = d_pv[8 * (size_t)i + 3];
            khwp[i] = d_pv[8 * (size_t)i + 4]; mcwp[i] = d_pv[8 * (size_t)i + 5];
        }
    }
    if (domc && g_world == 1) {
        if (quickmc) CG(cgpu_set_option("mc_smooth", 0));  // sw_pvaldist=0: plain counts, no smoothing
        printf("\n# MC-TEST STARTS");
        printf("\n# centering the replicates");
        CG(cgpu_mctest(reps, i0, datvec.data(), varadd, cm, ass.ptr.data(), ass.idx.data(), khp.data(), mcp.data(),
                       khwp.data(), mcwp.data()));
        printf("\n# calculating kh-pvalue"); dots(cm);
        printf("\n# calculating mc-pvalue"); dots(cm);
        printf("\n# calculating the variances");
        small_variance_warnings(inrep ? nullptr : repmats[i0].data(), mm, nb0, varadd);
        printf("\n# calculating weighted kh-pvalue"); dots(cm);
        printf("\n# calculating weighted mc-pvalue"); dots(cm);
        printf("\n# MC-TEST DONE");
    }

    // ---------------- statistics of the replicates (genrep)
    std::vector<double> npvec(cm), pvvec(cm), sevec(cm), pv0vec(cm), se0vec(cm), rssvec(cm), dfvec(cm), pfvec(cm),
        thvec(cm), beta(3 * (size_t)cm, 0.0);
    std::vector<double> cimat, csmat, eimat, ci0mat, cs0mat, ei0mat, cntmat;
    std::vector<cgpu_rowsum> r0;
    if (g_world > 1) {
        r0 = d_q0;
    } else if (doau || dobp || outrep || outcnt) {
        if (!inrep && fastrep) {  // genrep's rescaling branch (src/consel.c:757-771)
            printf("\n# calculate replicates of the statistics");
            const int kk1 = (int)rr1.size();
            std::vector<double> rs(kk1);
            for (int i = 0; i < kk1; i++) rs[i] = sqrt(rr[i0] / rr1[i]);
            cgpu_reps *scaled = nullptr;
            CG(cgpu_stats_build_rescaled(reps, i0, kk1, rs.data(), cm, ass.ptr.data(), ass.idx.data(), &scaled));
            dots(kk1);
            cgpu_reps_free(reps);  // the KH/SH tests are done; everything below reads the statistics only
            reps = scaled;
            bb.assign(kk1, nb0);
            rr = rr1;
            kk = kk1;
            i0 = 0;  // find_i0 again on the new scales (do_bptest :1293, do_bootrep :1463)
            for (int i = 1; i < kk; i++)
                if (fabs(rr[i] - 1.0) < fabs(rr[i0] - 1.0)) i0 = i;
        } else if (!inrep) {
            printf("\n# calculate replicates of the statistics");
            CG(cgpu_stats_build(reps, cm, ass.ptr.data(), ass.idx.data()));
            dots(kk);
        }
        // one query at the region threshold serves BP, the .cnt file and the row extremes
        std::vector<double> thr(cm, threshold);
        r0.resize((size_t)cm * kk);
        CG(cgpu_count_query(reps, 1, thr.data(), r0.data()));
    }
    // do_rmtmode drops the AU test for a single scale (src/consel.c:868); do_repmode has no such guard (:1231-1232):
    // `consel -R` on a one-scale .rep runs the (degenerate) fit and writes .ci / .cnt
    if (kk < 2 && !inrep) doau = false;

    if (dobp) {  // do_bptest
        printf("\n# BP-TEST STARTS");
        for (int jx = 0; jx < cm; jx++) {
            const cgpu_rowsum &q0 = r0[(size_t)jx * kk + i0];
            double x = quickmc ? q0.cnt : cgpu_smooth_count(&q0, nb0, threshold);  // pvaldist honours sw_pvaldist (:1298)
            npvec[jx] = 1.0 - (1.0 - x / nb0);
        }
        printf(" - DONE");
    }

    std::vector<double> bbd(kk);
    for (int k = 0; k < kk; k++) bbd[k] = (double)bb[k];

    if (doau && g_world > 1) {
        const size_t n2 = (size_t)cm * nalpha;
        for (int i = 0; i < cm; i++) {
            pvvec[i] = d_pv[8 * (size_t)i + 6]; sevec[i] = d_se[8 * (size_t)i + 6];
            pv0vec[i] = d_pv[8 * (size_t)i + 7]; se0vec[i] = d_se[8 * (size_t)i + 7];
            const double *a = &d_ax[7 * (size_t)i];
            pfvec[i] = a[0]; rssvec[i] = a[1]; dfvec[i] = a[2]; thvec[i] = a[5];
            if (d_ret[i]) {
                warn("regression degenerated: df[%d]=%g", i + 1, dfvec[i]);
            } else {
                if (pfvec[i] < 0.01) warn("theory does not fit well: pfit[%d]=%.4f", i + 1, pfvec[i]);
                beta[3 * (size_t)i] = a[3]; beta[3 * (size_t)i + 1] = a[4]; beta[3 * (size_t)i + 2] = a[6];
            }
        }
        cimat.assign(d_ci.begin(), d_ci.begin() + n2); csmat.assign(d_ci.begin() + n2, d_ci.begin() + 2 * n2);
        eimat.assign(d_ci.begin() + 2 * n2, d_ci.begin() + 3 * n2); ci0mat.assign(d_ci.begin() + 3 * n2, d_ci.begin() + 4 * n2);
        cs0mat.assign(d_ci.begin() + 4 * n2, d_ci.begin() + 5 * n2); ei0mat.assign(d_ci.begin() + 5 * n2, d_ci.begin() + 6 * n2);
        printf("\n# AU-TEST DONE");
    } else if (doau) {  // do_bootrep
        printf("\n# AU-TEST STARTS");
        printf("\n# counting the replicates (sort-free row summaries on the device)");
        dots(kk);
        printf("\n# calculating approximately unbiased p-values by ");
        printf(opt.fitmode == aufit::FIT_WLS ? "WLS (very fast)" : "MLE (fast)");
        printf(" fitting");
        fflush(stdout);

        // starting point of every item: the median of the r~1 row (src/consel.c:2765)
        std::vector<int> rank(cm, nb0 / 2);
        std::vector<double> med(cm);
        CG(cgpu_order_stat(reps, i0, rank.data(), med.data()));
        audriver::Input ain;
        ain.cm = cm; ain.kk = kk; ain.rr = rr.data(); ain.bb = bb.data(); ain.i0 = i0; ain.threshold = threshold;
        ain.median = med.data(); ain.r0 = r0.data(); ain.nalpha = nalpha; ain.alpha = alphavec.data();
        audriver::Output ao;
        auto query = [](void *ctx, int nq, const double *thr, cgpu_rowsum *out) {
            return cgpu_count_query(static_cast<const cgpu_reps *>(ctx), nq, thr, out);
        };
        auto query_rows = [](void *ctx, int row0, int nrows, int nq, const double *thr, cgpu_rowsum *out) {
            return cgpu_count_query_rows(static_cast<const cgpu_reps *>(ctx), row0, nrows, nq, thr, out);
        };
        if (audriver::run(ain, opt, query, reps, &ao, query_rows) != 0) die("%s", cgpu_last_error());
        for (int i = 0; i < cm; i++) {
            dprint(1, "\n# rank=%d item=%d", i + 1, orderv[i] + 1);
            dprint(1, "\n# ret=%d", ao.ret[i]);
            pvvec[i] = ao.pv[i]; sevec[i] = ao.se[i]; pv0vec[i] = ao.pv0[i]; se0vec[i] = ao.se0[i];
            rssvec[i] = ao.rss[i]; dfvec[i] = ao.df[i]; thvec[i] = ao.th[i]; pfvec[i] = ao.pf[i];
            if (ao.ret[i]) {
                warn("regression degenerated: df[%d]=%g", i + 1, dfvec[i]);
            } else {
                if (pfvec[i] < 0.01) warn("theory does not fit well: pfit[%d]=%.4f", i + 1, pfvec[i]);
                for (int c = 0; c < 3; c++) beta[3 * (size_t)i + c] = ao.beta[3 * (size_t)i + c];
            }
            putchar('.');
        }
        fflush(stdout);
        printf("\n# time elapsed for AU test is t=%g sec", ao.t_fit_sec);

        // ---- confidence limits (invtpval for au and bp)
        printf("\n# ALPHA:");
        for (int i = 0; i < nalpha; i++) printf("%g ", alphavec[i]);
        printf("\n# calculating confidence intervals");
        cimat = ao.ci; csmat = ao.cs; eimat = ao.ei; ci0mat = ao.ci0; cs0mat = ao.cs0; ei0mat = ao.ei0;
        for (int jx = 0; jx < cm; jx++) {
            for (int a = 0; a < nalpha; a++)
                for (int which = 0; which < 2; which++) {
                    const size_t e = 2 * ((size_t)jx * nalpha + a) + which;
                    dprint(1, "\n# invtpval item=%d alpha=%g ", jx + 1, alphavec[a]);
                    if (ao.ci_warn[e] == 1) warn("degenerecy in pval df=%g", ao.ci_warnval[e]);
                    if (ao.ci_warn[e] == 2) warn("no convergence in ci: ei=%6.3f", ao.ci_warnval[e]);
                    if (ao.ci_ret[e]) printf(" %s item=%2d, alpha=%4.2f", which ? "bp" : "au", jx + 1, alphavec[a]);
                }
            putchar('.');
        }
        fflush(stdout);
        printf("\n# AU-TEST DONE");
    }

    if (inrep && !doau) outrep = outcnt = false;  // do_repmode writes them only after the AU test (:1236-1240)
    // cntmat exists only when the AU test ran (do_bootrep allocates it, src/consel.c:1464, 1502-1507): without it
    // -- a single scale (:868), --no_au -- write_cnt2 (:1770) emits the CNT header alone
    if (outcnt && doau) {
        cntmat.resize((size_t)cm * kk);
        for (int i = 0; i < cm; i++)
            for (int k = 0; k < kk; k++) {
                const cgpu_rowsum &q = r0[(size_t)i * kk + k];
                cntmat[(size_t)i * kk + k] = smoothcnt ? cgpu_smooth_count(&q, bb[k], threshold) : q.cnt;
            }
    }

    if (g_world > 1 && g_rank > 0) {  // rank 0 writes the files
        cgpu_reps_free(reps);
        cgpu_comm_destroy();
        cgpu_shutdown();
        return 0;
    }

    // ---------------- outputs (write_pv :1611-1683, write_ci :1685-1711, write_rep, write_cnt)
    if (dobp || domc || doau || doba) {
        FILE *fp = stdout;
        if (fname_out) {
            fp = open_file(fname_out, ".pv", "w", &used);
            printf("\n# writing %s", used.c_str());
        } else {
            printf("\n# writing to stdout\n");
        }
        const int outbit = (dobp ? 1 : 0) + (doba ? 2 : 0) + (domc ? 4 : 0) + (doau ? 8 : 0);
        const int pvnum = (dobp ? 1 : 0) + (doba ? 1 : 0) + (domc ? 4 : 0) + (doau ? 2 : 0);
        const int auxnum = doau ? 7 : 0;
        fprintf(fp, "\n# ID:\n%d\n", 1);
        fprintf(fp, "\n# ITEM:\n"); write_ivec(fp, orderv.data(), cm);
        fprintf(fp, "\n# STAT:\n"); write_vec(fp, obsvec.data(), cm);
        fprintf(fp, "\n# BIT:\n%d\n", outbit);
        std::vector<double> pvmat((size_t)cm * pvnum), semat((size_t)cm * pvnum), auxmat((size_t)cm * auxnum);
        for (int i = 0; i < cm; i++) {
            int j = 0;
            auto put = [&](double p, double s) {
                pvmat[(size_t)i * pvnum + j] = p;
                semat[(size_t)i * pvnum + j] = s;
                j++;
            };
            if (dobp) put(npvec[i], seval(npvec[i], (double)nb0));
            if (doba) put(bapvec[i], 0.0);
            if (domc) {
                put(khp[i], seval(khp[i], (double)nb0));
                put(mcp[i], seval(mcp[i], (double)nb0));
                put(khwp[i], seval(khwp[i], (double)nb0));
                put(mcwp[i], seval(mcwp[i], (double)nb0));
            }
            if (doau) {
                put(pvvec[i], sevec[i]);
                put(pv0vec[i], se0vec[i]);
                double *ax = &auxmat[(size_t)i * auxnum];
                ax[0] = pfvec[i]; ax[1] = rssvec[i]; ax[2] = dfvec[i];
                ax[3] = beta[3 * (size_t)i]; ax[4] = beta[3 * (size_t)i + 1];
                ax[5] = thvec[i]; ax[6] = beta[3 * (size_t)i + 2];
            }
        }
        fprintf(fp, "\n# PV:\n"); write_mat(fp, pvmat.data(), cm, pvnum);
        fprintf(fp, "\n# SE:\n"); write_mat(fp, semat.data(), cm, pvnum);
        fprintf(fp, "\n# AX:\n"); write_mat(fp, auxnum ? auxmat.data() : pvmat.data(), cm, auxnum);
        if (fname_out) fclose(fp);
    }
    if (doau) {
        FILE *fp = stdout;
        if (fname_out) {
            fp = open_file(fname_out, ".ci", "w", &used);
            printf("\n# writing %s", used.c_str());
        } else {
            printf("\n# writing to stdout\n");
        }
        fprintf(fp, "\n# ITEM:\n"); write_ivec(fp, orderv.data(), cm);
        fprintf(fp, "\n# STAT:\n"); write_vec(fp, obsvec.data(), cm);
        fprintf(fp, "\n# ALPHA:\n"); write_vec(fp, alphavec.data(), nalpha);
        fprintf(fp, "\n# CI:\n"); write_mat(fp, cimat.data(), cm, nalpha);
        fprintf(fp, "\n# SE:\n"); write_mat(fp, csmat.data(), cm, nalpha);
        fprintf(fp, "\n# EI:\n"); write_mat(fp, eimat.data(), cm, nalpha);
        fprintf(fp, "\n# CI0:\n"); write_mat(fp, ci0mat.data(), cm, nalpha);
        fprintf(fp, "\n# SE0:\n"); write_mat(fp, cs0mat.data(), cm, nalpha);
        fprintf(fp, "\n# EI0:\n"); write_mat(fp, ei0mat.data(), cm, nalpha);
        if (fname_out) fclose(fp);
    }
    if (outrep) {
        std::vector<std::vector<double>> st(kk);
        std::vector<double *> ptrs(kk);
        for (int k = 0; k < kk; k++) {
            st[k].resize((size_t)cm * bb[k]);
            ptrs[k] = st[k].data();
        }
        // The reference writes the rows in whatever state its tests left them: do_bootrep sorts every row of every
        // scale in place (src/consel.c:1466-1471); without the AU test only do_bptest touches them -- its pvaldist
        // sorts the rows of scale i0 alone (:1298, cntdist mode 4 :1870), and not at all under --quick_mctest
        // (sw_pvaldist = 0, :323).
        CG(cgpu_stats_download(reps, doau ? 1 : 0, ptrs.data()));
        if (!doau && dobp && !quickmc)
            for (int j = 0; j < cm; j++) std::sort(st[i0].begin() + (size_t)j * bb[i0], st[i0].begin() + (size_t)(j + 1) * bb[i0]);
        const std::string repbase = (inrep && fname_in) ? std::string(fname_in) : outbase;  // fname_rep (:438, :443)
        if (fname_out) {
            FILE *fp = open_file(repbase.c_str(), ".rep", "wb", &used);
            printf("\n# writing %s", used.c_str());
            write_bivec(fp, orderv.data(), cm);
            write_bvec(fp, obsvec.data(), cm);
            write_bvec(fp, rr.data(), kk);
            write_bivec(fp, bb.data(), kk);
            write_bint(fp, kk);
            for (int k = 0; k < kk; k++) {
                write_bmat(fp, st[k].data(), cm, bb[k]);
                putchar('.');
            }
            fclose(fp);
        } else {
            printf("\n# writing to stdout\n");
            write_ivec(stdout, orderv.data(), cm);
            write_vec(stdout, obsvec.data(), cm);
            write_vec(stdout, rr.data(), kk);
            write_ivec(stdout, bb.data(), kk);
            printf("%d\n", kk);
            for (int k = 0; k < kk; k++) write_mat(stdout, st[k].data(), cm, bb[k]);
        }
    }
    if (outcnt) {
        FILE *fp = stdout;
        if (fname_out) {
            fp = open_file(fname_out, ".cnt", "w", &used);
            printf("\n# writing %s", used.c_str());
        } else {
            printf("\n# writing to stdout\n");
        }
        std::vector<double> rrmat((size_t)cm * kk), bbmat((size_t)cm * kk);
        for (int i = 0; i < cm; i++)
            for (int k = 0; k < kk; k++) {
                rrmat[(size_t)i * kk + k] = rr[k];
                bbmat[(size_t)i * kk + k] = (double)bb[k];
            }
        fprintf(fp, "\n# ITEM:\n"); write_ivec(fp, orderv.data(), cm);
        fprintf(fp, "\n# STAT:\n"); write_vec(fp, obsvec.data(), cm);
        fprintf(fp, "\n# R:\n"); write_mat(fp, rrmat.data(), cm, kk);
        fprintf(fp, "\n# B:\n"); write_mat(fp, bbmat.data(), cm, kk);
        fprintf(fp, "\n# CNT:\n"); write_mat(fp, cntmat.empty() ? nullptr : cntmat.data(), cm, kk);
        if (fname_out) fclose(fp);
    }
    cgpu_reps_free(reps);
    printf("\n# exit normally\n");
    cgpu_shutdown();
    return 0;
}
