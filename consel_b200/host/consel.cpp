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
// Not carried by this build (error at the command line): -g multi-gene mode, --chi* spherical fit, --pdm AIC correction,
// --quick_mctest.  They are off the hot path (SURVEY 8f "next").
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/consel_gpu.h"
#include "assoc.h"
#include "audriver.h"
#include "aufit.h"
#include "refio.h"
#include "specfun.h"

using namespace refio;

static void byebye() { die("error in command line"); }
static void dots(int n) {
    for (int i = 0; i < n; i++) putchar('.');
    fflush(stdout);
}
#define CG(call)                                           \
    do {                                                   \
        if ((call) != CGPU_OK) die("%s", cgpu_last_error()); \
    } while (0)



// ---------------------------------------------------------------- cnt mode (consel -C)
// do_cntmode (src/consel.c:1244-1283) + do_bootcnt (:1540-1592): BP and the AU fit from the
// count matrix of a `.cnt` file.  No replicates, no device: the reference's path is the
// same host arithmetic.  rcalpval (:2703-2737) with its default pfmin = 0 makes exactly one
// calcpval call, on the scales from the smallest r >= rmin that is below 1 upwards, and
// reports "degenerated" without fitting when there is no scale in [rmin, 1).
static void run_cntmode(const char *fname_in, const char *fname_out, bool dobp, bool doau, double threshold,
                        const aufit::Options &opt) {
    std::string used;
    FILE *fp = stdin;
    if (fname_in) {
        fp = open_file(fname_in, ".cnt", "r", &used);
        printf("\n# reading %s", used.c_str());
    } else {
        printf("\n# reading from stdin");
    }
    std::vector<int> orderv = read_ivec(fp);
    int cm = (int)orderv.size(), kk = 0;
    std::vector<double> obsvec = read_vec(fp, cm);
    std::vector<double> rrmat = read_mat(fp, &cm, &kk), bbmat = read_mat(fp, &cm, &kk), cntmat = read_mat(fp, &cm, &kk);
    if (fname_in) fclose(fp);
    printf("\n# CM:%d", cm);
    printf("\n# K:%d", kk);

    std::vector<double> npvec(cm), nsvec(cm), pvvec(cm), sevec(cm), pv0vec(cm), se0vec(cm), rssvec(cm), dfvec(cm),
        pfvec(cm), thvec(cm), beta(3 * (size_t)cm, 0.0);
    if (dobp) {
        for (int j = 0; j < cm; j++) {
            const double *rr = &rrmat[(size_t)j * kk];
            int i0 = 0;
            for (int i = 1; i < kk; i++)
                if (fabs(rr[i] - 1.0) < fabs(rr[i0] - 1.0)) i0 = i;
            const double b = bbmat[(size_t)j * kk + i0];
            npvec[j] = cntmat[(size_t)j * kk + i0] / b;
            nsvec[j] = sqrt(npvec[j] * (1.0 - npvec[j]) / b);
        }
    }
    if (doau) {
        printf("\n# AU-TEST STARTS");
        printf("\n# calculating approximately unbiased p-values by ");
        printf(opt.fitmode == aufit::FIT_WLS ? "WLS (very fast)" : "MLE (fast)");
        printf(" fitting");
        fflush(stdout);
        const double t0 = cpu_seconds();
        for (int i = 0; i < cm; i++) {
            const double *rr = &rrmat[(size_t)i * kk];
            std::vector<double> rs(rr, rr + kk);
            std::sort(rs.begin(), rs.end());
            int i1 = 0;
            while (i1 < kk && rs[i1] < opt.rrmin) i1++;
            int i2 = i1;
            while (i2 < kk && rs[i2] < 1.0) i2++;
            aufit::Fit f;
            int j = 1;
            if (i1 < i2) {
                f = aufit::calcpval(&cntmat[(size_t)i * kk], rr, &bbmat[(size_t)i * kk], kk, rs[i1], opt.rrmax, opt.kappa, opt);
                j = f.rc;
            }
            pvvec[i] = f.pv; sevec[i] = f.se; pv0vec[i] = f.pv0;
            if (cm > 1) se0vec[1] = f.se0;  // the reference passes `se0vec+1` for every item (:1567)
            rssvec[i] = f.rss; dfvec[i] = f.df;
            thvec[i] = threshold;  // unused in cnt mode
            dprint(1, "\n# rank=%d item=%d", i + 1, orderv[i] + 1);
            dprint(1, "\n# ret=%d", j);
            if (j) {
                pfvec[i] = 0.0;
                warn("regression degenerated: df[%d]=%g", i + 1, dfvec[i]);
            } else {
                pfvec[i] = specfun::pochisq(rssvec[i], (int)dfvec[i]);
                if (pfvec[i] < 0.01) warn("theory does not fit well: pfit[%d]=%.4f", i + 1, pfvec[i]);
                for (int c = 0; c < 3; c++) beta[3 * (size_t)i + c] = f.beta[c];
            }
            putchar('.');
            fflush(stdout);
        }
        printf("\n# time elapsed for AU test is t=%g sec", cpu_seconds() - t0);
        printf("\n# AU-TEST DONE");
    }
    if (!(dobp || doau)) return;
    // write_pv(sw_dobp,0,0,sw_doau) (src/consel.c:1611-1683)
    fp = stdout;
    if (fname_out) {
        fp = open_file(fname_out, ".pv", "w", &used);
        printf("\n# writing %s", used.c_str());
    } else {
        printf("\n# writing to stdout\n");
    }
    const int outbit = (dobp ? 1 : 0) + (doau ? 8 : 0), pvnum = (dobp ? 1 : 0) + (doau ? 2 : 0), auxnum = doau ? 7 : 0;
    fprintf(fp, "\n# ID:\n%d\n", 1);
    fprintf(fp, "\n# ITEM:\n"); write_ivec(fp, orderv.data(), cm);
    fprintf(fp, "\n# STAT:\n"); write_vec(fp, obsvec.data(), cm);
    fprintf(fp, "\n# BIT:\n%d\n", outbit);
    std::vector<double> pvmat((size_t)cm * pvnum), semat((size_t)cm * pvnum), auxmat((size_t)cm * auxnum);
    for (int i = 0; i < cm; i++) {
        int j = 0;
        if (dobp) { pvmat[(size_t)i * pvnum + j] = npvec[i]; semat[(size_t)i * pvnum + j] = nsvec[i]; j++; }
        if (doau) {
            pvmat[(size_t)i * pvnum + j] = pvvec[i]; semat[(size_t)i * pvnum + j] = sevec[i]; j++;
            pvmat[(size_t)i * pvnum + j] = pv0vec[i]; semat[(size_t)i * pvnum + j] = se0vec[i]; j++;
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

int main(int argc, char **argv) {
    const char *fname_in = nullptr, *fname_out_arg = nullptr, *fname_ass = nullptr, *fname_vt = nullptr, *fname_pa = nullptr;
    bool fastrep = false;
    bool outrep = false, outcnt = false, smoothcnt = false, nosort = false, inrep = false, incnt = false;
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
        else if (a == "-g" || a == "--pdm" ||
                 a == "--quick_mctest" || a.rfind("--chi", 0) == 0 || a == "--aictype" || a == "--ppn" ||
                 a == "--pmin" || a == "--congbp")
            die("switch %s is not supported by the B200 build of consel (rmt, rep and cnt modes only)", a.c_str());
        else byebye();
    }

    std::string outbase;
    if (fname_in && !fname_out_arg) fname_out_arg = fname_in;
    if (fname_out_arg) outbase = strip_extension(fname_out_arg);
    const char *fname_out = fname_out_arg ? outbase.c_str() : nullptr;

    if (incnt && inrep) die("only one of -C and -R can be specified");
    if (incnt) {
        run_cntmode(fname_in, fname_out, dobp, doau, threshold, opt);
        printf("\n# exit normally\n");
        return 0;
    }

    std::string used;
    std::vector<double> datvec, rr;
    std::vector<int> bb;
    std::vector<std::vector<double>> repmats;
    int mm = 0, kk = 0, cm = 0;
    Assoc ass;
    std::vector<double> obsvec;
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
        double x = -kHuge;
        for (int i = 0; i < mm; i++)
            if (datvec[i] > x) x = datvec[i];
        std::vector<double> w(mm);
        double y = 0.0;
        for (int i = 0; i < mm; i++) y += w[i] = exp(bapcoef * (datvec[i] - x));
        for (int i = 0; i < mm; i++) w[i] = w[i] / y;
        for (int i = 0; i < cm; i++) {
            double s = 0.0;
            for (int q = 0; q < ass.len(i); q++) s += w[ass.at(i)[q]];
            bapvec[i] = s;
        }
    }

    // ---------------- device: upload the replicate blocks
    CG(cgpu_init(gpu));
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
    if (domc) {
        printf("\n# MC-TEST STARTS");
        printf("\n# centering the replicates");
        CG(cgpu_mctest(reps, i0, datvec.data(), varadd, cm, ass.ptr.data(), ass.idx.data(), khp.data(), mcp.data(),
                       khwp.data(), mcwp.data()));
        printf("\n# calculating kh-pvalue"); dots(cm);
        printf("\n# calculating mc-pvalue"); dots(cm);
        printf("\n# calculating the variances"); dots(mm);
        printf("\n# calculating weighted kh-pvalue"); dots(cm);
        printf("\n# calculating weighted mc-pvalue"); dots(cm);
        printf("\n# MC-TEST DONE");
    }

    // ---------------- statistics of the replicates (genrep)
    std::vector<double> npvec(cm), pvvec(cm), sevec(cm), pv0vec(cm), se0vec(cm), rssvec(cm), dfvec(cm), pfvec(cm),
        thvec(cm), beta(3 * (size_t)cm, 0.0);
    std::vector<double> cimat, csmat, eimat, ci0mat, cs0mat, ei0mat, cntmat;
    std::vector<cgpu_rowsum> r0;
    if (doau || dobp || outrep || outcnt) {
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
    if (kk < 2) doau = false;

    if (dobp) {  // do_bptest
        printf("\n# BP-TEST STARTS");
        for (int jx = 0; jx < cm; jx++) {
            double x = cgpu_smooth_count(&r0[(size_t)jx * kk + i0], nb0, threshold);
            npvec[jx] = 1.0 - (1.0 - x / nb0);
        }
        printf(" - DONE");
    }

    std::vector<double> bbd(kk);
    for (int k = 0; k < kk; k++) bbd[k] = (double)bb[k];

    if (doau) {  // do_bootrep
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
    if (outcnt && (doau || dobp)) {
        cntmat.resize((size_t)cm * kk);
        for (int i = 0; i < cm; i++)
            for (int k = 0; k < kk; k++) {
                const cgpu_rowsum &q = r0[(size_t)i * kk + k];
                cntmat[(size_t)i * kk + k] = smoothcnt ? cgpu_smooth_count(&q, bb[k], threshold) : q.cnt;
            }
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
        // the reference writes the rows after do_bootrep/do_bptest sorted them in place
        const bool sorted = doau || dobp;
        CG(cgpu_stats_download(reps, sorted ? 1 : 0, ptrs.data()));
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
    if (outcnt && !cntmat.empty()) {
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
        fprintf(fp, "\n# CNT:\n"); write_mat(fp, cntmat.data(), cm, kk);
        if (fname_out) fclose(fp);
    }
    cgpu_reps_free(reps);
    printf("\n# exit normally\n");
    cgpu_shutdown();
    return 0;
}
