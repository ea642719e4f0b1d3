// makermt -> consel in one process, on device-resident replicate blocks (the shape
// the reference ships as makerep, src/makerep.c:304-376), single GPU or one process per GPU.
//
// Everything `consel foo` derives from the replicates (do_mctest :1352-1439, do_bptest :1286-1303,
// do_bootrep :1441-1538 of src/consel.c) runs here against the C ABI of libconselgpu.so; with a
// communicator (cgpu_comm_init) the replicate columns are sharded over the ranks and
//   - the scale nearest r = 1 is all-gathered once (KH/SH tests and the median need it whole),
//   - the O(K^2 B) loops of the KH/SH tests and the AU fits are split over ITEMS,
//   - every AU round is one collective count query (cgpu_count_query_dist),
//   - the finished per-item results are all-gathered, so every rank returns the full table.
// All reductions involved are exact (integer counts, max, min), so the table is bit-identical to
// the single-GPU one whatever the number of ranks.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/consel_gpu.h"
#include "audriver.h"
#include "pipeline.h"

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

static int au_run_dist_opt(int cm_own, int kk, const double *rr, const int *bb, int i0, double threshold, const double *median,
                           const cgpu_rowsum *r0, int nalpha, const double *alpha, const aufit::Options &opt,
                           audriver::DistQueryFn query, void *ctx, double *pv2, double *se2, double *ax7, double *ci6, int *ret) {
    audriver::Input in;
    in.cm = cm_own; in.kk = kk; in.rr = rr; in.bb = bb; in.i0 = i0; in.threshold = threshold;
    in.median = median; in.r0 = r0; in.nalpha = nalpha; in.alpha = alpha; in.confidence = ci6 != nullptr;
    audriver::Output o;
    if (int rc = audriver::run(in, opt, nullptr, ctx, &o, nullptr, query)) return rc;
    for (int i = 0; i < cm_own; i++) {
        pv2[2 * i] = o.pv[i]; pv2[2 * i + 1] = o.pv0[i];
        se2[2 * i] = o.se[i]; se2[2 * i + 1] = o.se0[i];
        double *ax = ax7 + 7 * (size_t)i;
        ax[0] = o.pf[i]; ax[1] = o.rss[i]; ax[2] = o.df[i]; ax[3] = o.beta[3 * (size_t)i];
        ax[4] = o.beta[3 * (size_t)i + 1]; ax[5] = o.th[i]; ax[6] = o.beta[3 * (size_t)i + 2];
        if (ret) ret[i] = o.ret[i];
    }
    if (ci6) {
        const size_t n2 = (size_t)cm_own * nalpha;
        const std::vector<double> *m[6] = {&o.ci, &o.cs, &o.ei, &o.ci0, &o.cs0, &o.ei0};
        for (int t = 0; t < 6; t++) std::copy(m[t]->begin(), m[t]->end(), ci6 + t * n2);
    }
    return 0;
}

extern "C" {

// AU test of THIS rank's items through a caller-supplied collective query (the production
// callback is cgpu_count_query_dist; the CPU tests drive it over gloo): see audriver.h.
int chost_au_run_dist(int cm_own, int kk, const double *rr, const int *bb, int i0, double threshold, const double *median,
                      const cgpu_rowsum *r0, int nalpha, const double *alpha, int fitmode, double kappa,
                      audriver::DistQueryFn query, void *ctx, double *pv2, double *se2, double *ax7, double *ci6, int *ret) {
    aufit::Options opt;
    opt.fitmode = fitmode;
    opt.kappa = kappa;
    return au_run_dist_opt(cm_own, kk, rr, bb, i0, threshold, median, r0, nalpha, alpha, opt, query, ctx, pv2, se2, ax7, ci6, ret);
}

// The consel half of the pipeline.  Items are given in output order (read_assoc's sort is done by
// the caller: ass_ptr/ass_idx already permuted).  Outputs (every rank gets all of them):
//   pv, se   cm x 8 in the .pv column order [bp, pp, kh, sh, wkh, wsh, au, np]; column 1 (pp, a
//            function of datvec alone) is left to the caller
//   cnt      cm x kk plain counts #{stat <= threshold} (the .cnt payload)
//   ax7      cm x 7 [pf, rss, df, beta0, beta1, th, beta2];  ci6  6 x cm x nalpha or NULL;  ret cm
//   timing   [0] mctest, [1] statistics + first query, [2] AU (median, fits, queries, gather) seconds,
//            [3] number of AU query rounds is not tracked here (0)
int chost_consel_run(cgpu_reps *reps, int K, const double *datvec, int kk, const double *rr, const int *bb, int cm,
                     const int *ass_ptr, const int *ass_idx, double threshold, double varadd, int fitmode, double kappa,
                     int do_mc, int do_au, int nalpha, const double *alpha, double *pv, double *se, double *cnt, double *ax7,
                     double *ci6, int *ret, double *timing) {
    aufit::Options opt;
    opt.fitmode = fitmode;
    opt.kappa = kappa;
    return pipeline::consel_run(reps, K, datvec, kk, rr, bb, cm, ass_ptr, ass_idx, threshold, varadd, opt, do_mc, do_au, nalpha, alpha,
                                pv, se, cnt, ax7, ci6, ret, timing, nullptr);
}

}  // extern "C"

int pipeline::consel_run(cgpu_reps *reps, int K, const double *datvec, int kk, const double *rr, const int *bb, int cm,
                         const int *ass_ptr, const int *ass_idx, double threshold, double varadd, const aufit::Options &opt,
                         int do_mc, int do_au, int nalpha, const double *alpha, double *pv, double *se, double *cnt, double *ax7,
                         double *ci6, int *ret, double *timing, cgpu_rowsum *q0_out) {
    if (!reps || !datvec || !rr || !bb || !ass_ptr || !ass_idx || !pv || !se || cm <= 0 || kk <= 0) return CGPU_EINVAL;
    int rank = 0, world = 1;
    cgpu_comm_info(&rank, &world);
    const bool dist = world > 1;
    int i0 = 0;
    for (int k = 1; k < kk; k++)  // find_i0 (src/consel.c:1778-1789): the scale nearest 1, first one on ties
        if (std::fabs(rr[k] - 1.0) < std::fabs(rr[i0] - 1.0)) i0 = k;
    const int nb0 = bb[i0];
    int e0 = 0, ne = cm;
    cgpu_comm_item_range(cm, &e0, &ne);
    if (!dist) {
        e0 = 0;
        ne = cm;
    }
    double t0 = now_s();
    cgpu_reps *whole = nullptr;  // scale i0 on every rank (sharded runs)
    struct Guard {
        cgpu_reps *&w;
        ~Guard() { if (w) cgpu_reps_free(w); }
    } guard{whole};
    if (dist && (do_mc || (do_au && kk >= 2)))
        if (int rc = cgpu_reps_gather_scale(reps, i0, &whole)) return rc;

    // ---- KH / SH / wKH / wSH (do_mctest)
    if (do_mc) {
        std::vector<double> four((size_t)4 * cm, 0.0);
        double *kh = four.data(), *sh = kh + cm, *wkh = sh + cm, *wsh = wkh + cm;
        int rc;
        if (dist) {
            rc = cgpu_mctest_rows(whole, 0, datvec, varadd, cm, ass_ptr, ass_idx, e0, e0 + ne, kh, sh, wkh, wsh);
            int rc2 = cgpu_comm_allreduce_host(four.data(), (long long)four.size(), 0);  // foreign entries are zero
            if (!rc) rc = rc2;
        } else {
            rc = cgpu_mctest(reps, i0, datvec, varadd, cm, ass_ptr, ass_idx, kh, sh, wkh, wsh);
        }
        if (rc) return rc;
        for (int i = 0; i < cm; i++) {
            pv[8 * (size_t)i + 2] = kh[i];
            pv[8 * (size_t)i + 3] = sh[i];
            pv[8 * (size_t)i + 4] = wkh[i];
            pv[8 * (size_t)i + 5] = wsh[i];
        }
    }
    if (timing) timing[0] = now_s() - t0;

    // ---- statistics + the count at `threshold` (genrep, do_bptest)
    t0 = now_s();
    if (int rc = cgpu_stats_build(reps, cm, ass_ptr, ass_idx)) return rc;
    std::vector<double> thr0((size_t)cm, threshold);
    std::vector<cgpu_rowsum> q0((size_t)cm * kk);
    {
        int any = 0;
        int rc = dist ? cgpu_count_query_dist(reps, 1, 1, thr0.data(), q0.data(), 1, &any) : cgpu_count_query(reps, 1, thr0.data(), q0.data());
        if (rc) return rc;
    }
    for (int i = 0; i < cm; i++) {
        const double sm = cgpu_smooth_count(&q0[(size_t)i * kk + i0], nb0, threshold);
        pv[8 * (size_t)i] = 1.0 - (1.0 - sm / nb0);  // do_bptest: bp = 1 - pvaldist
        if (cnt)
            for (int k = 0; k < kk; k++) cnt[(size_t)i * kk + k] = q0[(size_t)i * kk + k].cnt;
    }
    if (q0_out) std::copy(q0.begin(), q0.end(), q0_out);
    for (int i = 0; i < cm; i++)
        for (int c : {0, 2, 3, 4, 5}) {  // getseval (src/consel.c:2151-2158)
            const double p = pv[8 * (size_t)i + c];
            se[8 * (size_t)i + c] = std::sqrt(p * (1.0 - p) / nb0);
        }
    if (timing) timing[1] = now_s() - t0;
    if (!do_au || kk < 2) return CGPU_OK;

    // ---- AU test (do_bootrep): the median of the scale-i0 rows starts vcalpval
    t0 = now_s();
    std::vector<double> med((size_t)std::max(ne, 1));
    std::vector<int> rk((size_t)std::max(ne, 1), nb0 / 2);
    if (dist) {
        if (ne > 0) {
            std::vector<int> optr((size_t)ne + 1);
            for (int i = 0; i <= ne; i++) optr[i] = ass_ptr[e0 + i] - ass_ptr[e0];
            if (int rc = cgpu_stats_build(whole, ne, optr.data(), ass_idx + ass_ptr[e0])) return rc;
            if (int rc = cgpu_order_stat(whole, 0, rk.data(), med.data())) return rc;
        }
    } else if (int rc = cgpu_order_stat(reps, i0, rk.data(), med.data())) {
        return rc;
    }
    const int na = ci6 ? nalpha : 0;
    const size_t per = dist ? (size_t)((cm + world - 1) / world) : (size_t)cm;
    const size_t width = 2 + 2 + 7 + 1 + 6 * (size_t)na;  // doubles per item in the exchange block
    std::vector<double> pv2(2 * per, 0.0), se2(2 * per, 0.0), ax(7 * per, 0.0), ci(6 * per * std::max(na, 1), 0.0);
    std::vector<int> rt(per, 0);
    int rc;
    if (dist) {
        auto dq = [](void *ctx, int nq, const double *thr, cgpu_rowsum *out, int active, int *any) {
            return cgpu_count_query_dist(static_cast<const cgpu_reps *>(ctx), 0, nq, thr, out, active, any);
        };
        rc = au_run_dist_opt(ne, kk, rr, bb, i0, threshold, med.data(), q0.data() + (size_t)e0 * kk, na, alpha, opt,
                             dq, reps, pv2.data(), se2.data(), ax.data(), ci6 ? ci.data() : nullptr, rt.data());
    } else {
        audriver::Input in;
        in.cm = cm; in.kk = kk; in.rr = rr; in.bb = bb; in.i0 = i0; in.threshold = threshold;
        in.median = med.data(); in.r0 = q0.data(); in.nalpha = na; in.alpha = alpha; in.confidence = ci6 != nullptr;
        audriver::Output o;
        auto query = [](void *ctx, int nq, const double *thr, cgpu_rowsum *out) {
            return cgpu_count_query(static_cast<const cgpu_reps *>(ctx), nq, thr, out);
        };
        auto query_rows = [](void *ctx, int row0, int nrows, int nq, const double *thr, cgpu_rowsum *out) {
            return cgpu_count_query_rows(static_cast<const cgpu_reps *>(ctx), row0, nrows, nq, thr, out);
        };
        rc = audriver::run(in, opt, query, reps, &o, query_rows);
        if (!rc) {
            for (int i = 0; i < cm; i++) {
                pv2[2 * (size_t)i] = o.pv[i]; pv2[2 * (size_t)i + 1] = o.pv0[i];
                se2[2 * (size_t)i] = o.se[i]; se2[2 * (size_t)i + 1] = o.se0[i];
                double *a = &ax[7 * (size_t)i];
                a[0] = o.pf[i]; a[1] = o.rss[i]; a[2] = o.df[i]; a[3] = o.beta[3 * (size_t)i];
                a[4] = o.beta[3 * (size_t)i + 1]; a[5] = o.th[i]; a[6] = o.beta[3 * (size_t)i + 2];
                rt[i] = o.ret[i];
            }
            if (ci6) {
                const size_t n2 = (size_t)cm * na;
                const std::vector<double> *m[6] = {&o.ci, &o.cs, &o.ei, &o.ci0, &o.cs0, &o.ei0};
                for (int t = 0; t < 6; t++) std::copy(m[t]->begin(), m[t]->end(), ci.begin() + t * n2);
            }
        }
    }
    if (rc) return rc;
    // ---- every rank gets the whole table: one all-gather of fixed-size per-rank blocks
    std::vector<double> blk(per * width, 0.0), all;
    auto pack = [&](size_t i, double *b) {
        b[0] = pv2[2 * i]; b[1] = pv2[2 * i + 1]; b[2] = se2[2 * i]; b[3] = se2[2 * i + 1];
        for (int c = 0; c < 7; c++) b[4 + c] = ax[7 * i + c];
        b[11] = (double)rt[i];
        const size_t n2 = (dist ? (size_t)ne : (size_t)cm) * na;
        for (int t = 0; t < 6; t++)
            for (int a = 0; a < na; a++) b[12 + (size_t)t * na + a] = ci[t * n2 + i * na + a];
    };
    const size_t nown = dist ? (size_t)ne : (size_t)cm;
    for (size_t i = 0; i < nown; i++) pack(i, &blk[i * width]);
    const double *src = blk.data();
    if (dist) {
        all.resize(per * width * world);
        if (int rc2 = cgpu_comm_allgather_host(blk.data(), all.data(), (long long)(per * width * 8))) return rc2;
        src = all.data();  // rank p's items start at row p*per, exactly the padded layout
    }
    for (int i = 0; i < cm; i++) {
        const double *b = src + (size_t)i * width;
        pv[8 * (size_t)i + 6] = b[0]; pv[8 * (size_t)i + 7] = b[1];
        se[8 * (size_t)i + 6] = b[2]; se[8 * (size_t)i + 7] = b[3];
        if (ax7) for (int c = 0; c < 7; c++) ax7[7 * (size_t)i + c] = b[4 + c];
        if (ret) ret[i] = (int)b[11];
        if (ci6)
            for (int t = 0; t < 6; t++)
                for (int a = 0; a < na; a++) ci6[((size_t)t * cm + i) * na + a] = b[12 + (size_t)t * na + a];
    }
    if (timing) timing[2] = now_s() - t0;
    return CGPU_OK;
}


