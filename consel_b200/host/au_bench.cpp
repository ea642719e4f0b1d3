// Development tool (not shipped to users): times the HOST half of the AU test -- the lockstep
// driver of audriver.cpp with its fits -- at the item / scale / replicate counts of a BASELINE
// configuration, without a GPU.  The device's count query is replaced by a binary search over
// host-sorted statistic rows (same cgpu_rowsum answers as cgpu_count_query gives), so what is
// measured is exactly the code that bounds the consel side of the pipeline on the GPU box:
// prep / fit rounds, thread pool, qnorm, the MLE Newton steps.
//
//   au_host_bench [cm=1000] [B=10000] [repeat=3]        (CHOST_THREADS, CHOST_TIMING as usual)
//
// Prints a checksum of every p-value and confidence limit: any change to the host fit must keep
// it bit for bit.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/consel_gpu.h"
#include "audriver.h"

namespace {
struct Rows {
    int cm, kk;
    std::vector<int> bb;
    std::vector<std::vector<double>> sorted;  // [s] cm x bb[s], every row ascending
};

uint64_t g_state = 0x9E3779B97F4A7C15ull;
double u01() {
    g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17;
    return ((double)(g_state >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
double gauss() { return std::sqrt(-2.0 * std::log(u01())) * std::cos(6.283185307179586 * u01()); }

void summarize(const double *v, int B, double t, cgpu_rowsum *q) {
    const int c = (int)(std::upper_bound(v, v + B, t) - v);
    q->cnt = c;
    q->lo = c > 0 ? v[c - 1] : -1.0e30;
    q->hi = c < B ? v[c] : 1.0e30;
    q->min1 = v[0]; q->min2 = v[B > 1 ? 1 : 0]; q->max1 = v[B - 1]; q->max2 = v[B > 1 ? B - 2 : 0];
}

int query_rows(void *ctx, int row0, int nrows, int nq, const double *thr, cgpu_rowsum *out) {
    const Rows &R = *static_cast<const Rows *>(ctx);
    for (int e = 0; e < nrows; e++)
        for (int s = 0; s < R.kk; s++)
            for (int a = 0; a < nq; a++)
                summarize(&R.sorted[s][(size_t)(row0 + e) * R.bb[s]], R.bb[s], thr[(size_t)e * nq + a],
                          &out[((size_t)e * R.kk + s) * nq + a]);
    return 0;
}
int query_all(void *ctx, int nq, const double *thr, cgpu_rowsum *out) {
    return query_rows(ctx, 0, static_cast<const Rows *>(ctx)->cm, nq, thr, out);
}
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

int main(int argc, char **argv) {
    const int cm = argc > 1 ? atoi(argv[1]) : 1000, B = argc > 2 ? atoi(argv[2]) : 10000, repeat = argc > 3 ? atoi(argv[3]) : 3;
    const int kk = 10;
    const double rr[kk] = {0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4};
    Rows R;
    R.cm = cm; R.kk = kk; R.bb.assign(kk, B);
    R.sorted.resize(kk);
    // statistic of item i at scale r: d_i + (z + c_i z^2 / 8) / sqrt(r) -- signed distance d_i, a little curvature
    std::vector<double> d(cm), c(cm);
    for (int i = 0; i < cm; i++) { d[i] = -3.0 + 6.0 * (i + 0.5) / cm; c[i] = 0.4 * (u01() - 0.5); }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    for (int s = 0; s < kk; s++) {
        R.sorted[s].resize((size_t)cm * B);
        for (size_t e = 0; e < (size_t)cm * B; e++) {
            const double z = gauss();
            const int i = (int)(e / B);
            R.sorted[s][e] = d[i] + (z + c[i] * (z * z - 1.0) / 8.0) / std::sqrt(rr[s]);
        }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < hw; t++)
            th.emplace_back([&, t] {
                for (int i = (int)t; i < cm; i += (int)hw) std::sort(&R.sorted[s][(size_t)i * B], &R.sorted[s][(size_t)i * B] + B);
            });
        for (auto &x : th) x.join();
    }
    const int i0 = 5;
    std::vector<double> med(cm);
    std::vector<cgpu_rowsum> r0((size_t)cm * kk);
    std::vector<double> zero(cm, 0.0);
    query_all(&R, 1, zero.data(), r0.data());
    for (int i = 0; i < cm; i++) med[i] = R.sorted[i0][(size_t)i * B + B / 2];
    const double alpha[5] = {0.05, 0.10, 0.5, 0.90, 0.95};
    audriver::Input in;
    in.cm = cm; in.kk = kk; in.rr = rr; in.bb = R.bb.data(); in.i0 = i0; in.threshold = 0.0;
    in.median = med.data(); in.r0 = r0.data(); in.nalpha = 5; in.alpha = alpha; in.confidence = true;
    aufit::Options opt;
    double best = 1e30;
    uint64_t sum = 0;
    for (int it = 0; it < repeat; it++) {
        audriver::Output o;
        const double t0 = now();
        if (int rc = audriver::run(in, opt, query_all, &R, &o, query_rows)) { fprintf(stderr, "run failed %d\n", rc); return 1; }
        const double t = now() - t0;
        best = std::min(best, t);
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const std::vector<double> &v) {
            for (double x : v) { uint64_t b; memcpy(&b, &x, 8); h = (h ^ b) * 1099511628211ull; }
        };
        mix(o.pv); mix(o.se); mix(o.pv0); mix(o.se0); mix(o.rss); mix(o.df); mix(o.th); mix(o.pf); mix(o.beta);
        mix(o.ci); mix(o.cs); mix(o.ei); mix(o.ci0); mix(o.cs0); mix(o.ei0);
        if (it && h != sum) { fprintf(stderr, "results differ between repeats\n"); return 1; }
        sum = h;
        printf("pass %d: %.2f ms\n", it, t * 1e3);
    }
    printf("cm=%d kk=%d B=%d: AU host driver best %.2f ms, checksum %016llx\n", cm, kk, B, best * 1e3, (unsigned long long)sum);
    return 0;
}
