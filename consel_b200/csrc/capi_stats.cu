// consel-side entry points of the C ABI (include/consel_gpu.h): replicate statistics,
// sort-free count queries, order statistics and the KH/SH/wKH/wSH tests.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "cgpu_internal.h"

namespace cgpu {

#define CG_HUGE 1.0e30

// small RAII holder so early returns release device buffers
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { dev_free(p); }
    template <typename T>
    T *as() { return reinterpret_cast<T *>(p); }
    int alloc(size_t bytes) {
        cudaError_t e = dev_alloc(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error(CGPU_ENOMEM, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
            return CGPU_ENOMEM;
        }
        return CGPU_OK;
    }
    template <typename T>
    int upload(const T *h, size_t count, cudaStream_t st) {
        CGPU_TRY(alloc(count * sizeof(T)));
        if (count) CGPU_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
        return CGPU_OK;
    }
};

static int check_csr(int K, int cm, const int *ptr, const int *idx) {
    if (cm <= 0 || !ptr || !idx || ptr[0] != 0) {
        set_error(CGPU_EINVAL, "association: cm=%d", cm);
        return CGPU_EINVAL;
    }
    for (int e = 0; e < cm; e++) {
        if (ptr[e + 1] < ptr[e]) {
            set_error(CGPU_EINVAL, "association %d has negative length", e);
            return CGPU_EINVAL;
        }
        for (int q = ptr[e]; q < ptr[e + 1]; q++)
            if (idx[q] < 0 || idx[q] >= K) {
                set_error(CGPU_EINVAL, "ass out of range: ass[%d][%d]=%d", e + 1, q - ptr[e] + 1, idx[q] + 1);
                return CGPU_EINVAL;
            }
    }
    return CGPU_OK;
}

// cntdist3 (src/consel.c:1842-1860) from a row summary
static double smooth_count(const cgpu_rowsum &q, int B, double t) {
    const double n = (double)B;
    const long long cnt = (long long)q.cnt;
    double p;
    if (B < 2) return cnt > 0 ? n : 0.0;
    if (cnt <= 0) {
        if (q.min2 > q.min1) p = 0.5 + (t - q.min1) / (q.min2 - q.min1);
        else p = 0.0;
    } else if (cnt < B) {
        const double i = (double)(cnt - 1);
        if (q.hi > q.lo) p = 0.5 + i + (t - q.lo) / (q.hi - q.lo);
        else p = 0.5 + i;
    } else {
        if (q.max1 - q.max2 > 0) p = n - 0.5 + (t - q.max1) / (q.max1 - q.max2);
        else p = n;
    }
    if (p > n) p = n;
    else if (p < 0.0) p = 0.0;
    return p;
}

}  // namespace cgpu

using namespace cgpu;

extern "C" {

double cgpu_smooth_count(const cgpu_rowsum *q, int B, double t) { return smooth_count(*q, B, t); }

void cgpu_smooth_count_many(const cgpu_rowsum *q, int count, int B, double t, double *out) {
    for (int i = 0; i < count; i++) out[i] = smooth_count(q[i], B, t);
}

int cgpu_reps_upload(int K, int S, const int *bb, const double *const *rep, cgpu_reps **out) {
    CGPU_TRY(ensure_init());
    if (K <= 0 || S <= 0 || !bb || !rep || !out) {
        set_error(CGPU_EINVAL, "reps_upload: bad arguments");
        return CGPU_EINVAL;
    }
    cudaStream_t st = G().stream;
    DevReps *r = new DevReps();
    r->K = K;
    r->S = S;
    r->bb = (int *)calloc(S, sizeof(int));
    r->b0 = (int *)calloc(S, sizeof(int));
    r->nb = (int *)calloc(S, sizeof(int));
    r->d_rep = (double **)calloc(S, sizeof(double *));
    int rc = CGPU_OK;
    for (int s = 0; s < S && rc == CGPU_OK; s++) {
        int b0 = (int)((long long)bb[s] * G().shard_rank / G().shard_world);
        int b1 = (int)((long long)bb[s] * (G().shard_rank + 1) / G().shard_world);
        r->bb[s] = bb[s];
        r->b0[s] = b0;
        r->nb[s] = b1 - b0;
        cudaError_t e = dev_alloc(&r->d_rep[s], (size_t)K * std::max(1, b1 - b0) * sizeof(double));
        if (e == cudaSuccess && b1 > b0)
            e = cudaMemcpy2DAsync(r->d_rep[s], (size_t)(b1 - b0) * sizeof(double), rep[s] + b0,
                                  (size_t)bb[s] * sizeof(double), (size_t)(b1 - b0) * sizeof(double), K,
                                  cudaMemcpyDefault, st);
        if (e != cudaSuccess) {
            set_error(CGPU_ECUDA, "reps_upload scale %d: %s", s, cudaGetErrorString(e));
            rc = CGPU_ECUDA;
        }
    }
    if (rc == CGPU_OK && cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(CGPU_ECUDA, "reps_upload: sync failed");
        rc = CGPU_ECUDA;
    }
    if (rc != CGPU_OK) {
        cgpu_reps_free(reinterpret_cast<cgpu_reps *>(r));
        return rc;
    }
    *out = reinterpret_cast<cgpu_reps *>(r);
    return CGPU_OK;
}

// Statistics that already exist (a `.rep` file, consel -R: do_repmode src/consel.c:1195-1242):
// the handle has no replicate blocks, only the cm x B_s statistic rows, and serves the count
// queries / order statistics / downloads like one built by cgpu_stats_build.
int cgpu_stats_upload(int cm, int S, const int *bb, const double *const *stat, cgpu_reps **out) {
    CGPU_TRY(ensure_init());
    if (cm <= 0 || S <= 0 || !bb || !stat || !out) {
        set_error(CGPU_EINVAL, "stats_upload: bad arguments");
        return CGPU_EINVAL;
    }
    cudaStream_t st = G().stream;
    DevReps *r = new DevReps();
    r->K = 0;
    r->S = S;
    r->cm = cm;
    r->bb = (int *)calloc(S, sizeof(int));
    r->b0 = (int *)calloc(S, sizeof(int));
    r->nb = (int *)calloc(S, sizeof(int));
    r->d_rep = (double **)calloc(S, sizeof(double *));
    r->d_stat = (double **)calloc(S, sizeof(double *));
    int rc = CGPU_OK;
    for (int s = 0; s < S && rc == CGPU_OK; s++) {
        if (bb[s] < 0 || (bb[s] > 0 && !stat[s])) {
            set_error(CGPU_EINVAL, "stats_upload: bad block %d", s);
            rc = CGPU_EINVAL;
            break;
        }
        const int b0 = (int)((long long)bb[s] * G().shard_rank / G().shard_world);
        const int b1 = (int)((long long)bb[s] * (G().shard_rank + 1) / G().shard_world);
        r->bb[s] = bb[s];
        r->b0[s] = b0;
        r->nb[s] = b1 - b0;
        cudaError_t e = dev_alloc(&r->d_stat[s], (size_t)cm * std::max(1, b1 - b0) * sizeof(double));
        if (e == cudaSuccess && b1 > b0)
            e = cudaMemcpy2DAsync(r->d_stat[s], (size_t)(b1 - b0) * sizeof(double), stat[s] + b0,
                                  (size_t)bb[s] * sizeof(double), (size_t)(b1 - b0) * sizeof(double), cm,
                                  cudaMemcpyDefault, st);
        if (e != cudaSuccess) {
            set_error(CGPU_ECUDA, "stats_upload scale %d: %s", s, cudaGetErrorString(e));
            rc = CGPU_ECUDA;
        }
    }
    if (rc == CGPU_OK && cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(CGPU_ECUDA, "stats_upload: sync failed");
        rc = CGPU_ECUDA;
    }
    if (rc != CGPU_OK) {
        cgpu_reps_free(reinterpret_cast<cgpu_reps *>(r));
        return rc;
    }
    *out = reinterpret_cast<cgpu_reps *>(r);
    return CGPU_OK;
}

int cgpu_stats_build(cgpu_reps *rr, int cm, const int *ass_ptr, const int *ass_idx) {
    CGPU_TRY(ensure_init());
    DevReps *r = reinterpret_cast<DevReps *>(rr);
    if (!r) return CGPU_EINVAL;
    if (r->K <= 0) {
        set_error(CGPU_EINVAL, "stats_build: the handle holds uploaded statistics, no replicate blocks");
        return CGPU_EINVAL;
    }
    CGPU_TRY(check_csr(r->K, cm, ass_ptr, ass_idx));
    cudaStream_t st = G().stream;
    if (r->d_stat) {
        for (int s = 0; s < r->S; s++) dev_free(r->d_stat[s]);
        free(r->d_stat);
        r->d_stat = nullptr;
    }
    dev_free(r->d_ass_ptr);
    dev_free(r->d_ass_idx);
    r->d_ass_ptr = r->d_ass_idx = nullptr;
    r->cm = cm;
    r->h_edges.clear();  // new statistics: the cached row edges are stale
    CGPU_CUDA(dev_alloc(&r->d_ass_ptr, (size_t)(cm + 1) * sizeof(int)));
    CGPU_CUDA(dev_alloc(&r->d_ass_idx, (size_t)std::max(1, ass_ptr[cm]) * sizeof(int)));
    CGPU_CUDA(cudaMemcpyAsync(r->d_ass_ptr, ass_ptr, (size_t)(cm + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    CGPU_CUDA(cudaMemcpyAsync(r->d_ass_idx, ass_idx, (size_t)ass_ptr[cm] * sizeof(int), cudaMemcpyHostToDevice, st));
    r->d_stat = (double **)calloc(r->S, sizeof(double *));
    for (int s = 0; s < r->S; s++) {
        const int B = r->nb[s];
        CGPU_CUDA(dev_alloc(&r->d_stat[s], (size_t)cm * std::max(1, B) * sizeof(double)));
        if (B <= 0) continue;
        DevBuf m1, m2, a1;
        CGPU_TRY(m1.alloc((size_t)B * 8));
        CGPU_TRY(m2.alloc((size_t)B * 8));
        CGPU_TRY(a1.alloc((size_t)B * 4));
        CGPU_TRY(launch_col_top2(r->d_rep[s], r->K, B, B, m1.as<double>(), m2.as<double>(), a1.as<int>(), st));
        CGPU_TRY(launch_stat_build(r->d_rep[s], r->K, B, B, m1.as<double>(), m2.as<double>(), a1.as<int>(), cm,
                                   r->d_ass_ptr, r->d_ass_idx, r->d_stat[s], st));
    }
    CGPU_CUDA(cudaStreamSynchronize(st));
    return CGPU_OK;
}

// consel -f (genrep's rescaling branch, src/consel.c:757-771, repminmaxsrs :1961-1984): the
// statistics of S1 target scales, all derived from the replicates of scale i0 pulled towards
// their row means by rs[s] = sqrt(rr[i0]/rr1[s]) (computed by the caller, like the reference).
int cgpu_stats_build_rescaled(const cgpu_reps *src, int i0, int S1, const double *rs, int cm, const int *ass_ptr,
                              const int *ass_idx, cgpu_reps **out) {
    CGPU_TRY(ensure_init());
    const DevReps *q = reinterpret_cast<const DevReps *>(src);
    if (!q || q->K <= 0 || i0 < 0 || i0 >= q->S || S1 <= 0 || !rs || !out) {
        set_error(CGPU_EINVAL, "stats_build_rescaled: bad arguments");
        return CGPU_EINVAL;
    }
    const int K = q->K, B = q->nb[i0];
    CGPU_TRY(check_csr(K, cm, ass_ptr, ass_idx));
    if (B != q->bb[i0] || B <= 0) {
        set_error(CGPU_EINVAL, "stats_build_rescaled needs every replicate of scale %d on this device (row means are sequential sums)", i0);
        return CGPU_EINVAL;
    }
    cudaStream_t st = G().stream;
    DevReps *r = new DevReps();
    r->K = 0;
    r->S = S1;
    r->cm = cm;
    r->bb = (int *)calloc(S1, sizeof(int));
    r->b0 = (int *)calloc(S1, sizeof(int));
    r->nb = (int *)calloc(S1, sizeof(int));
    r->d_rep = (double **)calloc(S1, sizeof(double *));
    r->d_stat = (double **)calloc(S1, sizeof(double *));
    auto fail = [&](int rc) {
        cgpu_reps_free(reinterpret_cast<cgpu_reps *>(r));
        return rc;
    };
    DevBuf xs, m1, m2, a1, d_ptr, d_idx;
    int rc = xs.alloc((size_t)K * B * 8);
    if (rc == CGPU_OK) rc = m1.alloc((size_t)B * 8);
    if (rc == CGPU_OK) rc = m2.alloc((size_t)B * 8);
    if (rc == CGPU_OK) rc = a1.alloc((size_t)B * 4);
    if (rc == CGPU_OK) rc = d_ptr.upload(ass_ptr, (size_t)(cm + 1), st);
    if (rc == CGPU_OK) rc = d_idx.upload(ass_idx, (size_t)std::max(1, ass_ptr[cm]), st);
    if (rc != CGPU_OK) return fail(rc);
    for (int s = 0; s < S1; s++) {
        r->bb[s] = r->nb[s] = B;
        if (dev_alloc(&r->d_stat[s], (size_t)cm * B * sizeof(double)) != cudaSuccess) {
            set_error(CGPU_ENOMEM, "stats_build_rescaled: device allocation for scale %d", s);
            return fail(CGPU_ENOMEM);
        }
        rc = launch_center(q->d_rep[i0], K, B, B, xs.as<double>(), st, true, rs[s]);
        if (rc == CGPU_OK) rc = launch_col_top2(xs.as<double>(), K, B, B, m1.as<double>(), m2.as<double>(), a1.as<int>(), st);
        if (rc == CGPU_OK)
            rc = launch_stat_build(xs.as<double>(), K, B, B, m1.as<double>(), m2.as<double>(), a1.as<int>(), cm,
                                   d_ptr.as<int>(), d_idx.as<int>(), r->d_stat[s], st);
        if (rc != CGPU_OK) return fail(rc);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(CGPU_ECUDA, "stats_build_rescaled: sync failed");
        return fail(CGPU_ECUDA);
    }
    *out = reinterpret_cast<cgpu_reps *>(r);
    return CGPU_OK;
}

int cgpu_stats_download(const cgpu_reps *rr, int sorted, double *const *stat_out) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !r->d_stat || !stat_out) {
        set_error(CGPU_EINVAL, "stats_download: statistics not built");
        return CGPU_EINVAL;
    }
    cudaStream_t st = G().stream;
    for (int s = 0; s < r->S; s++) {
        if (r->nb[s] <= 0) continue;
        CGPU_CUDA(cudaMemcpyAsync(stat_out[s], r->d_stat[s], (size_t)r->cm * r->nb[s] * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
    }
    CGPU_CUDA(cudaStreamSynchronize(st));
    if (sorted) {
        // `.rep` payload only (consel -r): rows are sorted on the host after the copy; the
        // AU/BP path itself never sorts (count queries below).
        for (int s = 0; s < r->S; s++) {
            if (r->nb[s] != r->bb[s]) {
                set_error(CGPU_EINVAL, "stats_download(sorted) needs unsharded rows");
                return CGPU_EINVAL;
            }
            for (int i = 0; i < r->cm; i++) std::sort(stat_out[s] + (size_t)i * r->bb[s], stat_out[s] + (size_t)(i + 1) * r->bb[s]);
        }
    }
    return CGPU_OK;
}

int cgpu_count_query_rows(const cgpu_reps *rr, int row0, int nrows, int nq, const double *thr, cgpu_rowsum *out) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !r->d_stat || !thr || !out || nq < 1 || row0 < 0 || nrows < 0 || row0 + nrows > r->cm) {
        set_error(CGPU_EINVAL, "count_query: statistics not built or bad arguments");
        return CGPU_EINVAL;
    }
    if (nrows == 0) return CGPU_OK;
    (void)cudaSetDevice(G().device);  // the AU driver calls from a helper thread
    cudaStream_t st = G().stream;
    const int cm = nrows, S = r->S;
    // row edges are cached once per statistics build, by the first whole-range query
    const bool whole = row0 == 0 && nrows == r->cm;
    const bool have_edges = r->h_edges.size() == (size_t)r->cm * S * 4;
    if (!whole && !have_edges && S <= 32) {
        set_error(CGPU_EINVAL, "count_query_rows: query all rows once before querying a sub-range");
        return CGPU_EINVAL;
    }
    DevBuf d_thr, d_out, d_part;
    CGPU_TRY(d_thr.upload(thr, (size_t)cm * nq, st));
    const size_t nout = (size_t)cm * S * nq;
    CGPU_TRY(d_out.alloc(nout * sizeof(cgpu_rowsum)));
    if (S <= 32) {
        // one launch for all scales; long rows are split so few items still fill the SMs
        long long maxB = 1;
        for (int s = 0; s < S; s++) maxB = std::max<long long>(maxB, r->nb[s]);
        int nsplit = (int)std::min<long long>(16, std::max<long long>(1, (4LL * G().sm_count + (long long)cm * S - 1) / ((long long)cm * S)));
        nsplit = (int)std::min<long long>(nsplit, std::max<long long>(1, maxB / 1024));
        CGPU_TRY(d_part.alloc(nout * nsplit * sizeof(cgpu_rowsum)));
        CGPU_TRY(launch_rowsum_all_scales(r->d_stat, r->nb, S, cm, nq, d_thr.as<double>(), d_part.as<cgpu_rowsum>(),
                                          d_out.as<cgpu_rowsum>(), nsplit, st, !have_edges, row0));
    } else {
        for (int s = 0; s < S; s++) {
            RowSrc src{r->d_stat[s] + (size_t)row0 * r->nb[s], r->nb[s], nullptr, nullptr, nullptr, nullptr, nullptr};
            CGPU_TRY(launch_rowsum(0, src, cm, r->nb[s], nq, d_thr.as<double>(), d_out.as<cgpu_rowsum>(), S, s, st));
        }
    }
    CGPU_CUDA(cudaMemcpyAsync(out, d_out.p, nout * sizeof(cgpu_rowsum), cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    if (S <= 32) {
        if (!have_edges) {
            r->h_edges.resize((size_t)cm * S * 4);
            for (size_t e = 0; e < (size_t)cm * S; e++) {
                const cgpu_rowsum &q = out[e * nq];
                double *c = &r->h_edges[4 * e];
                c[0] = q.min1; c[1] = q.min2; c[2] = q.max1; c[3] = q.max2;
            }
        } else {
            for (size_t e = 0; e < (size_t)cm * S; e++) {
                const double *c = &r->h_edges[4 * ((size_t)row0 * S + e)];
                for (int q = 0; q < nq; q++) {
                    cgpu_rowsum &o = out[e * nq + q];
                    o.min1 = c[0]; o.min2 = c[1]; o.max1 = c[2]; o.max2 = c[3];
                }
            }
        }
    }
    return CGPU_OK;
}

int cgpu_count_query(const cgpu_reps *rr, int nq, const double *thr, cgpu_rowsum *out) {
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r) {
        set_error(CGPU_EINVAL, "count_query: statistics not built or bad arguments");
        return CGPU_EINVAL;
    }
    return cgpu_count_query_rows(rr, 0, r->cm, nq, thr, out);
}

int cgpu_order_stat(const cgpu_reps *rr, int s, const int *rank, double *val) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !r->d_stat || s < 0 || s >= r->S || !rank || !val) {
        set_error(CGPU_EINVAL, "order_stat: statistics not built or bad arguments");
        return CGPU_EINVAL;
    }
    if (r->nb[s] != r->bb[s]) {
        set_error(CGPU_EINVAL, "order_stat needs the whole row on one device (gather scale %d first)", s);
        return CGPU_EINVAL;
    }
    for (int i = 0; i < r->cm; i++)
        if (rank[i] < 0 || rank[i] >= r->bb[s]) {
            set_error(CGPU_EINVAL, "order_stat: rank[%d]=%d of %d", i, rank[i], r->bb[s]);
            return CGPU_EINVAL;
        }
    cudaStream_t st = G().stream;
    DevBuf d_rank, d_val;
    CGPU_TRY(d_rank.upload(rank, (size_t)r->cm, st));
    CGPU_TRY(d_val.alloc((size_t)r->cm * 8));
    CGPU_TRY(launch_order_stat(r->d_stat[s], r->cm, r->nb[s], r->nb[s], d_rank.as<int>(), d_val.as<double>(), st));
    CGPU_CUDA(cudaMemcpyAsync(val, d_val.p, (size_t)r->cm * 8, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    return CGPU_OK;
}

// KH/SH/wKH/wSH for the items [e0, e1) from the whole replicate block x (K x B, ld = B) of the
// scale nearest r = 1.  Host bookkeeping covers all items (it is O(nmem)); every device loop
// runs over the rows / members of the range only.
static int mctest_core(const double *x, int K, int B, const double *datvec, double varadd, int cm, const int *ass_ptr,
                       const int *ass_idx, int e0, int e1, double *kh, double *sh, double *wkh, double *wsh) {
    cudaStream_t st = G().stream;
    const int nmem = ass_ptr[cm];
    const int ne = e1 - e0;
    if (ne <= 0) return CGPU_OK;
    const int q_lo = ass_ptr[e0], q_hi = ass_ptr[e1], nq_own = q_hi - q_lo;  // members of the range

    // ---- host bookkeeping: members, and explicit complements (getcass, src/consel.c:2040-2063)
    // unless every item is a single tree -- then the complement of member i is "every j != i" and
    // the kernels walk it implicitly (an explicit list would be cm x (K-1) integers)
    bool all_single = true;
    for (int e = 0; e < cm; e++) all_single = all_single && (ass_ptr[e + 1] - ass_ptr[e] == 1);
    std::vector<int> cass_ptr(cm + 1, 0), cass_idx, mem_item(nmem), mem_tree(nmem);
    {
        std::vector<char> mark(K);
        for (int e = 0; e < cm; e++) {
            if (!all_single) std::fill(mark.begin(), mark.end(), 0);
            for (int q = ass_ptr[e]; q < ass_ptr[e + 1]; q++) {
                if (!all_single) {
                    if (mark[ass_idx[q]]) {
                        set_error(CGPU_EINVAL, "duplicate associations");
                        return CGPU_EINVAL;
                    }
                    mark[ass_idx[q]] = 1;
                }
                mem_item[q] = e;
                mem_tree[q] = ass_idx[q];
            }
            if (!all_single && e >= e0 && e < e1)
                for (int j = 0; j < K; j++)
                    if (!mark[j]) cass_idx.push_back(j);
            cass_ptr[e + 1] = (int)cass_idx.size();
        }
    }

    // first-argmax top-2 of datvec (for "max over all j != i" on the host)
    int da = 0, db = -1;
    for (int j = 1; j < K; j++)
        if (datvec[j] > datvec[da]) da = j;
    for (int j = 0; j < K; j++)
        if (j != da && (db < 0 || datvec[j] > datvec[db])) db = j;

    // ---- device: centred copy of the block
    DevBuf xc;
    CGPU_TRY(xc.alloc((size_t)K * B * 8));
    CGPU_TRY(launch_center(x, K, B, B, xc.as<double>(), st));

    // ---- unweighted KH contrasts (src/consel.c:2135-2143) and SH observed statistics
    std::vector<int> kia(cm), kja(cm);
    std::vector<double> kx0(cm), shobs(nmem);
    for (int e = e0; e < e1; e++) {
        const int a0 = ass_ptr[e], alen = ass_ptr[e + 1] - a0;
        int ja;
        if (alen == 1) {
            ja = (ass_idx[a0] == da) ? db : da;
        } else {
            double xx = -CG_HUGE;
            ja = cass_idx[cass_ptr[e]];
            for (int p = cass_ptr[e]; p < cass_ptr[e + 1]; p++)
                if (datvec[cass_idx[p]] > xx) { xx = datvec[cass_idx[p]]; ja = cass_idx[p]; }
        }
        double xx = CG_HUGE;
        int ia = ass_idx[a0];
        for (int q = a0; q < a0 + alen; q++) {
            double y = -datvec[ass_idx[q]];
            if (y < xx) { xx = y; ia = ass_idx[q]; }
        }
        kia[e] = ia;
        kja[e] = ja;
        kx0[e] = datvec[ja] - datvec[ia];
        // calcmaxass(datvec, NULL, ...) (src/consel.c:2026-2030): max over the complement with the -1e30 floor
        double mx = -CG_HUGE;
        if (datvec[ja] > mx) mx = datvec[ja];
        for (int q = a0; q < a0 + alen; q++) shobs[q] = mx - datvec[ass_idx[q]];
    }

    DevBuf d_ia, d_ja, d_thr, d_rs;
    CGPU_TRY(d_ia.upload(kia.data() + e0, (size_t)ne, st));
    CGPU_TRY(d_ja.upload(kja.data() + e0, (size_t)ne, st));
    CGPU_TRY(d_thr.upload(kx0.data() + e0, (size_t)ne, st));
    CGPU_TRY(d_rs.alloc((size_t)std::max(ne, nq_own) * sizeof(cgpu_rowsum)));
    std::vector<cgpu_rowsum> rs((size_t)std::max(ne, nq_own));  // row q - q_lo (members) or e - e0 (items)
    auto fetch = [&](int rows) -> int {
        CGPU_CUDA(cudaMemcpyAsync(rs.data(), d_rs.p, (size_t)rows * sizeof(cgpu_rowsum), cudaMemcpyDeviceToHost, st));
        CGPU_CUDA(cudaStreamSynchronize(st));
        return CGPU_OK;
    };
    const bool smooth = G().opt_mc_smooth != 0;  // sw_pvaldist; 0 = --quick_mctest
    auto pval = [&](const cgpu_rowsum &q, double t) {  // pvaldist :1876-1884
        return 1.0 - (smooth ? smooth_count(q, B, t) : q.cnt) / B;
    };

    {  // KH
        RowSrc src{xc.as<double>(), B, d_ia.as<int>(), d_ja.as<int>(), nullptr, nullptr, nullptr};
        CGPU_TRY(launch_rowsum(1, src, ne, B, 1, d_thr.as<double>(), d_rs.as<cgpu_rowsum>(), 1, 0, st));
        CGPU_TRY(fetch(ne));
        for (int e = e0; e < e1; e++) kh[e] = pval(rs[e - e0], kx0[e]);
    }

    // members / complements of the range on the device
    DevBuf d_mem_item, d_mem_tree, d_cass_ptr, d_cass_idx, d_tmp;
    CGPU_TRY(d_mem_item.upload(mem_item.data() + q_lo, (size_t)std::max(1, nq_own), st));
    CGPU_TRY(d_mem_tree.upload(mem_tree.data() + q_lo, (size_t)std::max(1, nq_own), st));
    auto reduce_items = [&](const std::vector<double> &thr, double *outp) {
        for (int e = e0; e < e1; e++) {
            double xx = 0.0;
            for (int q = ass_ptr[e]; q < ass_ptr[e + 1]; q++) {
                double p = pval(rs[q - q_lo], thr[q]);
                if (p > xx) xx = p;
            }
            outp[e] = xx;
        }
    };

    // every member as its own "row": rows of nq_own x B in chunks that fit the workspace
    const int chunk_rows = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, nq_own), ((size_t)2 << 30) / ((size_t)B * 8)));
    if (!all_single) {
        CGPU_TRY(d_cass_ptr.upload(cass_ptr.data(), (size_t)cm + 1, st));
        CGPU_TRY(d_cass_idx.upload(cass_idx.data(), std::max<size_t>(1, cass_idx.size()), st));
    }
    CGPU_TRY(d_tmp.alloc((size_t)chunk_rows * B * 8));

    {  // SH (unweighted)
        DevBuf d_t;
        CGPU_TRY(d_t.upload(shobs.data() + q_lo, (size_t)std::max(1, nq_own), st));
        if (all_single) {
            DevBuf m1, m2, a1;
            CGPU_TRY(m1.alloc((size_t)B * 8));
            CGPU_TRY(m2.alloc((size_t)B * 8));
            CGPU_TRY(a1.alloc((size_t)B * 4));
            CGPU_TRY(launch_col_top2(xc.as<double>(), K, B, B, m1.as<double>(), m2.as<double>(), a1.as<int>(), st));
            RowSrc src{xc.as<double>(), B, d_mem_tree.as<int>(), nullptr, m1.as<double>(), m2.as<double>(), a1.as<int>()};
            CGPU_TRY(launch_rowsum(2, src, nq_own, B, 1, d_t.as<double>(), d_rs.as<cgpu_rowsum>(), 1, 0, st));
            CGPU_TRY(fetch(nq_own));
        } else {
            for (int q0 = 0; q0 < nq_own; q0 += chunk_rows) {
                const int rows = std::min(chunk_rows, nq_own - q0);
                CGPU_TRY(launch_sh_rows(false, xc.as<double>(), K, B, B, nullptr, rows, d_mem_item.as<int>() + q0,
                                        d_mem_tree.as<int>() + q0, d_cass_ptr.as<int>(), d_cass_idx.as<int>(),
                                        d_tmp.as<double>(), st));
                RowSrc src{d_tmp.as<double>(), B, nullptr, nullptr, nullptr, nullptr, nullptr};
                CGPU_TRY(launch_rowsum(0, src, rows, B, 1, d_t.as<double>() + q0, d_rs.as<cgpu_rowsum>() + q0, 1, 0, st));
            }
            CGPU_TRY(fetch(nq_own));
        }
        reduce_items(shobs, sh);
    }

    // ---- weights (src/consel.c:1404-1415): the rows of the trees this range's members need
    DevBuf d_wt, d_nsmall, d_dat, d_rows;
    CGPU_TRY(d_wt.alloc((size_t)K * K * 8));
    CGPU_TRY(d_nsmall.alloc(4));
    CGPU_CUDA(cudaMemsetAsync(d_nsmall.p, 0, 4, st));
    if (e0 == 0 && e1 == cm) {
        CGPU_TRY(launch_pair_weight(xc.as<double>(), K, B, B, varadd, d_wt.as<double>(), d_nsmall.as<int>(), st));
    } else {
        std::vector<int> rows(mem_tree.begin() + q_lo, mem_tree.begin() + q_hi);
        std::sort(rows.begin(), rows.end());
        rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        CGPU_TRY(d_rows.upload(rows.data(), std::max<size_t>(1, rows.size()), st));
        CGPU_TRY(launch_pair_weight(xc.as<double>(), K, B, B, varadd, d_wt.as<double>(), d_nsmall.as<int>(), st,
                                    d_rows.as<int>(), (int)rows.size()));
    }
    CGPU_TRY(d_dat.upload(datvec, (size_t)K, st));

    // weighted observed statistics and contrasts
    DevBuf d_wobs, d_argj;
    CGPU_TRY(d_wobs.alloc((size_t)std::max(1, nq_own) * 8));
    CGPU_TRY(d_argj.alloc((size_t)std::max(1, nq_own) * 4));
    CGPU_TRY(launch_weighted_obs(d_dat.as<double>(), K, d_wt.as<double>(), nq_own, d_mem_item.as<int>(),
                                 d_mem_tree.as<int>(), all_single ? nullptr : d_cass_ptr.as<int>(),
                                 all_single ? nullptr : d_cass_idx.as<int>(), d_wobs.as<double>(), d_argj.as<int>(), st));
    std::vector<double> wobs(nmem);
    std::vector<int> argj(nmem);
    CGPU_CUDA(cudaMemcpyAsync(wobs.data() + q_lo, d_wobs.p, (size_t)nq_own * 8, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaMemcpyAsync(argj.data() + q_lo, d_argj.p, (size_t)nq_own * 4, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));

    {  // wKH: the weights only choose the contrast (src/consel.c:2121-2134, 2145)
        for (int e = e0; e < e1; e++) {
            double x0 = CG_HUGE;
            int qi = ass_ptr[e];
            for (int q = ass_ptr[e]; q < ass_ptr[e + 1]; q++)
                if (wobs[q] < x0) { x0 = wobs[q]; qi = q; }
            kia[e] = ass_idx[qi];
            kja[e] = argj[qi];  // the tree that attains the weighted maximum over the complement
            kx0[e] = datvec[kja[e]] - datvec[kia[e]];
        }
        CGPU_CUDA(cudaMemcpyAsync(d_ia.p, kia.data() + e0, (size_t)ne * 4, cudaMemcpyHostToDevice, st));
        CGPU_CUDA(cudaMemcpyAsync(d_ja.p, kja.data() + e0, (size_t)ne * 4, cudaMemcpyHostToDevice, st));
        CGPU_CUDA(cudaMemcpyAsync(d_thr.p, kx0.data() + e0, (size_t)ne * 8, cudaMemcpyHostToDevice, st));
        RowSrc src{xc.as<double>(), B, d_ia.as<int>(), d_ja.as<int>(), nullptr, nullptr, nullptr};
        CGPU_TRY(launch_rowsum(1, src, ne, B, 1, d_thr.as<double>(), d_rs.as<cgpu_rowsum>(), 1, 0, st));
        CGPU_TRY(fetch(ne));
        for (int e = e0; e < e1; e++) wkh[e] = pval(rs[e - e0], kx0[e]);
    }

    {  // wSH
        for (int q0 = 0; q0 < nq_own; q0 += chunk_rows) {
            const int rows = std::min(chunk_rows, nq_own - q0);
            if (all_single)
                CGPU_TRY(launch_wsh_single(xc.as<double>(), K, B, B, d_wt.as<double>(), d_mem_tree.as<int>() + q0, rows,
                                           d_tmp.as<double>(), st));
            else
                CGPU_TRY(launch_sh_rows(true, xc.as<double>(), K, B, B, d_wt.as<double>(), rows, d_mem_item.as<int>() + q0,
                                        d_mem_tree.as<int>() + q0, d_cass_ptr.as<int>(), d_cass_idx.as<int>(),
                                        d_tmp.as<double>(), st));
            RowSrc src{d_tmp.as<double>(), B, nullptr, nullptr, nullptr, nullptr, nullptr};
            CGPU_TRY(launch_rowsum(0, src, rows, B, 1, d_wobs.as<double>() + q0, d_rs.as<cgpu_rowsum>() + q0, 1, 0, st));
        }
        CGPU_TRY(fetch(nq_own));
        reduce_items(wobs, wsh);
    }
    return CGPU_OK;
}

static int mctest_check(const DevReps *r, int i0, const double *datvec, int cm, const int *ass_ptr, const int *ass_idx,
                        const double *kh, const double *sh, const double *wkh, const double *wsh) {
    if (!r || i0 < 0 || i0 >= r->S || !datvec || !kh || !sh || !wkh || !wsh) {
        set_error(CGPU_EINVAL, "mctest: bad arguments");
        return CGPU_EINVAL;
    }
    if (r->K <= 0) {
        set_error(CGPU_EINVAL, "mctest: the handle holds uploaded statistics, no replicate blocks");
        return CGPU_EINVAL;
    }
    CGPU_TRY(check_csr(r->K, cm, ass_ptr, ass_idx));
    if (r->bb[i0] < 2 || r->K < 2) {
        set_error(CGPU_EINVAL, "mctest: K=%d B=%d", r->K, r->bb[i0]);
        return CGPU_EINVAL;
    }
    return CGPU_OK;
}

int cgpu_mctest_rows(const cgpu_reps *rr, int i0, const double *datvec, double varadd, int cm, const int *ass_ptr,
                     const int *ass_idx, int e0, int e1, double *kh, double *sh, double *wkh, double *wsh) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    CGPU_TRY(mctest_check(r, i0, datvec, cm, ass_ptr, ass_idx, kh, sh, wkh, wsh));
    if (r->nb[i0] != r->bb[i0]) {
        set_error(CGPU_EINVAL, "mctest needs every replicate of scale %d on this device (row means are sequential sums): "
                               "cgpu_reps_gather_scale first", i0);
        return CGPU_EINVAL;
    }
    if (e0 < 0 || e1 > cm || e0 > e1) {
        set_error(CGPU_EINVAL, "mctest_rows: items [%d, %d) of %d", e0, e1, cm);
        return CGPU_EINVAL;
    }
    return mctest_core(r->d_rep[i0], r->K, r->bb[i0], datvec, varadd, cm, ass_ptr, ass_idx, e0, e1, kh, sh, wkh, wsh);
}

int cgpu_mctest(const cgpu_reps *rr, int i0, const double *datvec, double varadd, int cm, const int *ass_ptr,
                const int *ass_idx, double *kh, double *sh, double *wkh, double *wsh) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    CGPU_TRY(mctest_check(r, i0, datvec, cm, ass_ptr, ass_idx, kh, sh, wkh, wsh));
    if (r->nb[i0] == r->bb[i0])
        return mctest_core(r->d_rep[i0], r->K, r->bb[i0], datvec, varadd, cm, ass_ptr, ass_idx, 0, cm, kh, sh, wkh, wsh);
    if (comm_world() == 1) {
        set_error(CGPU_EINVAL, "mctest needs every replicate of scale %d on this device (row means are sequential sums)", i0);
        return CGPU_EINVAL;
    }
    // column-sharded handle + communicator: gather the scale, every rank tests its own items,
    // the four vectors are summed (the entries of foreign items are zero)
    cgpu_reps *whole = nullptr;
    CGPU_TRY(cgpu_reps_gather_scale(rr, i0, &whole));
    int e0 = 0, ne = 0;
    cgpu_comm_item_range(cm, &e0, &ne);
    std::vector<double> all((size_t)4 * cm, 0.0);
    int rc = cgpu_mctest_rows(whole, 0, datvec, varadd, cm, ass_ptr, ass_idx, e0, e0 + ne, all.data(), all.data() + cm,
                              all.data() + 2 * (size_t)cm, all.data() + 3 * (size_t)cm);
    cgpu_reps_free(whole);
    // (a failing rank still joins the collective so that the others do not hang; its error wins)
    int rc2 = cgpu_comm_allreduce_host(all.data(), (long long)all.size(), 0);
    if (rc != CGPU_OK) return rc;
    CGPU_TRY(rc2);
    memcpy(kh, all.data(), (size_t)cm * 8);
    memcpy(sh, all.data() + cm, (size_t)cm * 8);
    memcpy(wkh, all.data() + 2 * (size_t)cm, (size_t)cm * 8);
    memcpy(wsh, all.data() + 3 * (size_t)cm, (size_t)cm * 8);
    return CGPU_OK;
}

int cgpu_reps_gather_scale(const cgpu_reps *rr, int s, cgpu_reps **whole) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !whole || s < 0 || s >= r->S || r->K <= 0) {
        set_error(CGPU_EINVAL, "reps_gather_scale: bad arguments");
        return CGPU_EINVAL;
    }
    (void)cudaSetDevice(G().device);
    cudaStream_t st = G().stream;
    const int K = r->K, bb = r->bb[s], world = comm_world();
    DevReps *w = new DevReps();
    w->K = K;
    w->S = 1;
    w->bb = (int *)calloc(1, sizeof(int));
    w->b0 = (int *)calloc(1, sizeof(int));
    w->nb = (int *)calloc(1, sizeof(int));
    w->d_rep = (double **)calloc(1, sizeof(double *));
    w->bb[0] = w->nb[0] = bb;
    auto fail = [&](int rc) {
        cgpu_reps_free(reinterpret_cast<cgpu_reps *>(w));
        return rc;
    };
    if (dev_alloc(&w->d_rep[0], (size_t)K * std::max(1, bb) * 8) != cudaSuccess) {
        set_error(CGPU_ENOMEM, "reps_gather_scale: device allocation of %zu bytes", (size_t)K * bb * 8);
        return fail(CGPU_ENOMEM);
    }
    if (world == 1 || r->nb[s] == bb) {
        if (cudaMemcpyAsync(w->d_rep[0], r->d_rep[s], (size_t)K * bb * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return fail(CGPU_ECUDA);
    } else {
        // equal-sized messages: every rank's block is padded to nbmax columns
        const int nbmax = (bb + world - 1) / world + 1;
        DevBuf send, recv;
        int rc = send.alloc((size_t)K * nbmax * 8);
        if (rc == CGPU_OK) rc = recv.alloc((size_t)world * K * nbmax * 8);
        if (rc != CGPU_OK) return fail(rc);
        cudaError_t e = cudaSuccess;
        if (r->nb[s] > 0)
            e = cudaMemcpy2DAsync(send.p, (size_t)nbmax * 8, r->d_rep[s], (size_t)r->nb[s] * 8, (size_t)r->nb[s] * 8, K,
                                  cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) {
            set_error(CGPU_ECUDA, "reps_gather_scale: %s", cudaGetErrorString(e));
            return fail(CGPU_ECUDA);
        }
        rc = comm_allgather_f64(send.as<double>(), recv.as<double>(), (size_t)K * nbmax, st);
        if (rc != CGPU_OK) return fail(rc);
        for (int p = 0; p < world; p++) {
            const int c0 = (int)((long long)bb * p / world), c1 = (int)((long long)bb * (p + 1) / world);
            if (c1 <= c0) continue;
            e = cudaMemcpy2DAsync(w->d_rep[0] + c0, (size_t)bb * 8, recv.as<double>() + (size_t)p * K * nbmax, (size_t)nbmax * 8,
                                  (size_t)(c1 - c0) * 8, K, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) {
                set_error(CGPU_ECUDA, "reps_gather_scale: %s", cudaGetErrorString(e));
                return fail(CGPU_ECUDA);
            }
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(CGPU_ECUDA, "reps_gather_scale: sync failed");
        return fail(CGPU_ECUDA);
    }
    *whole = reinterpret_cast<cgpu_reps *>(w);
    return CGPU_OK;
}

int cgpu_count_query_dist(const cgpu_reps *rr, int all_rows, int nq, const double *thr, cgpu_rowsum *out, int active,
                          int *any_active) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !r->d_stat || nq < 1 || nq > 16) {
        set_error(CGPU_EINVAL, "count_query_dist: statistics not built or bad arguments");
        return CGPU_EINVAL;
    }
    const int world = comm_world(), rank = comm_rank();
    const int cm = r->cm, S = r->S;
    if (world == 1) {
        if (any_active) *any_active = active ? 1 : 0;
        if (!thr || !out) return cm ? CGPU_EINVAL : CGPU_OK;
        return cgpu_count_query_rows(rr, 0, cm, nq, thr, out);
    }
    (void)cudaSetDevice(G().device);  // the AU driver calls from a helper thread
    cudaStream_t st = G().stream;
    const int per = (cm + world - 1) / world, cm_pad = per * world;
    const int row0 = std::min(cm, rank * per), nrows = std::min(cm, row0 + per) - row0;
    if (nrows > 0 && !all_rows && (!thr || !out) && active) {
        set_error(CGPU_EINVAL, "count_query_dist: active rank without thresholds");
        return CGPU_EINVAL;
    }
    const size_t n_thr = (size_t)cm_pad * nq, n_out = (size_t)cm_pad * S * nq, n_loc = (size_t)per * S * nq;
    const bool have_edges = r->h_edges.size() == (size_t)cm * S * 4;
    // one pinned staging block per call site would be nicer; these are tens of KB
    DevBuf d_send, d_thr, d_flag, d_q, d_part, d_planes, d_red, d_edge_l, d_edge_all;
    CGPU_TRY(d_send.alloc(((size_t)per * nq + 1) * 8));
    CGPU_TRY(d_thr.alloc(n_thr * 8));
    CGPU_TRY(d_flag.alloc((size_t)(world + 1) * 8));
    CGPU_TRY(d_q.alloc(std::max<size_t>(1, (size_t)cm * S * nq) * sizeof(cgpu_rowsum)));
    CGPU_TRY(d_planes.alloc(3 * n_out * 8));
    CGPU_TRY(d_red.alloc(3 * (all_rows ? n_out : n_loc) * 8));
    std::vector<double> h_send((size_t)per * nq + 1, 1.0e300);
    h_send[(size_t)per * nq] = active ? 1.0 : 0.0;
    if (all_rows) {
        if (!thr || !out) {
            set_error(CGPU_EINVAL, "count_query_dist(all_rows) needs thresholds and an output block");
            return CGPU_EINVAL;
        }
        std::vector<double> h_all(n_thr, 1.0e300);
        memcpy(h_all.data(), thr, (size_t)cm * nq * 8);
        CGPU_CUDA(cudaMemcpyAsync(d_thr.p, h_all.data(), n_thr * 8, cudaMemcpyHostToDevice, st));
        CGPU_CUDA(cudaMemcpyAsync(d_send.as<double>() + (size_t)per * nq, &h_send[(size_t)per * nq], 8, cudaMemcpyHostToDevice, st));
        CGPU_CUDA(cudaStreamSynchronize(st));  // h_all is pageable: the copy has consumed it once this returns
        CGPU_TRY(comm_allgather_f64(d_send.as<double>() + (size_t)per * nq, d_flag.as<double>(), 1, st));
    } else {
        if (thr && nrows > 0) memcpy(h_send.data(), thr, (size_t)nrows * nq * 8);
        CGPU_CUDA(cudaMemcpyAsync(d_send.p, h_send.data(), h_send.size() * 8, cudaMemcpyHostToDevice, st));
        CGPU_TRY(comm_group_start());
        int rc1 = comm_allgather_f64(d_send.as<double>(), d_thr.as<double>(), (size_t)per * nq, st);
        int rc2 = comm_allgather_f64(d_send.as<double>() + (size_t)per * nq, d_flag.as<double>(), 1, st);
        CGPU_TRY(comm_group_end());
        CGPU_TRY(rc1);
        CGPU_TRY(rc2);
    }
    // every item over this rank's columns
    if (S <= 32) {
        long long maxB = 1;
        for (int s = 0; s < S; s++) maxB = std::max<long long>(maxB, r->nb[s]);
        int nsplit = (int)std::min<long long>(16, std::max<long long>(1, (4LL * G().sm_count + (long long)cm * S - 1) / ((long long)cm * S)));
        nsplit = (int)std::min<long long>(nsplit, std::max<long long>(1, maxB / 1024));
        CGPU_TRY(d_part.alloc(std::max<size_t>(1, (size_t)cm * S * nq) * nsplit * sizeof(cgpu_rowsum)));
        CGPU_TRY(launch_rowsum_all_scales(r->d_stat, r->nb, S, cm, nq, d_thr.as<double>(), d_part.as<cgpu_rowsum>(),
                                          d_q.as<cgpu_rowsum>(), nsplit, st, !have_edges, 0));
    } else {
        for (int s = 0; s < S; s++) {
            RowSrc src{r->d_stat[s], r->nb[s], nullptr, nullptr, nullptr, nullptr, nullptr};
            CGPU_TRY(launch_rowsum(0, src, cm, r->nb[s], nq, d_thr.as<double>(), d_q.as<cgpu_rowsum>(), S, s, st));
        }
    }
    CGPU_TRY(launch_rowsum_planes(d_q.as<cgpu_rowsum>(), (size_t)cm * S * nq, n_out, d_planes.as<double>(), st));
    double *pl = d_planes.as<double>(), *rd = d_red.as<double>();
    const size_t n_red = all_rows ? n_out : n_loc;
    CGPU_TRY(comm_group_start());
    int rcs[3];
    for (int f = 0; f < 3; f++)
        rcs[f] = all_rows ? comm_allreduce_f64(pl + f * n_out, rd + f * n_red, n_out, f, st)
                          : comm_reduce_scatter_f64(pl + f * n_out, rd + f * n_red, n_loc, f, st);
    CGPU_TRY(comm_group_end());
    for (int f = 0; f < 3; f++) CGPU_TRY(rcs[f]);
    std::vector<double> h_edges_all;
    if (!have_edges) {  // once per statistics build: the two smallest / two largest values of every row, over all ranks
        const size_t ne = (size_t)cm * S * 4;
        CGPU_TRY(d_edge_l.alloc(std::max<size_t>(1, ne) * 8));
        CGPU_TRY(d_edge_all.alloc(std::max<size_t>(1, ne) * world * 8));
        CGPU_TRY(launch_rowsum_edges(d_q.as<cgpu_rowsum>(), (size_t)cm * S, nq, d_edge_l.as<double>(), st));
        CGPU_TRY(comm_allgather_f64(d_edge_l.as<double>(), d_edge_all.as<double>(), ne, st));
        h_edges_all.resize(ne * world);
        CGPU_CUDA(cudaMemcpyAsync(h_edges_all.data(), d_edge_all.p, ne * world * 8, cudaMemcpyDeviceToHost, st));
    }
    std::vector<double> h_red(3 * n_red), h_flag(world);
    CGPU_CUDA(cudaMemcpyAsync(h_red.data(), rd, 3 * n_red * 8, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaMemcpyAsync(h_flag.data(), d_flag.p, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    if (!have_edges) {
        const size_t nrs = (size_t)cm * S;
        r->h_edges.assign(nrs * 4, 0.0);
        for (size_t e = 0; e < nrs; e++) {
            double mn1 = CG_HUGE, mn2 = CG_HUGE, mx1 = -CG_HUGE, mx2 = -CG_HUGE;
            for (int p = 0; p < world; p++) {
                const double *c = &h_edges_all[((size_t)p * nrs + e) * 4];
                for (int t = 0; t < 2; t++) {
                    const double v = c[t];
                    if (v < mn1) { mn2 = mn1; mn1 = v; } else if (v < mn2) mn2 = v;
                    const double u = c[2 + t];
                    if (u > mx1) { mx2 = mx1; mx1 = u; } else if (u > mx2) mx2 = u;
                }
            }
            double *o = &r->h_edges[4 * e];
            o[0] = mn1; o[1] = mn2; o[2] = mx1; o[3] = mx2;
        }
    }
    if (any_active) {
        int any = 0;
        for (int p = 0; p < world; p++) any |= h_flag[p] != 0.0;
        *any_active = any;
    }
    const int rows_out = all_rows ? cm : nrows, base = all_rows ? 0 : row0;
    if (out)
        for (int i = 0; i < rows_out; i++)
            for (int s = 0; s < S; s++) {
                const double *c = &r->h_edges[4 * ((size_t)(base + i) * S + s)];
                for (int q = 0; q < nq; q++) {
                    const size_t e = ((size_t)i * S + s) * nq + q;
                    cgpu_rowsum &o = out[e];
                    o.cnt = h_red[e];
                    o.lo = h_red[n_red + e];
                    o.hi = h_red[2 * n_red + e];
                    o.min1 = c[0]; o.min2 = c[1]; o.max1 = c[2]; o.max2 = c[3];
                }
            }
    return CGPU_OK;
}

}  // extern "C"
