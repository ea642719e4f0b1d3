// consel's replicate-statistics loops on the device (src/consel.c, H4-H11 of SURVEY 2.3).
//
// Everything the AU/BP/KH/SH machinery needs from a row v[0..B) of replicate
// statistics is a "row summary" against a threshold t:
//     cnt = #{v <= t}, lo = max{v <= t}, hi = min{v > t}, the two smallest and the two
//     largest values
// from which cntdist2/cntdist3 (src/consel.c:1806-1860) are reproduced bit for bit
// (cgpu_smooth_count), so no row is ever sorted: sort_vec (src/consel.c:1466-1471,
// 80% of consel at K=15) becomes one streaming pass per query round.  All reductions
// used here (integer count, max, min) are exact and order independent, which is why
// the results do not depend on tiling or on how replicates are sharded over GPUs.
//
// Layouts: replicate block rep[k*ld + b] (tree-major, as in .rmt); statistic block
// stat[e*ld + b] (item-major, as in .rep).  These kernels are HBM-streaming: 8 bytes
// per (row, replicate) element per pass.
#include <algorithm>

#include "cgpu_internal.h"

namespace cgpu {

#define CG_HUGE 1.0e30  // src/consel.c:155

// ------------------------------------------------------------------ top-2 per replicate
// For column b: m1 = max_k x, a1 = first index attaining it, m2 = max_{k != a1} x, all
// with the reference's -1e30 floor, so max_{j != i} x_j == (i == a1 ? m2 : m1) exactly
// as the prefix/suffix scan of calcmaxs (src/consel.c:1894-1916) yields it.
// The K rows are cut into `ks` ranges (grid.y) so that a block of few replicates (B = 10^4 gives
// only 79 column tiles) still fills the SMs; every thread keeps four row loads in flight.  The
// partial triples are merged in row order by col_top2_merge_kernel, which keeps the "first index
// attaining the maximum" rule of the sequential scan (a later range only wins with v > m1).
__global__ void __launch_bounds__(128)
col_top2_kernel(const double *__restrict__ x, int K, int B, int ld, int rows_per, double *__restrict__ m1o,
                double *__restrict__ m2o, int *__restrict__ a1o) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int k0 = blockIdx.y * rows_per, k1 = min(K, k0 + rows_per);
    double m1 = -CG_HUGE, m2 = -CG_HUGE;
    int a1 = -1;
    auto take = [&](double v, int k) {
        if (v > m1) {
            m2 = m1;
            m1 = v;
            a1 = k;
        } else if (v > m2) {
            m2 = v;
        }
    };
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
        const double v0 = x[(size_t)k * ld + b], v1 = x[(size_t)(k + 1) * ld + b];
        const double v2 = x[(size_t)(k + 2) * ld + b], v3 = x[(size_t)(k + 3) * ld + b];
        take(v0, k);
        take(v1, k + 1);
        take(v2, k + 2);
        take(v3, k + 3);
    }
    for (; k < k1; k++) take(x[(size_t)k * ld + b], k);
    const size_t o = (size_t)blockIdx.y * B + b;
    m1o[o] = m1;
    m2o[o] = m2;
    a1o[o] = a1;
}

__global__ void __launch_bounds__(128)
col_top2_merge_kernel(const double *__restrict__ pm1, const double *__restrict__ pm2, const int *__restrict__ pa1, int ks,
                      int B, double *__restrict__ m1o, double *__restrict__ m2o, int *__restrict__ a1o) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double m1 = pm1[b], m2 = pm2[b];
    int a1 = pa1[b];
    for (int p = 1; p < ks; p++) {
        const double q1 = pm1[(size_t)p * B + b], q2 = pm2[(size_t)p * B + b];
        if (q1 > m1) {  // strictly greater: an equal maximum in a later range keeps the earlier index
            m2 = m1 > q2 ? m1 : q2;
            m1 = q1;
            a1 = pa1[(size_t)p * B + b];
        } else {
            m2 = m2 > q1 ? m2 : q1;
        }
    }
    m1o[b] = m1;
    m2o[b] = m2;
    a1o[b] = a1;
}

// repminmaxs (src/consel.c:1939-1959): stat[e][b] = min_{i in A_e} (max_{j!=i} x_j - x_i)
__global__ void __launch_bounds__(128)
stat_build_kernel(const double *__restrict__ x, int K, int B, int ld, const double *__restrict__ m1v,
                  const double *__restrict__ m2v, const int *__restrict__ a1v, int cm,
                  const int *__restrict__ ass_ptr, const int *__restrict__ ass_idx,
                  double *__restrict__ stat) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double m1 = m1v[b], m2 = m2v[b];
    const int a1 = a1v[b];
    for (int e = blockIdx.y; e < cm; e += gridDim.y) {
        double y = CG_HUGE;
        for (int q = ass_ptr[e]; q < ass_ptr[e + 1]; q++) {
            int i = ass_idx[q];
            double st = __dsub_rn(i == a1 ? m2 : m1, x[(size_t)i * ld + b]);
            if (st < y) y = st;
        }
        stat[(size_t)e * ld + b] = y;
    }
}

// ------------------------------------------------------------------ row summaries
struct RowAcc {
    unsigned long long cnt;
    double lo, hi;
};

__device__ __forceinline__ void edge_insert(double v, double &mn1, double &mn2, double &mx1, double &mx2) {
    if (v < mn1) { mn2 = mn1; mn1 = v; } else if (v < mn2) mn2 = v;
    if (v > mx1) { mx2 = mx1; mx1 = v; } else if (v > mx2) mx2 = v;
}
__device__ __forceinline__ void edge_merge(double &mn1, double &mn2, double &mx1, double &mx2, double o1,
                                           double o2, double p1, double p2) {
    // two smallest of {mn1<=mn2, o1<=o2}; two largest of {mx1>=mx2, p1>=p2}
    double a = fmin(mn1, o1), bb = fmax(mn1, o1);
    mn2 = fmin(bb, fmin(mn2, o2));
    mn1 = a;
    double c = fmax(mx1, p1), d = fmin(mx1, p1);
    mx2 = fmax(d, fmax(mx2, p2));
    mx1 = c;
}

constexpr int RS_THREADS = 256;
constexpr int RS_MAXQ = 16;

template <int MODE>
__device__ __forceinline__ double row_value(const RowSrc &s, int row, int b, int i_row, int j_row) {
    if (MODE == 0) return s.x[(size_t)row * s.ld + b];
    if (MODE == 1) return __dsub_rn(s.x[(size_t)j_row * s.ld + b], s.x[(size_t)i_row * s.ld + b]);
    return __dsub_rn(i_row == s.a1[b] ? s.m2[b] : s.m1[b], s.x[(size_t)i_row * s.ld + b]);
}

// one CTA per row; nq thresholds per row: thr[row*nq + q]; out[(row*out_stride + out_off)*nq + q]
// QCAP: compile-time capacity of the per-thread summaries (1 for the KH/SH tests, whose rows
// have one threshold: 48 registers and full occupancy instead of ~200 and 12 %)
template <int MODE, int QCAP>
__global__ void __launch_bounds__(RS_THREADS)
rowsum_kernel(RowSrc src, int B, int nq, const double *__restrict__ thr, cgpu_rowsum *__restrict__ out,
              int out_stride, int out_off) {
    const int row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ double s_thr[QCAP];
    __shared__ unsigned long long s_cnt[RS_THREADS / 32][QCAP];
    __shared__ double s_lo[RS_THREADS / 32][QCAP], s_hi[RS_THREADS / 32][QCAP];
    __shared__ double s_edge[RS_THREADS / 32][4];
    if (tid < nq) s_thr[tid] = thr[(size_t)row * nq + tid];
    __syncthreads();
    const int i_row = (MODE == 0) ? 0 : src.ia[row];
    const int j_row = (MODE == 1) ? src.ja[row] : 0;

    unsigned long long cnt[QCAP];
    double lo[QCAP], hi[QCAP];
#pragma unroll
    for (int q = 0; q < QCAP; q++) {
        cnt[q] = 0;
        lo[q] = -CG_HUGE;
        hi[q] = CG_HUGE;
    }
    double mn1 = CG_HUGE, mn2 = CG_HUGE, mx1 = -CG_HUGE, mx2 = -CG_HUGE;
    auto take = [&](double v) {
        edge_insert(v, mn1, mn2, mx1, mx2);
#pragma unroll
        for (int q = 0; q < QCAP; q++) {
            if (q < nq) {
                const double t = s_thr[q];
                if (v <= t) {
                    cnt[q]++;
                    lo[q] = fmax(lo[q], v);
                } else {
                    hi[q] = fmin(hi[q], v);
                }
            }
        }
    };
    int b = tid;
    for (; b + 3 * RS_THREADS < B; b += 4 * RS_THREADS) {  // four rows of loads in flight
        const double v0 = row_value<MODE>(src, row, b, i_row, j_row);
        const double v1 = row_value<MODE>(src, row, b + RS_THREADS, i_row, j_row);
        const double v2 = row_value<MODE>(src, row, b + 2 * RS_THREADS, i_row, j_row);
        const double v3 = row_value<MODE>(src, row, b + 3 * RS_THREADS, i_row, j_row);
        take(v0);
        take(v1);
        take(v2);
        take(v3);
    }
    for (; b < B; b += RS_THREADS) take(row_value<MODE>(src, row, b, i_row, j_row));
    // warp reduce, then across warps through shared memory
#pragma unroll
    for (int off = 16; off; off >>= 1) {
#pragma unroll
        for (int q = 0; q < QCAP; q++) {
            if (q < nq) {
                cnt[q] += __shfl_xor_sync(0xffffffffu, cnt[q], off);
                lo[q] = fmax(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], off));
                hi[q] = fmin(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], off));
            }
        }
        double o1 = __shfl_xor_sync(0xffffffffu, mn1, off), o2 = __shfl_xor_sync(0xffffffffu, mn2, off);
        double p1 = __shfl_xor_sync(0xffffffffu, mx1, off), p2 = __shfl_xor_sync(0xffffffffu, mx2, off);
        edge_merge(mn1, mn2, mx1, mx2, o1, o2, p1, p2);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < QCAP; q++)
            if (q < nq) {
                s_cnt[wid][q] = cnt[q];
                s_lo[wid][q] = lo[q];
                s_hi[wid][q] = hi[q];
            }
        s_edge[wid][0] = mn1; s_edge[wid][1] = mn2; s_edge[wid][2] = mx1; s_edge[wid][3] = mx2;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < RS_THREADS / 32; w++)
            edge_merge(mn1, mn2, mx1, mx2, s_edge[w][0], s_edge[w][1], s_edge[w][2], s_edge[w][3]);
        s_edge[0][0] = mn1; s_edge[0][1] = mn2; s_edge[0][2] = mx1; s_edge[0][3] = mx2;
    }
    __syncthreads();
    if (tid < nq) {
        unsigned long long c = 0;
        double l = -CG_HUGE, h = CG_HUGE;
        for (int w = 0; w < RS_THREADS / 32; w++) {
            c += s_cnt[w][tid];
            l = fmax(l, s_lo[w][tid]);
            h = fmin(h, s_hi[w][tid]);
        }
        cgpu_rowsum r;
        r.cnt = (double)c;
        r.lo = l;
        r.hi = h;
        r.min1 = s_edge[0][0]; r.min2 = s_edge[0][1]; r.max1 = s_edge[0][2]; r.max2 = s_edge[0][3];
        out[((size_t)row * out_stride + out_off) * nq + tid] = r;
    }
}

// All scales of a count query in ONE launch: grid (row, scale, split).  Rows are cut into
// `nsplit` column ranges so that a query over few long rows (K=105, B=100000) still
// fills the machine; the partial summaries are merged by rowsum_merge_kernel (count: sum,
// lo: max, hi: min, edges: two smallest / two largest — all exact).
struct ScaleList {
    const double *stat[32];
    int B[32];
};

// NQ (compile time) >= nq thresholds; surplus thresholds are +HUGE and never stored.  Each
// thread keeps four independent 8-byte loads in flight: the kernel is a pure stream over
// the statistic blocks (8 B per row element and pass) and must be HBM bound.
// EDGES: also track the two smallest / two largest values of the row (needed once per
// statistics build; later queries take them from the host-side cache and skip the work)
template <int NQ, bool EDGES>
__global__ void __launch_bounds__(RS_THREADS)
rowsum_multi_kernel(ScaleList sl, int row0, int nq, int q0, int nqt, int nsplit, const double *__restrict__ thr,
                    cgpu_rowsum *__restrict__ part) {
    // thresholds q0 .. q0+nq-1 of the nqt thresholds of each row; rows row0 .. row0+gridDim.x-1 of
    // the statistic blocks (thresholds and results are indexed from 0)
    const int row = blockIdx.x, sc = blockIdx.y, sp = blockIdx.z, S = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int B = sl.B[sc];
    const double *__restrict__ v = sl.stat[sc] + (size_t)(row0 + row) * B;
    const int chunk = (B + nsplit - 1) / nsplit;
    const int b_lo = sp * chunk, b_hi = min(B, b_lo + chunk);
    __shared__ unsigned long long s_cnt[RS_THREADS / 32][NQ];
    __shared__ double s_lo[RS_THREADS / 32][NQ], s_hi[RS_THREADS / 32][NQ];
    __shared__ double s_edge[RS_THREADS / 32][4];
    double t[NQ];
    unsigned int cnt[NQ];
    double lo[NQ], hi[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        t[q] = q < nq ? thr[(size_t)row * nqt + q0 + q] : CG_HUGE;
        cnt[q] = 0;
        lo[q] = -CG_HUGE;
        hi[q] = CG_HUGE;
    }
    double mn1 = CG_HUGE, mn2 = CG_HUGE, mx1 = -CG_HUGE, mx2 = -CG_HUGE;
    auto take = [&](double x) {
        if (EDGES) edge_insert(x, mn1, mn2, mx1, mx2);
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            if (x <= t[q]) {
                cnt[q]++;
                lo[q] = fmax(lo[q], x);
            } else {
                hi[q] = fmin(hi[q], x);
            }
        }
    };
    constexpr int U = 4;
    int b = b_lo + tid;
    for (; b + (U - 1) * RS_THREADS < b_hi; b += U * RS_THREADS) {
        double x[U];
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = v[b + u * RS_THREADS];
#pragma unroll
        for (int u = 0; u < U; u++) take(x[u]);
    }
    for (; b < b_hi; b += RS_THREADS) take(v[b]);

    unsigned long long c64[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) c64[q] = cnt[q];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            c64[q] += __shfl_xor_sync(0xffffffffu, c64[q], off);
            lo[q] = fmax(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], off));
            hi[q] = fmin(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], off));
        }
        double o1 = __shfl_xor_sync(0xffffffffu, mn1, off), o2 = __shfl_xor_sync(0xffffffffu, mn2, off);
        double p1 = __shfl_xor_sync(0xffffffffu, mx1, off), p2 = __shfl_xor_sync(0xffffffffu, mx2, off);
        edge_merge(mn1, mn2, mx1, mx2, o1, o2, p1, p2);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            s_cnt[wid][q] = c64[q];
            s_lo[wid][q] = lo[q];
            s_hi[wid][q] = hi[q];
        }
        s_edge[wid][0] = mn1; s_edge[wid][1] = mn2; s_edge[wid][2] = mx1; s_edge[wid][3] = mx2;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < RS_THREADS / 32; w++)
            edge_merge(mn1, mn2, mx1, mx2, s_edge[w][0], s_edge[w][1], s_edge[w][2], s_edge[w][3]);
        s_edge[0][0] = mn1; s_edge[0][1] = mn2; s_edge[0][2] = mx1; s_edge[0][3] = mx2;
    }
    __syncthreads();
    if (tid < nq) {
        unsigned long long c = 0;
        double l = -CG_HUGE, h = CG_HUGE;
        for (int w = 0; w < RS_THREADS / 32; w++) {
            c += s_cnt[w][tid];
            l = fmax(l, s_lo[w][tid]);
            h = fmin(h, s_hi[w][tid]);
        }
        cgpu_rowsum r;
        r.cnt = (double)c;
        r.lo = l;
        r.hi = h;
        r.min1 = s_edge[0][0]; r.min2 = s_edge[0][1]; r.max1 = s_edge[0][2]; r.max2 = s_edge[0][3];
        part[(((size_t)row * S + sc) * nsplit + sp) * nqt + q0 + tid] = r;
    }
}

// Several thresholds per row, sorted: an element falls into exactly one interval
// (t_(j-1), t_j] of the sorted thresholds, j = #{q : x > t_q}.  Each thread keeps a private
// (count, min, max) per interval in shared memory and every threshold's (count, lo, hi) is
// a prefix sum / prefix max / suffix min over the intervals afterwards: NQ compares and
// three updates per element instead of NQ x (compare, count, max, min).  Exact as before
// (integer counts, max and min are order independent).  Row edges are never needed here
// (they come from the host cache).
template <int NQ>
__global__ void __launch_bounds__(RS_THREADS)
rowsum_sorted_kernel(ScaleList sl, int row0, int nq, int q0, int nqt, int nsplit, const double *__restrict__ thr,
                     cgpu_rowsum *__restrict__ part) {
    const int row = blockIdx.x, sc = blockIdx.y, sp = blockIdx.z, S = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int B = sl.B[sc];
    const double *__restrict__ v = sl.stat[sc] + (size_t)(row0 + row) * B;
    const int chunk = (B + nsplit - 1) / nsplit;
    const int b_lo = sp * chunk, b_hi = min(B, b_lo + chunk);
    constexpr int NI = NQ + 1;  // intervals
    extern __shared__ __align__(16) unsigned char rss_raw[];
    double *imin = reinterpret_cast<double *>(rss_raw);                 // [NI][RS_THREADS]
    double *imax = imin + NI * RS_THREADS;                              // [NI][RS_THREADS]
    unsigned int *icnt = reinterpret_cast<unsigned int *>(imax + NI * RS_THREADS);  // [NI][RS_THREADS]
    __shared__ double s_t[NQ];
    __shared__ int s_perm[NQ];
    __shared__ unsigned long long s_cnt[RS_THREADS / 32][NQ];
    __shared__ double s_lo[RS_THREADS / 32][NQ], s_hi[RS_THREADS / 32][NQ];
    if (tid == 0) {  // insertion sort of <= 8 thresholds, remembering where each came from
        for (int q = 0; q < NQ; q++) {
            const double t = q < nq ? thr[(size_t)row * nqt + q0 + q] : CG_HUGE;
            int p = q;
            while (p > 0 && s_t[p - 1] > t) {
                s_t[p] = s_t[p - 1];
                s_perm[p] = s_perm[p - 1];
                p--;
            }
            s_t[p] = t;
            s_perm[p] = q;
        }
    }
#pragma unroll
    for (int j = 0; j < NI; j++) {
        imin[j * RS_THREADS + tid] = CG_HUGE;
        imax[j * RS_THREADS + tid] = -CG_HUGE;
        icnt[j * RS_THREADS + tid] = 0;
    }
    __syncthreads();
    double t[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) t[q] = s_t[q];
    auto take = [&](double x) {
        if (x != x) return;  // NaN never counts and never becomes a neighbour (as fmax/fmin ignore it)
        int j = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) j += x > t[q] ? 1 : 0;
        const int e = j * RS_THREADS + tid;
        icnt[e]++;
        imin[e] = fmin(imin[e], x);
        imax[e] = fmax(imax[e], x);
    };
    constexpr int U = 4;
    int b = b_lo + tid;
    for (; b + (U - 1) * RS_THREADS < b_hi; b += U * RS_THREADS) {
        double x[U];
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = v[b + u * RS_THREADS];
#pragma unroll
        for (int u = 0; u < U; u++) take(x[u]);
    }
    for (; b < b_hi; b += RS_THREADS) take(v[b]);

    // per thread: threshold q (sorted order) sees intervals 0..q below it and q+1..NQ above it
    unsigned long long c64[NQ];
    double lo[NQ], hi[NQ];
    {
        unsigned long long c = 0;
        double l = -CG_HUGE;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            c += icnt[q * RS_THREADS + tid];
            l = fmax(l, imax[q * RS_THREADS + tid]);
            c64[q] = c;
            lo[q] = l;
        }
        double h = CG_HUGE;
#pragma unroll
        for (int q = NQ - 1; q >= 0; q--) {
            h = fmin(h, imin[(q + 1) * RS_THREADS + tid]);
            hi[q] = h;
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            c64[q] += __shfl_xor_sync(0xffffffffu, c64[q], off);
            lo[q] = fmax(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], off));
            hi[q] = fmin(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], off));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            s_cnt[wid][q] = c64[q];
            s_lo[wid][q] = lo[q];
            s_hi[wid][q] = hi[q];
        }
    }
    __syncthreads();
    if (tid < NQ && s_perm[tid] < nq) {
        unsigned long long c = 0;
        double l = -CG_HUGE, h = CG_HUGE;
        for (int w = 0; w < RS_THREADS / 32; w++) {
            c += s_cnt[w][tid];
            l = fmax(l, s_lo[w][tid]);
            h = fmin(h, s_hi[w][tid]);
        }
        cgpu_rowsum r;
        r.cnt = (double)c;
        r.lo = l;
        r.hi = h;
        r.min1 = r.min2 = r.max1 = r.max2 = 0.0;  // filled from the host-side edge cache
        part[(((size_t)row * S + sc) * nsplit + sp) * nqt + q0 + s_perm[tid]] = r;
    }
}

__global__ void __launch_bounds__(128)
rowsum_merge_kernel(const cgpu_rowsum *__restrict__ part, int total, int nsplit, int nq, cgpu_rowsum *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // (row*S + sc)*nq + q
    if (e >= total) return;
    const int rs = e / nq, q = e % nq;
    cgpu_rowsum r = part[((size_t)rs * nsplit) * nq + q];
    for (int sp = 1; sp < nsplit; sp++) {
        const cgpu_rowsum p = part[((size_t)rs * nsplit + sp) * nq + q];
        r.cnt += p.cnt;  // integers below 2^53: exact
        r.lo = fmax(r.lo, p.lo);
        r.hi = fmin(r.hi, p.hi);
        edge_merge(r.min1, r.min2, r.max1, r.max2, p.min1, p.min2, p.max1, p.max2);
    }
    out[e] = r;
}

// row summaries -> three planar arrays (count | lo | hi) of `padded` entries each, the layout the
// NCCL reductions (sum | max | min) of a sharded count query want; padding entries are neutral
__global__ void __launch_bounds__(256)
rowsum_planes_kernel(const cgpu_rowsum *__restrict__ q, size_t n, size_t padded, double *__restrict__ planes) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= padded) return;
    double c = 0.0, lo = -CG_HUGE, hi = CG_HUGE;
    if (e < n) {
        c = q[e].cnt;
        lo = q[e].lo;
        hi = q[e].hi;
    }
    planes[e] = c;
    planes[padded + e] = lo;
    planes[2 * padded + e] = hi;
}
// (min1, min2, max1, max2) of every (row, scale), taken from the first threshold's summary
__global__ void __launch_bounds__(256)
rowsum_edges_kernel(const cgpu_rowsum *__restrict__ q, size_t nrs, int nq, double *__restrict__ edges) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= nrs) return;
    const cgpu_rowsum r = q[e * nq];
    edges[4 * e] = r.min1;
    edges[4 * e + 1] = r.min2;
    edges[4 * e + 2] = r.max1;
    edges[4 * e + 3] = r.max2;
}
int launch_rowsum_planes(const cgpu_rowsum *d_q, size_t n, size_t padded, double *d_planes, cudaStream_t st) {
    if (padded == 0) return CGPU_OK;
    rowsum_planes_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, st>>>(d_q, n, padded, d_planes);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}
int launch_rowsum_edges(const cgpu_rowsum *d_q, size_t nrs, int nq, double *d_edges, cudaStream_t st) {
    if (nrs == 0) return CGPU_OK;
    rowsum_edges_kernel<<<(unsigned)((nrs + 255) / 256), 256, 0, st>>>(d_q, nrs, nq, d_edges);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

static size_t rss_smem(int nq) { return (size_t)(nq + 1) * RS_THREADS * (8 + 8 + 4); }

int launch_rowsum_all_scales(const double *const *d_stat, const int *B, int S, int cm, int nq, const double *d_thr,
                             cgpu_rowsum *d_part, cgpu_rowsum *d_out, int nsplit, cudaStream_t st, bool edges, int row0) {
    if (cm <= 0 || S <= 0) return CGPU_OK;
    if (S > 32 || nq < 1 || nq > RS_MAXQ) {
        set_error(CGPU_EINVAL, "count query with %d scales / %d thresholds", S, nq);
        return CGPU_EINVAL;
    }
    ScaleList sl;
    for (int s = 0; s < S; s++) {
        sl.stat[s] = d_stat[s];
        sl.B[s] = B[s];
    }
    const dim3 grid(cm, S, nsplit);
    double sumB = 0;
    for (int sc = 0; sc < S; sc++) sumB += B[sc];
    KT kt(nq == 1 ? "rowsum_multi_kernel<1>" : "rowsum_multi/sorted_kernel", st, 8.0 * cm * sumB * ((nq + 7) / 8));  // one pass per 8 thresholds
    // more than 8 thresholds per row: two passes of <= 8 (64 registers, 3 CTAs/SM) stream faster
    // than one pass of 16 (204 registers)
#define CGPU_RSM(NQ_)                                                                                                     \
    do {                                                                                                                  \
        if (edges) rowsum_multi_kernel<NQ_, true><<<grid, RS_THREADS, 0, st>>>(sl, row0, n, q0, nq, nsplit, d_thr, d_part);     \
        else rowsum_multi_kernel<NQ_, false><<<grid, RS_THREADS, 0, st>>>(sl, row0, n, q0, nq, nsplit, d_thr, d_part);          \
    } while (0)
    for (int q0 = 0; q0 < nq; q0 += 8) {
        const int n = std::min(8, nq - q0);
        if (n <= 1) CGPU_RSM(1);
        else if (n <= 2) CGPU_RSM(2);
        else if (!edges && n <= 4) rowsum_sorted_kernel<4><<<grid, RS_THREADS, rss_smem(4), st>>>(sl, row0, n, q0, nq, nsplit, d_thr, d_part);
        else if (!edges) rowsum_sorted_kernel<8><<<grid, RS_THREADS, rss_smem(8), st>>>(sl, row0, n, q0, nq, nsplit, d_thr, d_part);
        else if (n <= 4) CGPU_RSM(4);
        else CGPU_RSM(8);
    }
#undef CGPU_RSM
    CGPU_LAUNCH_CHECK();
    const int total = cm * S * nq;
    rowsum_merge_kernel<<<(total + 127) / 128, 128, 0, st>>>(d_part, total, nsplit, nq, d_out);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

int launch_rowsum(int mode, const RowSrc &src, int rows, int B, int nq, const double *d_thr, cgpu_rowsum *d_out,
                  int out_stride, int out_off, cudaStream_t st) {
    if (rows <= 0) return CGPU_OK;
    if (nq < 1 || nq > RS_MAXQ) {
        set_error(CGPU_EINVAL, "nq=%d (1..%d)", nq, RS_MAXQ);
        return CGPU_EINVAL;
    }
    KT kt(mode == 0 ? "rowsum_kernel<plain>" : mode == 1 ? "rowsum_kernel<kh>" : "rowsum_kernel<sh>", st, (mode == 1 ? 16.0 : 8.0) * rows * B);
#define CGPU_RS(M_)                                                                                                  \
    do {                                                                                                             \
        if (nq == 1) rowsum_kernel<M_, 1><<<rows, RS_THREADS, 0, st>>>(src, B, nq, d_thr, d_out, out_stride, out_off);        \
        else rowsum_kernel<M_, RS_MAXQ><<<rows, RS_THREADS, 0, st>>>(src, B, nq, d_thr, d_out, out_stride, out_off);          \
    } while (0)
    if (mode == 0) CGPU_RS(0);
    else if (mode == 1) CGPU_RS(1);
    else CGPU_RS(2);
#undef CGPU_RS
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// ------------------------------------------------------------------ order statistic
// k-th smallest (0-based) of each row on the order-preserving integer image of a double.
__device__ __forceinline__ unsigned long long dkey(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// Radix select, most significant byte first: eight passes, each a 256-bin shared-memory
// histogram of the current byte over the elements that match the prefix found so far, then a
// scan of the bins for the one that holds rank k (rows stay in L1/L2 between the passes).
__global__ void __launch_bounds__(256)
order_stat_kernel(const double *__restrict__ stat, int B, int ld, const int *__restrict__ rank,
                  double *__restrict__ val) {
    const int row = blockIdx.x, tid = threadIdx.x;
    const double *v = stat + (size_t)row * ld;
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_k;
    if (tid == 0) {
        s_prefix = 0;
        s_k = (unsigned int)rank[row];
    }
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        s_hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int b = tid; b < B; b += 256) {
            const unsigned long long key = dkey(v[b]);
            if ((key & mask) == prefix) atomicAdd(&s_hist[(unsigned int)(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int k = s_k, c = 0;
            int bin = 0;
            for (; bin < 255; bin++) {
                c = s_hist[bin];
                if (k < c) break;
                k -= c;
            }
            s_k = k;
            s_prefix = prefix | ((unsigned long long)bin << shift);
        }
        __syncthreads();
    }
    if (tid == 0) val[row] = dunkey(s_prefix);
}

int launch_order_stat(const double *d_stat, int rows, int B, int ld, const int *d_rank, double *d_val,
                      cudaStream_t st) {
    if (rows <= 0) return CGPU_OK;
    KT kt("order_stat_kernel", st, 8.0 * rows * B);
    order_stat_kernel<<<rows, 256, 0, st>>>(d_stat, B, ld, d_rank, d_val);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// ------------------------------------------------------------------ mc-test pieces
// centring (src/consel.c:1378-1382): one warp per tree; the row sum is accumulated in
// the reference's order (b ascending, one rounded add per element): lanes fetch 32
// consecutive values, every lane replays the 32 adds through shuffles.
constexpr int CR_CHUNK = 256;  // elements staged per warp and step
__global__ void __launch_bounds__(128)
center_rows_kernel(const double *__restrict__ x, int K, int B, int ld, double *__restrict__ xc, bool rescale, double rs) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + w;
    __shared__ double stage[4][CR_CHUNK];
    if (row >= K) return;
    const double *src = x + (size_t)row * ld;
    double *buf = stage[w];
    // The row sum is one chain of rounded adds in replicate order.  The next 256 values are
    // fetched (coalesced, 8 per lane) while every lane replays the chain over the staged ones
    // from shared memory: the only serial latency left is the add itself.
    double nxt[CR_CHUNK / 32];
#pragma unroll
    for (int u = 0; u < CR_CHUNK / 32; u++) {
        const int b = u * 32 + lane;
        nxt[u] = b < B ? src[b] : 0.0;
    }
    double s = 0.0;
    for (int b0 = 0; b0 < B; b0 += CR_CHUNK) {
        __syncwarp();
#pragma unroll
        for (int u = 0; u < CR_CHUNK / 32; u++) buf[u * 32 + lane] = nxt[u];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < CR_CHUNK / 32; u++) {
            const int b = b0 + CR_CHUNK + u * 32 + lane;
            nxt[u] = b < B ? src[b] : 0.0;
        }
        const int m = min(CR_CHUNK, B - b0);
        if (m == CR_CHUNK) {
#pragma unroll 32
            for (int t = 0; t < CR_CHUNK; t++) s = __dadd_rn(s, buf[t]);
        } else {
            for (int t = 0; t < m; t++) s = __dadd_rn(s, buf[t]);
        }
    }
    const double mean = __ddiv_rn(s, (double)B);
    double *dst = xc + (size_t)row * ld;
    if (rescale) {  // repminmaxsrs (src/consel.c:1978): rs*(x - mean) + mean, three roundings
        for (int b = lane; b < B; b += 32) dst[b] = __dadd_rn(__dmul_rn(rs, __dsub_rn(src[b], mean)), mean);
    } else {
        for (int b = lane; b < B; b += 32) dst[b] = __dsub_rn(src[b], mean);
    }
}

// pairwise variances -> weights (src/consel.c:1404-1415):
//   wt[i][j] = 1/sqrt((sum_b (x_i - x_j)^2 + varadd)/(B-1)), sum in replicate order.
// 16x16 pairs per CTA, one pair per thread, replicates staged through shared memory.
constexpr int PV_T = 32, PV_BC = 64;  // 32 x 32 pairs per CTA (256 threads, 2 x 2 pairs each), 64 replicates per stage
// rows == nullptr: all pairs, lower triangle computed once and mirrored (grid t x t).
// rows != nullptr: only the weight rows i = rows[0..nrows) against every j (grid t x ceil(nrows/32)):
// a rank of a sharded run needs the rows of its own members only.  (x_i - x_j)^2 is the same
// value whichever of the two is "i", so both forms give bit-identical weights.
// Every thread owns 2 x 2 pairs: four shared-memory loads feed twelve fp64 operations per
// replicate (one pair per thread was bound by its two loads per three operations).
__global__ void __launch_bounds__(256)
pair_weight_kernel(const double *__restrict__ xc, int K, int B, int ld, double varadd, double *__restrict__ wt,
                   int *__restrict__ nsmall, const int *__restrict__ rows, int nrows) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (!rows && tj > ti) return;
    __shared__ double si[PV_T][PV_BC + 1], sj[PV_T][PV_BC + 1];
    __shared__ int s_row[PV_T];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // pairs (row ty / ty+16 of tile ti) x (tx / tx+16 of tile tj)
    if (threadIdx.x < PV_T) {
        const int r = ti * PV_T + threadIdx.x;
        s_row[threadIdx.x] = rows ? (r < nrows ? rows[r] : -1) : (r < K ? r : -1);
    }
    __syncthreads();
    // staging: thread e loads rows r = e/64 (+4, +8, ... +28) at column e%64 of both tiles; the next
    // chunk is fetched into registers while the current one is consumed from shared memory
    constexpr int PER = PV_T * PV_BC / 256;  // 8 elements per thread and tile
    const int lc = threadIdx.x % PV_BC, lr = threadIdx.x / PV_BC;
    double ni[PER], nj[PER];
    auto fetch = [&](int b0) {
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int r = lr + u * (256 / PV_BC);
            const int b = b0 + lc;
            const int gi = s_row[r], gj = tj * PV_T + r;
            ni[u] = (gi >= 0 && b < B) ? xc[(size_t)gi * ld + b] : 0.0;
            nj[u] = (gj < K && b < B) ? xc[(size_t)gj * ld + b] : 0.0;
        }
    };
    fetch(0);
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int b0 = 0; b0 < B; b0 += PV_BC) {
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int r = lr + u * (256 / PV_BC);
            si[r][lc] = ni[u];
            sj[r][lc] = nj[u];
        }
        __syncthreads();
        fetch(b0 + PV_BC);
        const int m = min(PV_BC, B - b0);
        // per pair: the chain of rounded adds in replicate order (the squares are independent)
#pragma unroll 8
        for (int c = 0; c < m; c++) {
            const double a0 = si[ty][c], a1 = si[ty + 16][c], b0v = sj[tx][c], b1v = sj[tx + 16][c];
            double d;
            d = __dsub_rn(a0, b0v); acc[0][0] = __dadd_rn(acc[0][0], __dmul_rn(d, d));
            d = __dsub_rn(a0, b1v); acc[0][1] = __dadd_rn(acc[0][1], __dmul_rn(d, d));
            d = __dsub_rn(a1, b0v); acc[1][0] = __dadd_rn(acc[1][0], __dmul_rn(d, d));
            d = __dsub_rn(a1, b1v); acc[1][1] = __dadd_rn(acc[1][1], __dmul_rn(d, d));
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = s_row[ty + 16 * p], j = tj * PV_T + tx + 16 * q;
            if (i < 0 || j >= K) continue;
            if (i == j) {
                wt[(size_t)i * K + i] = 0.0;
                continue;
            }
            if (!rows && j > i) continue;
            // VARADDWARN (src/consel.c:154,1409) counts every pair once: in the row-subset form the pair is
            // counted by the row with the larger tree index -- the caller only uses the count on single-device runs
            if (acc[p][q] < varadd * 100.0 && (!rows || j < i)) atomicAdd(nsmall, 1);
            const double w = __ddiv_rn(1.0, __dsqrt_rn(__ddiv_rn(__dadd_rn(acc[p][q], varadd), (double)(B - 1))));
            wt[(size_t)i * K + j] = w;
            if (!rows) wt[(size_t)j * K + i] = w;
        }
}

// SH statistic rows for general association sets (src/consel.c:2026-2030, 2017-2024):
//   unweighted: out[q][b] = max_{j in cass(e)} x[j][b] - x[i_q][b]
//   weighted:   out[q][b] = max_{j in cass(e)} wt[i_q][j] * (x[j][b] - x[i_q][b])
// q runs over the members (CSR positions) of the items; comp marks complement by item.
template <bool WEIGHTED>
__global__ void __launch_bounds__(128)
sh_rows_kernel(const double *__restrict__ x, int K, int B, int ld, const double *__restrict__ wt,
               const int *__restrict__ mem_item, const int *__restrict__ mem_tree,
               const int *__restrict__ cass_ptr, const int *__restrict__ cass_idx, double *__restrict__ out) {
    const int q = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int e = mem_item[q], i = mem_tree[q];
    const double x0 = x[(size_t)i * ld + b];
    double mx = -CG_HUGE;
    if (WEIGHTED) {
        const double *wp = wt + (size_t)i * K;
        for (int p = cass_ptr[e]; p < cass_ptr[e + 1]; p++) {
            int j = cass_idx[p];
            double y = __dmul_rn(wp[j], __dsub_rn(x[(size_t)j * ld + b], x0));
            if (y > mx) mx = y;
        }
        out[(size_t)q * ld + b] = mx;
    } else {
        for (int p = cass_ptr[e]; p < cass_ptr[e + 1]; p++) {
            double y = x[(size_t)cass_idx[p] * ld + b];
            if (y > mx) mx = y;
        }
        out[(size_t)q * ld + b] = __dsub_rn(mx, x0);
    }
}

// Weighted SH rows when every item is a single tree (the default identity association): the
// complement of member i is "every j != i", so no index list is needed, and a CTA finishes
// WS_MI members for 128 replicates at once -- x[j][b] is loaded once per j and used by all of its
// members, the weights wt[i_m][j] of a chunk of j sit in shared memory.  Per (member, j, b) one
// rounded subtract, one rounded multiply and a compare, exactly the operations (and the ascending
// j order) of calcmaxass with weights (src/consel.c:2017-2024); the FP64 pipe is the bound.
constexpr int WS_MI = 8, WS_JC = 128;
__global__ void __launch_bounds__(128)
wsh_single_kernel(const double *__restrict__ x, int K, int B, int ld, const double *__restrict__ wt,
                  const int *__restrict__ mem_tree, int nmem, double *__restrict__ out) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    const int q0 = blockIdx.y * WS_MI;
    __shared__ __align__(16) double s_w[WS_JC][WS_MI];  // the WS_MI weights of one j are contiguous: 128-bit shared loads
    __shared__ int s_tree[WS_MI];
    if (threadIdx.x < WS_MI) s_tree[threadIdx.x] = q0 + threadIdx.x < nmem ? mem_tree[q0 + threadIdx.x] : -1;
    __syncthreads();
    const bool live = b < B;
    double x0[WS_MI], mx[WS_MI];
#pragma unroll
    for (int m = 0; m < WS_MI; m++) {
        const int t = s_tree[m];
        x0[m] = (live && t >= 0) ? x[(size_t)t * ld + b] : 0.0;
        mx[m] = -CG_HUGE;
    }
    for (int j0 = 0; j0 < K; j0 += WS_JC) {
        const int jn = min(WS_JC, K - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < WS_MI * WS_JC; e += 128) {
            const int m = e / WS_JC, jj = e % WS_JC, t = s_tree[m];
            // the member's own tree is not in its complement: a NaN weight fails every compare
            s_w[jj][m] = (t >= 0 && jj < jn && j0 + jj != t) ? wt[(size_t)t * K + j0 + jj] : __longlong_as_double(0x7ff8000000000000ll);
        }
        __syncthreads();
        if (!live) continue;
        double xn = x[(size_t)j0 * ld + b];
        for (int jj = 0; jj < jn; jj++) {
            const double xj = xn;
            if (jj + 1 < jn) xn = x[(size_t)(j0 + jj + 1) * ld + b];
            double w[WS_MI];
#pragma unroll
            for (int m = 0; m < WS_MI; m += 2) {
                const double2 w2 = *reinterpret_cast<const double2 *>(&s_w[jj][m]);
                w[m] = w2.x;
                w[m + 1] = w2.y;
            }
#pragma unroll
            for (int m = 0; m < WS_MI; m++) {
                const double y = __dmul_rn(w[m], __dsub_rn(xj, x0[m]));
                if (y > mx[m]) mx[m] = y;
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int m = 0; m < WS_MI; m++)
        if (q0 + m < nmem) out[(size_t)(q0 + m) * ld + b] = mx[m];
}

// observed statistics and KH contrasts that depend on the weights (src/consel.c:2121-2134,
// 2017-2024 applied to datvec): one thread per member.
__global__ void __launch_bounds__(128)
weighted_obs_kernel(const double *__restrict__ dat, int K, const double *__restrict__ wt, int nmem,
                    const int *__restrict__ mem_item, const int *__restrict__ mem_tree,
                    const int *__restrict__ cass_ptr, const int *__restrict__ cass_idx,
                    double *__restrict__ obs, int *__restrict__ argj) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nmem) return;
    const int i = mem_tree[q];
    const double x0 = dat[i];
    const double *wp = wt + (size_t)i * K;
    double mx = -CG_HUGE;
    int jj = -1;
    if (cass_ptr) {
        const int e = mem_item[q];
        jj = cass_idx[cass_ptr[e]];
        for (int p = cass_ptr[e]; p < cass_ptr[e + 1]; p++) {
            int j = cass_idx[p];
            double y = __dmul_rn(wp[j], __dsub_rn(dat[j], x0));
            if (y > mx) { mx = y; jj = j; }
        }
    } else {  // singleton item: the complement is every other tree, ascending
        jj = i == 0 ? 1 : 0;
        for (int j = 0; j < K; j++) {
            if (j == i) continue;
            double y = __dmul_rn(wp[j], __dsub_rn(dat[j], x0));
            if (y > mx) { mx = y; jj = j; }
        }
    }
    obs[q] = mx;
    argj[q] = jj;  // the TREE that attains the maximum (first one in complement order)
}

int launch_col_top2(const double *x, int K, int B, int ld, double *m1, double *m2, int *a1, cudaStream_t st) {
    if (B <= 0) return CGPU_OK;
    const int tiles = (B + 127) / 128;
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    int ks = std::max(1, std::min({(8 * sms + tiles - 1) / tiles, (K + 15) / 16, 64}));
    const int rows_per = (K + ks - 1) / ks;
    ks = (K + rows_per - 1) / rows_per;
    if (ks == 1) {
        KT kt("col_top2_kernel", st, 8.0 * K * B);
        col_top2_kernel<<<dim3(tiles, 1), 128, 0, st>>>(x, K, B, ld, rows_per, m1, m2, a1);
        CGPU_LAUNCH_CHECK();
        return CGPU_OK;
    }
    double *pm = nullptr;
    int *pa = nullptr;
    CGPU_CUDA(dev_alloc(&pm, (size_t)2 * ks * B * sizeof(double)));
    CGPU_CUDA(dev_alloc(&pa, (size_t)ks * B * sizeof(int)));
    {
        KT kt("col_top2_kernel", st, 8.0 * K * B);
        col_top2_kernel<<<dim3(tiles, ks), 128, 0, st>>>(x, K, B, ld, rows_per, pm, pm + (size_t)ks * B, pa);
        CGPU_LAUNCH_CHECK();
    }
    {
        KT kt("col_top2_merge_kernel", st);
        col_top2_merge_kernel<<<tiles, 128, 0, st>>>(pm, pm + (size_t)ks * B, pa, ks, B, m1, m2, a1);
        CGPU_LAUNCH_CHECK();
    }
    dev_free(pm);
    dev_free(pa);
    return CGPU_OK;
}
int launch_stat_build(const double *x, int K, int B, int ld, const double *m1, const double *m2, const int *a1,
                      int cm, const int *ass_ptr, const int *ass_idx, double *stat, cudaStream_t st) {
    if (B <= 0 || cm <= 0) return CGPU_OK;
    int gy = std::min(cm, 64);
    KT kt("stat_build_kernel", st, 16.0 * cm * B);  // one replicate row read + one statistic row written per item (singleton items)
    stat_build_kernel<<<dim3((B + 127) / 128, gy), 128, 0, st>>>(x, K, B, ld, m1, m2, a1, cm, ass_ptr, ass_idx, stat);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}
int launch_center(const double *x, int K, int B, int ld, double *xc, cudaStream_t st, bool rescale, double rs) {
    KT kt("center_rows_kernel", st, 16.0 * K * B);
    center_rows_kernel<<<(K + 3) / 4, 128, 0, st>>>(x, K, B, ld, xc, rescale, rs);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}
int launch_pair_weight(const double *xc, int K, int B, int ld, double varadd, double *wt, int *nsmall,
                       cudaStream_t st, const int *d_rows, int nrows) {
    int t = (K + PV_T - 1) / PV_T;
    const int ty = d_rows ? (nrows + PV_T - 1) / PV_T : t;
    if (ty <= 0) return CGPU_OK;
    KT kt("pair_weight_kernel", st, 3.0 * B * (d_rows ? (double)nrows * (K - 1) : 0.5 * K * (K - 1)));  // (sub, mul, add) per pair and replicate
    pair_weight_kernel<<<dim3(t, ty), 256, 0, st>>>(xc, K, B, ld, varadd, wt, nsmall, d_rows, nrows);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}
int launch_wsh_single(const double *x, int K, int B, int ld, const double *wt, const int *mem_tree, int nmem, double *out,
                      cudaStream_t st) {
    if (nmem <= 0 || B <= 0) return CGPU_OK;
    KT kt("wsh_single_kernel", st, 3.0 * B * (double)nmem * (K - 1));  // (sub, mul, compare) per (member, other tree, replicate)
    wsh_single_kernel<<<dim3((B + 127) / 128, (nmem + WS_MI - 1) / WS_MI), 128, 0, st>>>(x, K, B, ld, wt, mem_tree, nmem, out);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}
int launch_sh_rows(bool weighted, const double *x, int K, int B, int ld, const double *wt, int nmem,
                   const int *mem_item, const int *mem_tree, const int *cass_ptr, const int *cass_idx, double *out,
                   cudaStream_t st) {
    if (nmem <= 0 || B <= 0) return CGPU_OK;
    dim3 grid((B + 127) / 128, nmem);
    KT kt(weighted ? "sh_rows_kernel<weighted>" : "sh_rows_kernel", st);
    if (weighted) sh_rows_kernel<true><<<grid, 128, 0, st>>>(x, K, B, ld, wt, mem_item, mem_tree, cass_ptr, cass_idx, out);
    else sh_rows_kernel<false><<<grid, 128, 0, st>>>(x, K, B, ld, wt, mem_item, mem_tree, cass_ptr, cass_idx, out);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}
int launch_weighted_obs(const double *dat, int K, const double *wt, int nmem, const int *mem_item,
                        const int *mem_tree, const int *cass_ptr, const int *cass_idx, double *obs, int *argj,
                        cudaStream_t st) {
    if (nmem <= 0) return CGPU_OK;
    KT kt("weighted_obs_kernel", st, 0.0);
    weighted_obs_kernel<<<(nmem + 127) / 128, 128, 0, st>>>(dat, K, wt, nmem, mem_item, mem_tree, cass_ptr, cass_idx, obs, argj);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

}  // namespace cgpu
