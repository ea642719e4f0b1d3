// Replicate sums rep[k][b] = ( sum_j c[b,j] * L[k][j] ) / r — the inner product that is
// >= 82% of makermt (src/makermt.c:65-70, marked "time consuming" at :68).
//
// contract_kernel<TK, MODE, ALGO>: order-preserving FP64 contraction.
//   * CTA tile = KT = 32*TK trees x 64 replicates; warp w owns replicate group
//     g = 8*blockIdx.x + w (8 replicates, one BJ8 count word per site), lane l owns
//     trees k = ktile*KT + l + 32*i.  Every thread keeps 8*TK fp64 accumulators in
//     registers and walks the sites j = 0..n-1 IN ORDER, so each output sees exactly
//     the reference's sequence of roundings:  x = rn(x + rn(c*L))  (MODE 0), then the
//     final true division by r.  MODE 1 (CGPU_SUM_FAST) fuses the two roundings (DFMA).
//   * ALGO 0 (default) is dense and branch free: a converter warp expands the count
//     bytes of each stage to fp64 in shared memory once, so the consumers' inner loop is
//     LDS + FP64 only, and every (tree, replicate, site) term is formed, including 0*L,
//     so non-finite inputs behave exactly like the reference.
//     ALGO 1 is the sparse variant: the count byte of (replicate, site) is warp-uniform,
//     so "c == 0 -> nothing to add" (x + (+-0) == x) and "c == 1 -> no multiply"
//     (1*L == L) are divergence-free branches that skip ~exp(-r) of the terms; it is
//     kept behind CGPU_CONTRACT_ALGO=1 until its branch overhead is tuned.
//   * operands are staged through shared memory by a dedicated producer warp with
//     cp.async.bulk (TMA bulk copies) completing on mbarriers: per stage one contiguous
//     tile of the site-major matrix LT[ktile][j0..j0+JC)[KT] and the eight count rows
//     of the CTA's replicate groups; consumers release a stage through an "empty"
//     mbarrier.  Site-major LT makes the per-site read of 32*TK trees conflict free.
//
// Bound: FP64 pipe (and the issue slots of the per-(replicate,site) branches), not HBM:
// algorithmic HBM bytes per launch are 8*K*n (L once) + n*nb (counts) + 8*K*nb (rep),
// see DESIGN.md.
#include <stdlib.h>

#include "cgpu_internal.h"

namespace cgpu {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ L -> site-major tiles
// LT[kt][j][kk] = L[kt*KT+kk][j], zero padded to NP sites and KT trees per tile.
__global__ void __launch_bounds__(256)
transpose_L_kernel(const double *__restrict__ L, int K, int n, int NP, int KT, double *__restrict__ LT,
                   int *__restrict__ nonfinite) {
    __shared__ double tile[32][33];
    const int j0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    int bad = 0;
    for (int r = ty; r < 32; r += 8) {
        int k = k0 + r, j = j0 + tx;
        double v = (k < K && j < n) ? L[(size_t)k * n + j] : 0.0;
        if (!isfinite(v)) bad = 1;
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int j = j0 + r, k = k0 + tx;  // write: consecutive threads -> consecutive trees
        if (j < NP) {
            int kt = k / KT, kk = k % KT;
            LT[((size_t)kt * NP + j) * KT + kk] = tile[tx][r];
        }
    }
    if (bad) atomicExch(nonfinite, 1);
}

int launch_transpose_L(const double *d_L, int K, int n, int NP, int TK, double *d_LT, int *d_nonfinite,
                       cudaStream_t st) {
    const int KT = 32 * TK;
    const int ktiles = (K + KT - 1) / KT;
    dim3 grid(NP / 32, ktiles * KT / 32);
    transpose_L_kernel<<<grid, 256, 0, st>>>(d_L, K, n, NP, KT, d_LT, d_nonfinite);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// ------------------------------------------------------------------ contraction
constexpr int CT_CONSUMERS = 8;    // consumer warps = replicate groups per CTA
constexpr int CT_THREADS = 32 * (CT_CONSUMERS + 2);  // + producer warp + converter warp

template <int TK>
struct CtSmem {
    static constexpr int KT = 32 * TK;
    static constexpr int JC = TK >= 4 ? 2048 / KT : 32;  // sites per pipeline stage
    static constexpr int TILE_DBL = JC * KT;              // doubles of LT per stage (<= 16 KB)
    static constexpr int STAGES = TK == 2 ? 3 : 4;
    static constexpr size_t l_bytes = (size_t)STAGES * TILE_DBL * sizeof(double);
    static constexpr size_t c_bytes = (size_t)STAGES * CT_CONSUMERS * JC * sizeof(unsigned long long);
    static constexpr size_t d_bytes = (size_t)STAGES * CT_CONSUMERS * JC * 8 * sizeof(double);
    static constexpr size_t m_bytes = (size_t)STAGES * CT_CONSUMERS * JC * sizeof(uint32_t);
    static constexpr size_t total = l_bytes + c_bytes + d_bytes + m_bytes + 3 * STAGES * sizeof(uint64_t);
};

// ALGO 0: dense, branch free.  A converter warp expands the count bytes of each stage to
//         fp64 once (shared memory), so the consumers' inner loop is loads + FP64 only.
// ALGO 1: sparse, warp-uniform branches on the count byte (requires finite L)
// ALGO 2: strict mode with half the FP64 instructions (requires finite L): x + rn(c*L) equals
//         fma(c, L, x) whenever c*L is exact, i.e. c in {0,1,2,4,8,...}.  The converter warp
//         zeroes the multiplier of every other count (3,5,6,7,...: ~6% of the terms at r=1)
//         and flags it in a per-site mask; the consumers run the dense DFMA loop and then
//         add rn(c*L) for the flagged (replicate, site) pairs with a separate DMUL + DADD.
//         Every accumulator still receives exactly one update per site, in site order.
template <int TK, int MODE, int ALGO>
__global__ void __launch_bounds__(CT_THREADS, (TK >= 8 ? 1 : 2))
contract_kernel(const double *__restrict__ LT, const unsigned long long *__restrict__ C,
                double *__restrict__ rep, int K, int NP, int nb, int ldrep, int col0, double r) {
    using SM = CtSmem<TK>;
    constexpr int KT = SM::KT, JC = SM::JC, NS = SM::STAGES, TILE = SM::TILE_DBL;
    extern __shared__ __align__(128) unsigned char smraw[];
    double *Ls = reinterpret_cast<double *>(smraw);
    unsigned long long *Cs = reinterpret_cast<unsigned long long *>(smraw + SM::l_bytes);
    double *Cd = reinterpret_cast<double *>(smraw + SM::l_bytes + SM::c_bytes);
    uint32_t *Ms = reinterpret_cast<uint32_t *>(smraw + SM::l_bytes + SM::c_bytes + SM::d_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smraw + SM::l_bytes + SM::c_bytes + SM::d_bytes + SM::m_bytes);
    uint64_t *ready = full + NS;   // counts of the stage are available as fp64
    uint64_t *empty = ready + NS;  // every consumer is done with the stage

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ngroups = (nb + 7) / 8;
    const int g0 = blockIdx.x * CT_CONSUMERS;
    const int ktile = blockIdx.y;
    const int nchunks = NP / JC;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], 1);
            mbar_init(&empty[s], CT_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (wid == CT_CONSUMERS) {
        // ---------------- producer warp: one lane drives the bulk copies
        if (lane == 0) {
            const int nact = min(CT_CONSUMERS, ngroups - g0);
            const uint32_t bytes = (uint32_t)(TILE * sizeof(double) + (size_t)nact * JC * 8);
            const double *lsrc = LT + (size_t)ktile * NP * KT;
            for (int c = 0; c < nchunks; c++) {
                const int s = c % NS;
                mbar_wait(&empty[s], ((c / NS) & 1) ^ 1);
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(Ls + (size_t)s * TILE, lsrc + (size_t)c * TILE, TILE * sizeof(double), &full[s]);
                for (int w = 0; w < nact; w++)
                    bulk_g2s(Cs + ((size_t)s * CT_CONSUMERS + w) * JC, C + (size_t)(g0 + w) * NP + (size_t)c * JC, JC * 8, &full[s]);
            }
        }
        return;
    }
    if (wid == CT_CONSUMERS + 1) {
        // ---------------- converter warp: count bytes -> fp64, once per stage
        if (ALGO != 1) {
            const int nact = min(CT_CONSUMERS, ngroups - g0);
            for (int c = 0; c < nchunks; c++) {
                const int s = c % NS;
                mbar_wait(&full[s], (c / NS) & 1);
                // the stage's fp64 area was last read in round c-NS; `empty` of that round
                // completed before the producer refilled `full`, so it is free to overwrite
                for (int e = lane; e < nact * JC; e += 32) {
                    const unsigned long long c64 = Cs[(size_t)s * CT_CONSUMERS * JC + e];
                    double v[8];
                    uint32_t mask = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const uint32_t cnt = (uint32_t)(c64 >> (8 * t)) & 0xffu;
                        const bool odd_one = ALGO == 2 && (cnt & (cnt - 1)) != 0;  // not 0 and not a power of two
                        v[t] = odd_one ? 0.0 : (double)cnt;
                        mask |= odd_one ? (1u << t) : 0u;
                    }
                    double4 *dst = reinterpret_cast<double4 *>(Cd + ((size_t)s * CT_CONSUMERS * JC + e) * 8);
                    dst[0] = make_double4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_double4(v[4], v[5], v[6], v[7]);
                    if (ALGO == 2) Ms[(size_t)s * CT_CONSUMERS * JC + e] = mask;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[s]);
            }
        }
        return;
    }

    // ---------------- consumer warps
    const int g = g0 + wid;
    const bool active = g < ngroups;
    double acc[8][TK];
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
        for (int i = 0; i < TK; i++) acc[t][i] = 0.0;

    for (int c = 0; c < nchunks; c++) {
        const int s = c % NS;
        mbar_wait(&full[s], (c / NS) & 1);
        if (ALGO != 1) mbar_wait(&ready[s], (c / NS) & 1);
        if (active) {
            const double *lrow = Ls + (size_t)s * TILE + lane;
            if (ALGO != 1) {
                const double2 *drow = reinterpret_cast<const double2 *>(Cd + ((size_t)s * CT_CONSUMERS + wid) * JC * 8);
                const uint32_t *mrow = Ms + ((size_t)s * CT_CONSUMERS + wid) * JC;
                const unsigned long long *crow = Cs + ((size_t)s * CT_CONSUMERS + wid) * JC;
                constexpr bool FUSE = MODE == CGPU_SUM_FAST || ALGO == 2;
#pragma unroll 2
                for (int jj = 0; jj < JC; jj++) {
                    double lv[TK];
#pragma unroll
                    for (int i = 0; i < TK; i++) lv[i] = lrow[jj * KT + 32 * i];
#pragma unroll
                    for (int t2 = 0; t2 < 4; t2++) {
                        const double2 dc = drow[jj * 4 + t2];
#pragma unroll
                        for (int i = 0; i < TK; i++) {
                            acc[2 * t2][i] = FUSE ? __fma_rn(dc.x, lv[i], acc[2 * t2][i])
                                                  : __dadd_rn(acc[2 * t2][i], __dmul_rn(dc.x, lv[i]));
                            acc[2 * t2 + 1][i] = FUSE ? __fma_rn(dc.y, lv[i], acc[2 * t2 + 1][i])
                                                      : __dadd_rn(acc[2 * t2 + 1][i], __dmul_rn(dc.y, lv[i]));
                        }
                    }
                    if (ALGO == 2 && MODE == CGPU_SUM_STRICT) {
                        const uint32_t mask = mrow[jj];
                        if (mask) {  // warp-uniform: some replicate of this group drew the site 3,5,6,7,... times
                            const unsigned long long c64 = crow[jj];
#pragma unroll
                            for (int t = 0; t < 8; t++) {
                                if (mask & (1u << t)) {
                                    const double dcn = (double)((uint32_t)(c64 >> (8 * t)) & 0xffu);
#pragma unroll
                                    for (int i = 0; i < TK; i++) acc[t][i] = __dadd_rn(acc[t][i], __dmul_rn(dcn, lv[i]));
                                }
                            }
                        }
                    }
                }
            } else {
                const unsigned long long *crow = Cs + ((size_t)s * CT_CONSUMERS + wid) * JC;
#pragma unroll 2
                for (int jj = 0; jj < JC; jj++) {
                    const unsigned long long c64 = crow[jj];
                    double lv[TK];
#pragma unroll
                    for (int i = 0; i < TK; i++) lv[i] = lrow[jj * KT + 32 * i];
                    const uint32_t clo = (uint32_t)c64, chi = (uint32_t)(c64 >> 32);
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const uint32_t cnt = ((t < 4 ? clo : chi) >> (8 * (t & 3))) & 0xffu;
                        if (cnt == 0) continue;
                        if (cnt == 1) {
#pragma unroll
                            for (int i = 0; i < TK; i++) acc[t][i] = __dadd_rn(acc[t][i], lv[i]);
                        } else {
                            const double dc = __uint2double_rn(cnt);
#pragma unroll
                            for (int i = 0; i < TK; i++)
                                acc[t][i] = MODE == CGPU_SUM_STRICT ? __dadd_rn(acc[t][i], __dmul_rn(dc, lv[i]))
                                                                    : __fma_rn(dc, lv[i], acc[t][i]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    if (active) {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int b = g * 8 + t;
            if (b >= nb) continue;
#pragma unroll
            for (int i = 0; i < TK; i++) {
                const int k = ktile * KT + lane + 32 * i;
                if (k < K) rep[(size_t)k * ldrep + col0 + b] = __ddiv_rn(acc[t][i], r);
            }
        }
    }
}

template <int TK, int MODE, int ALGO>
static int launch_contract_t(const DevMatrix &m, const unsigned long long *d_counts, int nb, double r,
                             double *d_rep, int ldrep, int col0, cudaStream_t st) {
    auto kern = contract_kernel<TK, MODE, ALGO>;
    static bool attr_done = false;
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CtSmem<TK>::total));
        attr_done = true;
    }
    const int ngroups = (nb + 7) / 8;
    dim3 grid((ngroups + CT_CONSUMERS - 1) / CT_CONSUMERS, m.ktiles);
    kern<<<grid, CT_THREADS, CtSmem<TK>::total, st>>>(m.d_LT, d_counts, d_rep, m.K, m.NP, nb, ldrep, col0, r);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// Default: strict mode on finite data uses ALGO 2 (DFMA + flagged fix-ups), anything else
// the plain dense ALGO 0.  CGPU_CONTRACT_ALGO=0/1/2 forces a variant (development A/B
// switch); the sparse and fix-up variants need finite inputs because they never form 0*L.
static int contract_algo(const DevMatrix &m, int sum_mode) {
    static int env = [] {
        const char *e = getenv("CGPU_CONTRACT_ALGO");
        return e ? atoi(e) : -1;
    }();
    const int forced = G().opt_contract_algo >= 0 ? G().opt_contract_algo : env;
    int algo = forced >= 0 ? forced : (sum_mode == CGPU_SUM_STRICT ? 2 : 0);
    if (!m.finite) algo = 0;
    if (algo == 2 && sum_mode != CGPU_SUM_STRICT) algo = 0;
    return algo;
}

template <int TK>
static int launch_contract_tk(const DevMatrix &m, const unsigned long long *d_counts, int nb, double r,
                              double *d_rep, int ldrep, int col0, int sum_mode, cudaStream_t st) {
    const int algo = contract_algo(m, sum_mode);
    if (sum_mode == CGPU_SUM_STRICT) {
        if (algo == 2) return launch_contract_t<TK, CGPU_SUM_STRICT, 2>(m, d_counts, nb, r, d_rep, ldrep, col0, st);
        if (algo == 1) return launch_contract_t<TK, CGPU_SUM_STRICT, 1>(m, d_counts, nb, r, d_rep, ldrep, col0, st);
        return launch_contract_t<TK, CGPU_SUM_STRICT, 0>(m, d_counts, nb, r, d_rep, ldrep, col0, st);
    }
    if (algo == 1) return launch_contract_t<TK, CGPU_SUM_FAST, 1>(m, d_counts, nb, r, d_rep, ldrep, col0, st);
    return launch_contract_t<TK, CGPU_SUM_FAST, 0>(m, d_counts, nb, r, d_rep, ldrep, col0, st);
}

int launch_contract(const DevMatrix &m, const unsigned long long *d_counts, int nb, double r, double *d_rep,
                    int ldrep, int col0, int sum_mode, cudaStream_t st) {
    if (nb <= 0) return CGPU_OK;
    switch (m.TK) {
        case 1: return launch_contract_tk<1>(m, d_counts, nb, r, d_rep, ldrep, col0, sum_mode, st);
        case 2: return launch_contract_tk<2>(m, d_counts, nb, r, d_rep, ldrep, col0, sum_mode, st);
        case 4: return launch_contract_tk<4>(m, d_counts, nb, r, d_rep, ldrep, col0, sum_mode, st);
        case 8: return launch_contract_tk<8>(m, d_counts, nb, r, d_rep, ldrep, col0, sum_mode, st);
    }
    set_error(CGPU_EINVAL, "bad tree tile %d", m.TK);
    return CGPU_EINVAL;
}

}  // namespace cgpu
