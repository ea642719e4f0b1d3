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
//   * ALGO 0 (default) is dense and branch free: the count byte is turned into fp64
//     through a 256-entry shared table and every (tree, replicate, site) term is formed,
//     including 0*L, so non-finite inputs behave exactly like the reference.
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
constexpr int CT_TILE_DBL = 2048;  // doubles of LT per pipeline stage (16 KB)
constexpr int CT_STAGES = 4;
constexpr int CT_CONSUMERS = 8;    // consumer warps = replicate groups per CTA
constexpr int CT_THREADS = 32 * (CT_CONSUMERS + 1);

template <int TK>
struct CtSmem {
    static constexpr int KT = 32 * TK;
    static constexpr int JC = CT_TILE_DBL / KT;
    static constexpr size_t l_bytes = (size_t)CT_STAGES * CT_TILE_DBL * sizeof(double);
    static constexpr size_t c_bytes = (size_t)CT_STAGES * CT_CONSUMERS * JC * sizeof(unsigned long long);
    static constexpr size_t total = l_bytes + c_bytes + 2 * CT_STAGES * sizeof(uint64_t);
};

// ALGO 0: dense, branch free (count -> fp64 through a 256-entry shared table)
// ALGO 1: sparse, warp-uniform branches on the count byte (requires finite L)
template <int TK, int MODE, int ALGO>
__global__ void __launch_bounds__(CT_THREADS, (TK >= 8 ? 1 : 2))
contract_kernel(const double *__restrict__ LT, const unsigned long long *__restrict__ C,
                double *__restrict__ rep, int K, int NP, int nb, int ldrep, int col0, double r) {
    using SM = CtSmem<TK>;
    constexpr int KT = SM::KT, JC = SM::JC;
    extern __shared__ __align__(128) unsigned char smraw[];
    double *Ls = reinterpret_cast<double *>(smraw);
    unsigned long long *Cs = reinterpret_cast<unsigned long long *>(smraw + SM::l_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smraw + SM::l_bytes + SM::c_bytes);
    uint64_t *empty = full + CT_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ngroups = (nb + 7) / 8;
    const int g0 = blockIdx.x * CT_CONSUMERS;
    const int ktile = blockIdx.y;
    const int nchunks = NP / JC;

    __shared__ double dtab[256];
    if (tid < 256) dtab[tid] = (double)tid;
    if (tid == 0) {
        for (int s = 0; s < CT_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CT_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (wid == CT_CONSUMERS) {
        // ---------------- producer warp: one lane drives the bulk copies
        if (lane == 0) {
            const int nact = min(CT_CONSUMERS, ngroups - g0);
            const uint32_t bytes = (uint32_t)(CT_TILE_DBL * sizeof(double) + (size_t)nact * JC * 8);
            const double *lsrc = LT + (size_t)ktile * NP * KT;
            for (int c = 0; c < nchunks; c++) {
                const int s = c % CT_STAGES;
                mbar_wait(&empty[s], ((c / CT_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(Ls + (size_t)s * CT_TILE_DBL, lsrc + (size_t)c * JC * KT, CT_TILE_DBL * sizeof(double), &full[s]);
                for (int w = 0; w < nact; w++)
                    bulk_g2s(Cs + ((size_t)s * CT_CONSUMERS + w) * JC, C + (size_t)(g0 + w) * NP + (size_t)c * JC, JC * 8, &full[s]);
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
        const int s = c % CT_STAGES;
        mbar_wait(&full[s], (c / CT_STAGES) & 1);
        if (active) {
            const double *lrow = Ls + (size_t)s * CT_TILE_DBL + lane;
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
                    if (ALGO == 0) {
                        const double dc = dtab[cnt];
#pragma unroll
                        for (int i = 0; i < TK; i++)
                            acc[t][i] = MODE == CGPU_SUM_STRICT ? __dadd_rn(acc[t][i], __dmul_rn(dc, lv[i]))
                                                                : __fma_rn(dc, lv[i], acc[t][i]);
                    } else {
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

// CGPU_CONTRACT_ALGO=1 selects the sparse variant (development A/B switch); it needs
// finite inputs because it never forms 0*L.
static int contract_algo(const DevMatrix &m) {
    static int env = [] {
        const char *e = getenv("CGPU_CONTRACT_ALGO");
        return e ? atoi(e) : 0;
    }();
    return (env == 1 && m.finite) ? 1 : 0;
}

template <int TK>
static int launch_contract_tk(const DevMatrix &m, const unsigned long long *d_counts, int nb, double r,
                              double *d_rep, int ldrep, int col0, int sum_mode, cudaStream_t st) {
    const int algo = contract_algo(m);
    if (sum_mode == CGPU_SUM_STRICT) {
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
