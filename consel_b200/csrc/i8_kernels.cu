// CGPU_SUM_INT8: the replicate sums as an EXACT fixed-point contraction on the 5th-gen
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
//   L[k][j] = M[k][j] * 2^-F with M an integer (F chosen so that EVERY entry is exact),
//   M = sum_s d_s * 256^s with signed digits d_s in [-128,127]  (S slices, "Ozaki splitting")
//   T_s[k][b] = sum_j c[b][j] * d_s[k][j]          int8 x uint8 -> int32, one GEMM with the
//                                                  slices stacked along M (rows k*S+s)
//   rep[k][b] = rn( rn( sum_s T_s * 256^s * 2^-F ) / r )
//
// The integer sum is exact, so the result is the correctly rounded sum (one rounding) and
// is identical for every tiling, batch split and GPU count; it differs from the
// reference's sequentially rounded sum by <= n*2^-53 relative (tests: bit-equal to a
// big-integer restatement, <= 1e-9 relative to the reference).
//
// Kernels
//   fx_scan_kernel     exponent range of L (lowest set mantissa bit, highest magnitude bit)
//   fx_slice_kernel    L -> digit matrix A[(k*S+s)][j] (int8, K-major rows of NP bytes)
//   gemm_i8_persistent_kernel
//                      D[128 x 256] tiles, one CTA per SM: TMA (128B swizzle) -> 4-stage smem ring
//                      -> tcgen05.mma kind::i8 -> two TMEM accumulator stages -> eight epilogue
//                      warps (tcgen05.ld).  The epilogue either stores the int32 digit sums or,
//                      for S in {1,2,4,8}, finishes the job in place: transpose through shared
//                      memory, exact fold of the digits, ONE rounding (an FMA of two exact fp64
//                      group sums), scale, divide by r, coalesced store of the replicate block.
//   gemm_i8_kernel     one tile per CTA (kept as the "i8_oneshot" measurement variant)
//   fx_combine_kernel  the recombination as a separate pass over the int32 sums (other S)
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <vector>

#include "cgpu_internal.h"

namespace cgpu {

// ------------------------------------------------------------------ fixed-point set-up
// per entry x = (-1)^s * m * 2^e with m a 53-bit integer: lo = e + ctz(m) (exponent of the
// lowest set bit), hi = e + 52 - clz-adjust (exponent of the highest set bit)
__global__ void __launch_bounds__(256)
fx_scan_kernel(const double *__restrict__ L, size_t count, int *__restrict__ out /* [0]=min lo, [1]=max hi, [2]=nonfinite */) {
    int lo = 1 << 30, hi = -(1 << 30), bad = 0;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(L[e]);
        const int ex = (int)((u >> 52) & 0x7ff);
        unsigned long long m = u & 0xfffffffffffffull;
        if (ex == 0x7ff) { bad = 1; continue; }
        int e2;
        if (ex == 0) {
            if (m == 0) continue;  // zero contributes nothing
            e2 = -1074;
        } else {
            m |= 1ull << 52;
            e2 = ex - 1075;
        }
        lo = min(lo, e2 + (int)__ffsll((long long)m) - 1);
        hi = max(hi, e2 + 63 - (int)__clzll((long long)m));
    }
    for (int off = 16; off; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], lo);
        atomicMax(&out[1], hi);
        if (bad) atomicExch(&out[2], 1);
    }
}

// A[(k*S+s)*NP + j] = digit s of M = L[k][col(j)]*2^F, col(j) = rep[j] (pattern slots) or j; ncols columns;
// rows >= K*S and columns >= ncols stay zero
__global__ void __launch_bounds__(256)
fx_slice_kernel(const double *__restrict__ L, int K, int n, int ncols, const unsigned int *__restrict__ rep, int NP, int S, int F,
                signed char *__restrict__ A) {
    const size_t total = (size_t)K * ncols;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / ncols), j = (int)(e % ncols);
        const unsigned long long u = (unsigned long long)__double_as_longlong(L[(size_t)k * n + (rep ? rep[j] : (unsigned int)j)]);
        const int ex = (int)((u >> 52) & 0x7ff);
        unsigned long long m = u & 0xfffffffffffffull;
        int e2 = -1074;
        if (ex) {
            m |= 1ull << 52;
            e2 = ex - 1075;
        }
        // M = +-m * 2^(e2+F); F >= -lo guarantees the shift loses no set bit
        const int sh = e2 + F;
        unsigned __int128 mag = m;
        if (sh >= 0) mag <<= sh; else mag >>= -sh;
        __int128 v = (u >> 63) ? -(__int128)mag : (__int128)mag;
        for (int s = 0; s < S; s++) {
            int d = (int)(v & 0xff);
            if (d >= 128) d -= 256;  // balanced digit; the borrow moves to the next slice
            A[((size_t)k * S + s) * NP + j] = (signed char)d;
            v = (v - d) >> 8;
        }
    }
}

// Correctly rounded (nearest-even) conversion of a 128-bit integer: values beyond 63 bits are
// first shortened to 63 bits with the shifted-out bits OR-ed into the last place (round to
// odd); 63 >= 53 + 2 bits, so the hardware's single rounding of that integer is the
// rounding of the full value.  The power of two is applied exactly afterwards.
__device__ __forceinline__ double i128_to_double_rn(__int128 v) {
    const bool neg = v < 0;
    const unsigned __int128 a = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
    const unsigned long long hi = (unsigned long long)(a >> 64), lo = (unsigned long long)a;
    double r;
    if (hi == 0 && (long long)lo >= 0) {
        r = (double)(long long)lo;
    } else {
        const int msb = hi ? 127 - __clzll((long long)hi) : 63;
        const int sh = msb - 62;  // 1..65
        unsigned long long q = (unsigned long long)(a >> sh);
        const unsigned __int128 lost = a & ((((unsigned __int128)1) << sh) - 1);
        q |= (lost != 0) ? 1ull : 0ull;
        r = (double)(long long)q * __longlong_as_double((long long)(1023 + sh) << 52);
    }
    return neg ? -r : r;
}

// 2^-F applied to an exactly representable value: one multiply when 2^-F is a normal double
struct FxScale {
    double mul;   // 2^-F, or 0 when the exponent is outside the normal range
    int F;
};
__device__ __forceinline__ double fx_scale(double x, const FxScale &sc) {
    return sc.mul != 0.0 ? x * sc.mul : scalbn(x, -sc.F);
}

// exact sum of the S digit sums d[s*stride]: digits are folded four at a time in 64 bits
// (|group| < 2^56), the groups in 128 bits
template <int S>
__device__ __forceinline__ __int128 fx_digits_to_int(const int *d, int stride) {
    __int128 t = 0;
#pragma unroll
    for (int g = (S + 3) / 4 - 1; g >= 0; g--) {
        long long x = 0;
#pragma unroll
        for (int sdig = (S < 4 * g + 4 ? S : 4 * g + 4) - 1; sdig >= 4 * g; sdig--) x = x * 256 + (long long)d[sdig * stride];
        t = (t << 32) + (__int128)x;
    }
    return t;
}

// P: int32 partial sums [Mpad][ldp] (row k*S+s, column = replicate of the batch)
__global__ void __launch_bounds__(256)
fx_combine_kernel(const int *__restrict__ P, int ldp, int K, int S, int F, int nb, double r, double *__restrict__ rep,
                  int ldrep, int col0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (b >= nb || k >= K) return;
    __int128 t = 0;
    for (int s = S - 1; s >= 0; s--) t = (t << 8) + (__int128)P[((size_t)k * S + s) * ldp + b];
    const double x = scalbn(i128_to_double_rn(t), -F);  // exact scaling of the rounded sum
    rep[(size_t)k * ldrep + col0 + b] = __ddiv_rn(x, r);
}

// ------------------------------------------------------------------ tcgen05 GEMM
// Parameters of the fused recombination epilogue (rep == nullptr: plain int32 partial sums).
// Requires 32 % S == 0 so that the digits of a tree never straddle an epilogue warp.
struct FxEpilogue {
    double *rep;
    double r;
    FxScale sc;
    int K, S, nb, ldrep, col0;
    int dbg;  // measurement only: 1 = no conversion/division, 2 = no arithmetic at all, 3 = no transpose either
};

// one 32 x 32 block of digit sums (rows = (tree, digit), transposed in `tile`, stride 33):
// lane l finishes replicate column b for each of the 32/S trees of the block
// S <= 8 on the fused path: the digits fold into two 64-bit groups g1 (digits 4..7) and g0
// (digits 0..3), |g| <= 128*n'*(256^4-1)/255 < 2^31.01 * n' < 2^53 (n' < 2^21 is checked on the host), so both convert
// to fp64 exactly and ONE fused multiply-add rounds the exact value g1*2^32 + g0 once.
template <int S>
__device__ __forceinline__ double fx_digits_to_double(const int *d, int stride) {
    long long g0 = 0, g1 = 0;
#pragma unroll
    for (int sdig = (S < 4 ? S : 4) - 1; sdig >= 0; sdig--) g0 = g0 * 256 + (long long)d[sdig * stride];
    if (S <= 4) return (double)g0;
#pragma unroll
    for (int sdig = S - 1; sdig >= 4; sdig--) g1 = g1 * 256 + (long long)d[sdig * stride];
    return __fma_rn((double)g1, 4294967296.0, (double)g0);
}

template <int S>
__device__ __forceinline__ void fx_finish_block(const int *tile, int lane, int k0, int b, const FxEpilogue &ep) {
    static_assert(S <= 8, "the fused recombination folds at most two 64-bit groups");
    constexpr int TREES = 32 / S;
    // the TREES results of this lane are independent: every stage is done for all of them
    // before the next one, so the dependent chains (fold, conversion, division) overlap
    // instead of running back to back with only two warps per scheduler to hide them
    double x[TREES];
#pragma unroll
    for (int j = 0; j < TREES; j++) x[j] = fx_digits_to_double<S>(tile + (j * S) * 33 + lane, 33) * ep.sc.mul;
    if (ep.dbg == 1) {
#pragma unroll
        for (int j = 0; j < TREES; j++)
            if (k0 + j < ep.K && b < ep.nb) ep.rep[(size_t)(k0 + j) * ep.ldrep + ep.col0 + b] = x[j];
        return;
    }
#pragma unroll
    for (int j = 0; j < TREES; j++) x[j] = __ddiv_rn(x[j], ep.r);
#pragma unroll
    for (int j = 0; j < TREES; j++) {
        const int k = k0 + j;
        if (k < ep.K && b < ep.nb) ep.rep[(size_t)k * ep.ldrep + ep.col0 + b] = x[j];
    }
}

// ------------------------------------------------------------------ tcgen05 GEMM tiles
constexpr int G8_BM = 128, G8_BN = 256, G8_BK = 128;  // BK in bytes == int8 elements (one 128B swizzle atom)
constexpr int G8_STAGES = 4;  // one-shot kernel; the persistent kernel is a template over (stages, epilogue warps)
constexpr int G8_THREADS = 192;  // one-shot kernel: warp0 TMA, warp1 MMA (+TMEM alloc), warps 2-5 epilogue
// persistent kernel configurations: (4 stages, 8 epilogue warps) keeps the tensor pipe fed on long
// tiles; (3 stages, 12 epilogue warps) for short tiles, where the fused epilogue of a tile takes as
// long as its MMAs and the freed 48 KB hold four more transpose buffers
template <int STAGES, int EW> struct G8P {
    static constexpr int threads = 64 + 32 * EW;
    static constexpr size_t epi_bytes = (size_t)EW * 32 * 33 * 4;  // per epilogue warp a 32 x 32 int32 transpose tile (stride 33)
    static constexpr size_t smem = STAGES * ((size_t)128 * 128 + (size_t)256 * 128) + 1024 /*align*/ + 256 /*barriers*/ + epi_bytes;
};
constexpr size_t G8_A_BYTES = (size_t)G8_BM * G8_BK, G8_B_BYTES = (size_t)G8_BN * G8_BK;
constexpr size_t G8_SMEM = G8_STAGES * (G8_A_BYTES + G8_B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(s2u(b)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            s2u(dst)),
        "l"(map), "r"(s2u(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile, rows of 128 bytes, 128B swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem) {
    uint64_t d = (uint64_t)((s2u(smem) & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}

// D (int32, [Mpad][ldp]) = A (int8 digits, [Mpad][NP]) x B^T (uint8 counts, [Npad][NP])
__global__ void __launch_bounds__(G8_THREADS, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int NP,
               int *__restrict__ P, int ldp, int ncols_valid) {
    extern __shared__ unsigned char g8_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(g8_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = base;
    unsigned char *sB = base + G8_STAGES * G8_A_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(base + G8_STAGES * (G8_A_BYTES + G8_B_BYTES));
    uint64_t *empty = full + G8_STAGES;
    uint64_t *acc_done = empty + G8_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * G8_BM, n0 = blockIdx.y * G8_BN;
    const int nkb = NP / G8_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G8_STAGES; s++) {
            mb_init(&full[s], 1);
            mb_init(&empty[s], 1);
        }
        mb_init(acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // one warp allocates the accumulator columns in tensor memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(G8_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % G8_STAGES;
                mb_wait(&empty[s], ((kb / G8_STAGES) & 1) ^ 1);
                mb_expect(&full[s], (uint32_t)(G8_A_BYTES + G8_B_BYTES));
                tma_load_2d(sA + s * G8_A_BYTES, &mapA, kb * G8_BK, m0, &full[s]);
                tma_load_2d(sB + s * G8_B_BYTES, &mapB, kb * G8_BK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=S32, A=signed 8 bit, B=unsigned 8 bit, both K-major, N=256, M=128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(G8_BN >> 3) << 17) | ((uint32_t)(G8_BM >> 4) << 24);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % G8_STAGES;
                mb_wait(&full[s], (kb / G8_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t ad = umma_desc_sw128(sA + s * G8_A_BYTES), bd = umma_desc_sw128(sB + s * G8_B_BYTES);
#pragma unroll
                for (int k = 0; k < G8_BK / 32; k++)  // UMMA_K = 32 int8 = 32 bytes = 2 descriptor units
                    umma_i8(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
            }
            umma_commit(acc_done);
        }
    } else {
        // epilogue: warp w may only touch TMEM lanes 32*(w%4) .. +31
        const int q = warp & 3;
        mb_wait(acc_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
        int *dst = P + (size_t)row * ldp + n0;
#pragma unroll 1
        for (int c = 0; c < G8_BN; c += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n0 + c < ncols_valid) {
                uint4 *d4 = reinterpret_cast<uint4 *>(dst + c);
#pragma unroll
                for (int i = 0; i < 8; i++) d4[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(G8_BN) : "memory");
}

// Persistent variant: one CTA per SM walks the tile list (m fastest, so CTAs running at the
// same time share count rows in L2); the shared-memory ring runs across tiles, and two
// accumulator stages in tensor memory (2 x 256 columns) let the epilogue of tile i drain
// while the MMAs of tile i+1 are issued.
template <int STAGES, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_i8_persistent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int NP,
                          int *__restrict__ P, int ldp, int m_tiles, int n_tiles, FxEpilogue ep) {
    extern __shared__ unsigned char g8_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(g8_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = base;
    unsigned char *sB = base + STAGES * G8_A_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(base + STAGES * (G8_A_BYTES + G8_B_BYTES));
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;  // [2]
    uint64_t *acc_empty = acc_full + 2;      // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    int *epi = reinterpret_cast<int *>(base + STAGES * (G8_A_BYTES + G8_B_BYTES) + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = NP / G8_BK;
    const int total_tiles = m_tiles * n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mb_init(&full[s], 1);
            mb_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mb_init(&acc_full[a], 1);
            mb_init(&acc_empty[a], EW);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(2 * G8_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;  // k-block counter across tiles
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile % m_tiles) * G8_BM, n0 = (tile / m_tiles) * G8_BN;
                // The count rows stream from HBM (each tile of them is used by only m_tiles CTAs):
                // at ~2 us of latency the four stages in flight cannot feed the tensor pipe, so
                // the rows of this CTA's NEXT tile are pulled into L2 one tile ahead.
                const int nt = tile + gridDim.x;
                // (short tiles only: a long tile's rows would evict the operands in use from L2)
                const bool pre = NP <= 8192 && nt < total_tiles && nt / m_tiles != tile / m_tiles;
                const int n0n = (nt / m_tiles) * G8_BN;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    mb_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    mb_expect(&full[s], (uint32_t)(G8_A_BYTES + G8_B_BYTES));
                    tma_load_2d(sA + s * G8_A_BYTES, &mapA, kb * G8_BK, m0, &full[s]);
                    tma_load_2d(sB + s * G8_B_BYTES, &mapB, kb * G8_BK, n0, &full[s]);
                    if (pre) tma_prefetch_2d(&mapB, kb * G8_BK, n0n);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(G8_BN >> 3) << 17) | ((uint32_t)(G8_BM >> 4) << 24);
            int it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tcount++) {
                const int acc = tcount & 1;
                mb_wait(&acc_empty[acc], ((tcount >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + (uint32_t)(acc * G8_BN);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES;
                    mb_wait(&full[s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t ad = umma_desc_sw128(sA + s * G8_A_BYTES), bd = umma_desc_sw128(sB + s * G8_B_BYTES);
#pragma unroll
                    for (int k = 0; k < G8_BK / 32; k++) umma_i8(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // epilogue warp: TMEM lane quarter q (= warp % 4, the hardware's access rule), column half h
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        // column blocks (32 wide) of this warp: the 8 blocks of a tile are dealt 3 / 3 / 2 to the
        // three warps of a lane quarter (or 4 / 4 with eight epilogue warps)
        constexpr int NSPLIT = EW / 4, NBLK = G8_BN / 32;
        const int blk_lo = h * NBLK / NSPLIT, blk_hi = (h + 1) * NBLK / NSPLIT;
        int *tile_s = epi + (warp - 2) * (32 * 33);
        int tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tcount++) {
            const int acc = tcount & 1;
            const int m0 = (tile % m_tiles) * G8_BM, n0 = (tile / m_tiles) * G8_BN;
            mb_wait(&acc_full[acc], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int *dst = P + (size_t)(m0 + q * 32 + lane) * ldp + n0;
#pragma unroll 1
            for (int c = blk_lo * 32; c < blk_hi * 32; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G8_BN + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 32 == blk_hi * 32) {  // last read of this accumulator by this warp: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(&acc_empty[acc])) : "memory");
                }
                if (ep.rep == nullptr) {
                    uint4 *d4 = reinterpret_cast<uint4 *>(dst + c);
#pragma unroll
                    for (int i = 0; i < 8; i++) d4[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    continue;
                }
                // Fused recombination (what fx_combine_kernel does from HBM): this lane holds
                // row (k*S+s) of 32 replicates; transpose through shared memory so that lane l
                // owns replicate column c+l and walks the S digits of each of the 32/S trees.
                __syncwarp();
                if (ep.dbg != 3) {
#pragma unroll
                    for (int i = 0; i < 32; i++) tile_s[lane * 33 + i] = (int)v[i];
                }
                __syncwarp();
                const int b = n0 + c + lane;  // replicate (column of the batch)
                const int k0 = (m0 + q * 32) / ep.S;
                switch (ep.S) {
                    case 1: fx_finish_block<1>(tile_s, lane, k0, b, ep); break;
                    case 2: fx_finish_block<2>(tile_s, lane, k0, b, ep); break;
                    case 4: fx_finish_block<4>(tile_s, lane, k0, b, ep); break;
                    default: fx_finish_block<8>(tile_s, lane, k0, b, ep); break;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * G8_BN) : "memory");
}

// ------------------------------------------------------------------ 2-CTA GEMM (cta_group::2)
// The one-CTA kernel above is bound by shared-memory bandwidth, not by the tensor pipe: a
// 128x256x32 instruction reads 12 KB of operands while TMA writes the next 12 KB, 184 B/clk
// against the 128 B/clk an SM's shared memory moves, which caps it near 70 % of the int8 issue
// rate (measured: 0.68).  Here two CTAs of a cluster (one TPC) issue ONE 256x256x32 instruction
// per step: each CTA stages only its own 128 replicate rows and HALF of the digit rows
// (16 KB + 16 KB per 128-byte k-block instead of 16 + 32), the hardware shares the digit halves
// between the two tensor cores.
//
// Roles are also swapped with respect to the kernel above: the REPLICATES are the M side (TMEM
// lanes) and the (tree, digit) rows the N side (TMEM columns), so an epilogue thread owns one
// replicate and reads the S digits of a tree from consecutive registers -- the exact fold, the
// single rounding and the division need no shared-memory transpose, and a warp stores 32
// consecutive replicates of a tree with one coalesced 256-byte store.
//
// Protocol (both CTAs run the same code; rank 0 of the cluster is the leader):
//   producers   each CTA's warp 0 waits on its own empty[s], then issues its two TMA loads with
//               cta_group::2 completing on the LEADER's full[s]; the leader alone arms full[s]
//               with the bytes of both CTAs
//   MMA         the leader's warp 1 issues tcgen05.mma.cta_group::2 and commits (multicast) to
//               empty[s] of both CTAs, and to acc_full[a] of both CTAs after a tile
//   epilogue    every epilogue warp of both CTAs drains its own TMEM lanes and arrives on the
//               LEADER's acc_empty[a] (count 2 x EW)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: rank 0's copy
constexpr size_t G2_OP_BYTES = (size_t)128 * G8_BK;  // one operand tile of a CTA: 128 rows x 128 bytes

template <int STAGES, int EW> struct G2P {
    static constexpr int threads = 64 + 32 * EW;
    static constexpr size_t smem = STAGES * 2 * G2_OP_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            s2u(dst)),
        "l"(map), "r"(s2u(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {  // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s2u(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mb_arrive_leader(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(s2u(bar) & PEER_BIT_MASK) : "memory");
}

// Tile order of the pair kernel.  The ~74 pairs that run at the same time sweep the site axis
// roughly in step, so every operand byte is fetched from HBM once per WAVE that uses it: a wave
// should be a near-square patch of the tile grid.  Panels of G2_PANEL count-tile columns are
// walked digit-tile by digit-tile (panel width 8: 74 consecutive tiles cover 8 x ~9 (17 operand tiles per
// wave instead of 75 with a plain row-major order: at C5 that is 0.36 instead of 1.6 TB per step).
__device__ __forceinline__ void g2_tile_coords(int tile, int c_tiles, int d_tiles, int pw, int &ct, int &dt) {
    if (pw <= 0) {  // plain order, digit tiles fastest ("i8_panel" 0: measurement)
        ct = tile / d_tiles;
        dt = tile - ct * d_tiles;
        return;
    }
    const int per_panel = pw * d_tiles;
    const int panel = tile / per_panel, rem = tile - panel * per_panel;
    const int width = min(pw, c_tiles - panel * pw);
    dt = rem / width;
    ct = panel * pw + (rem - dt * width);
}

// this thread's replicate b, 32/S trees starting at k0: digits in consecutive registers
template <int S>
__device__ __forceinline__ void fx_finish_regs(const uint32_t (&v)[32], int k0, int b, const FxEpilogue &ep) {
    static_assert(S <= 8, "the fused recombination folds at most two 64-bit groups");
    constexpr int TREES = 32 / S;
    double x[TREES];
#pragma unroll
    for (int j = 0; j < TREES; j++) {
        long long g0 = 0, g1 = 0;
#pragma unroll
        for (int sdig = (S < 4 ? S : 4) - 1; sdig >= 0; sdig--) g0 = g0 * 256 + (long long)(int)v[j * S + sdig];
        if (S <= 4) {
            x[j] = (double)g0 * ep.sc.mul;
        } else {
#pragma unroll
            for (int sdig = S - 1; sdig >= 4; sdig--) g1 = g1 * 256 + (long long)(int)v[j * S + sdig];
            x[j] = __fma_rn((double)g1, 4294967296.0, (double)g0) * ep.sc.mul;
        }
    }
    if (ep.dbg != 1) {
#pragma unroll
        for (int j = 0; j < TREES; j++) x[j] = __ddiv_rn(x[j], ep.r);
    }
    if (b < ep.nb) {
#pragma unroll
        for (int j = 0; j < TREES; j++)
            if (k0 + j < ep.K) ep.rep[(size_t)(k0 + j) * ep.ldrep + ep.col0 + b] = x[j];
    }
}

// mapC: count rows [nb_pad][NP] (uint8), box 128 rows; mapD: digit rows [Mpad][NP] (int8), box 128 rows.
// A pair's tile: 256 replicates (ct) x 256 digit rows (dt), visited in the order of g2_tile_coords.
// P (unfused path): [p_rows][ldp] int32.
template <int STAGES, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_i8_2cta_kernel(const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapD, int NP,
                    int *__restrict__ P, int ldp, int p_rows, int c_tiles, int d_tiles, int panel_w, int ksplit, FxEpilogue ep) {
    extern __shared__ unsigned char g8_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(g8_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sC = base;
    unsigned char *sD = base + STAGES * G2_OP_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(base + STAGES * 2 * G2_OP_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;  // [2]
    uint64_t *acc_empty = acc_full + 2;   // [2] (the leader's copy is the one in use)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nkb = NP / G8_BK;
    // work items: (tile, part of the site axis).  ksplit > 1 (few tiles, very long rows: C4) lets every
    // pair work; the int32 partial sums of the parts are added with red.global (exact, order free)
    const int total_tiles = c_tiles * d_tiles * ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mb_init(&full[s], 1);
            mb_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mb_init(&acc_full[a], 1);
            mb_init(&acc_empty[a], 2 * EW);  // one arrival per epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs allocates the pair's accumulator columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(2 * G8_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // barriers of both CTAs are initialised before anybody signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;  // k-block counter across tiles
            for (int tile = pair; tile < total_tiles; tile += npairs) {
                int ct, dt;
                g2_tile_coords(tile / ksplit, c_tiles, d_tiles, panel_w, ct, dt);
                const int kp = tile % ksplit;
                const int kb0 = (int)((long long)nkb * kp / ksplit), kb1 = (int)((long long)nkb * (kp + 1) / ksplit);
                const int row_c = ct * 256 + (int)rank * 128;
                const int row_d = dt * 256 + (int)rank * 128;
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const int s = it % STAGES;
                    mb_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    if (leader) mb_expect(&full[s], (uint32_t)(4 * G2_OP_BYTES));  // both operands of both CTAs
                    tma_load_2d_2sm(sC + s * G2_OP_BYTES, &mapC, kb * G8_BK, row_c, &full[s]);
                    tma_load_2d_2sm(sD + s * G2_OP_BYTES, &mapD, kb * G8_BK, row_d, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // instruction descriptor: D=S32, A=unsigned 8 bit (counts), B=signed 8 bit (digits), both K-major, N=256, M=256
            const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            int it = 0, tcount = 0;
            for (int tile = pair; tile < total_tiles; tile += npairs, tcount++) {
                const int acc = tcount & 1;
                mb_wait(&acc_empty[acc], ((tcount >> 1) & 1) ^ 1);  // both epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + (uint32_t)(acc * G8_BN);
                const int kp = tile % ksplit;
                const int kb0 = (int)((long long)nkb * kp / ksplit), kb1 = (int)((long long)nkb * (kp + 1) / ksplit);
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const int s = it % STAGES;
                    mb_wait(&full[s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t ad = umma_desc_sw128(sC + s * G2_OP_BYTES), bd = umma_desc_sw128(sD + s * G2_OP_BYTES);
#pragma unroll
                    for (int k = 0; k < G8_BK / 32; k++) umma_i8_2sm(d_tmem, ad + 2 * k, bd + 2 * k, idesc, ((kb - kb0) | k) ? 1u : 0u);
                    umma_commit_2sm(&empty[s]);
                }
                umma_commit_2sm(&acc_full[acc]);
            }
        }
    } else {
        // epilogue warp: TMEM lane quarter q (= warp % 4), column range h of NSPLIT
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        constexpr int NSPLIT = EW / 4, NBLK = G8_BN / 32;
        const int blk_lo = h * NBLK / NSPLIT, blk_hi = (h + 1) * NBLK / NSPLIT;
        int tcount = 0;
        for (int tile = pair; tile < total_tiles; tile += npairs, tcount++) {
            const int acc = tcount & 1;
            int ct, dt;
            g2_tile_coords(tile / ksplit, c_tiles, d_tiles, panel_w, ct, dt);
            const int b = ct * 256 + (int)rank * 128 + q * 32 + lane;  // this thread's replicate (column of the batch)
            const int n_base = dt * 256;                               // first digit row of the tile
            mb_wait(&acc_full[acc], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = blk_lo * 32; c < blk_hi * 32; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G8_BN + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 32 == blk_hi * 32) {  // last read of this accumulator by this warp: hand it back to the leader
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mb_arrive_leader(&acc_empty[acc]);
                }
                const int n0 = n_base + c;
                if (ep.rep == nullptr) {
                    if (b < ldp) {
                        if (ksplit > 1) {
#pragma unroll
                            for (int i = 0; i < 32; i++)
                                if (n0 + i < p_rows) atomicAdd(&P[(size_t)(n0 + i) * ldp + b], (int)v[i]);  // RED.ADD: no return value used
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; i++)
                                if (n0 + i < p_rows) P[(size_t)(n0 + i) * ldp + b] = (int)v[i];
                        }
                    }
                    continue;
                }
                if (ep.dbg >= 2) continue;
                const int k0 = n0 / ep.S;
                switch (ep.S) {
                    case 1: fx_finish_regs<1>(v, k0, b, ep); break;
                    case 2: fx_finish_regs<2>(v, k0, b, ep); break;
                    case 4: fx_finish_regs<4>(v, k0, b, ep); break;
                    default: fx_finish_regs<8>(v, k0, b, ep); break;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // the peer may still be reading operands / signalling barriers of this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * G8_BN) : "memory");
}

// ------------------------------------------------------------------ measured int8 tensor peak
// The bare issue rate of tcgen05.mma kind::i8 (M=128, N=256, K=32 per instruction) from operands
// that already sit in shared memory: no TMA, no epilogue, two accumulator stages alternating like
// the GEMM above.  One CTA per SM; `iters` k-blocks of four instructions each.  This is the
// roofline denominator bench.py grades gemm_i8_persistent_kernel against (MEASURED_PEAKS.json
// carries bf16 only).
__global__ void __launch_bounds__(64, 1) i8_peak_kernel(int iters) {
    extern __shared__ unsigned char g8_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(g8_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = base, *sB = base + 2 * G8_A_BYTES;
    uint64_t *done = reinterpret_cast<uint64_t *>(base + 2 * (G8_A_BYTES + G8_B_BYTES));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pseudo-random operand bytes: the power a tensor instruction draws depends on how many operand
    // bits toggle, and the sustained rate is a power-limited figure
    for (size_t i = threadIdx.x; i < 2 * (G8_A_BYTES + G8_B_BYTES) / 16; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
            w[t] = h;
        }
        reinterpret_cast<uint4 *>(base)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (threadIdx.x == 0) {
        mb_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's async proxy
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "n"(2 * G8_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(G8_BN >> 3) << 17) | ((uint32_t)(G8_BM >> 4) << 24);
        for (int it = 0; it < iters; it++) {
            const int s = it & 1;  // two operand stages and two accumulators, like the pipelined kernel
            const uint64_t ad = umma_desc_sw128(sA + s * G8_A_BYTES), bd = umma_desc_sw128(sB + s * G8_B_BYTES);
            const uint32_t d_tmem = tmem + (uint32_t)(((it >> 5) & 1) * G8_BN);
#pragma unroll
            for (int k = 0; k < G8_BK / 32; k++) umma_i8(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (it & 31) | k ? 1u : 0u);
        }
        umma_commit(done);
        mb_wait(done, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * G8_BN) : "memory");
}

// the same loop back to back for `seconds`: the rate of the second half, i.e. under whatever
// clock the power limit settles at (a long GEMM-bound step runs against this, not the burst rate)
int bench_i8_peak_sustained(double seconds, double *tops, cudaStream_t st) {
    const size_t smem = 2 * (G8_A_BYTES + G8_B_BYTES) + 1024 + 64;
    CGPU_CUDA(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int iters = 1 << 15;
    const int launches = std::max(4, (int)(seconds / 2.6e-3)) / 2 * 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < launches; i++) {
        if (i == launches / 2) cudaEventRecord(a, st);
        i8_peak_kernel<<<sms, 64, smem, st>>>(iters);
    }
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    CGPU_LAUNCH_CHECK();
    *tops = 2.0 * G8_BM * G8_BN * G8_BK * (double)iters * sms * (launches / 2) / (ms * 1e-3) / 1e12;
    return CGPU_OK;
}

int bench_i8_peak(double *tops, cudaStream_t st) {
    const size_t smem = 2 * (G8_A_BYTES + G8_B_BYTES) + 1024 + 64;
    CGPU_CUDA(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int iters = 1 << 15;  // 32768 k-blocks of 128x256x128: ~2.5 ms at the datasheet rate
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a, st);
        i8_peak_kernel<<<sms, 64, smem, st>>>(iters);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    CGPU_LAUNCH_CHECK();
    *tops = 2.0 * G8_BM * G8_BN * G8_BK * (double)iters * sms / (best * 1e-3) / 1e12;
    return CGPU_OK;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return (EncodeTiledFn) nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// rows x NP bytes, row-major, box = box_rows x 128 bytes, 128B swizzle
static int make_map(CUtensorMap *map, const void *ptr, size_t rows, int NP, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error(CGPU_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
        return CGPU_ECUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)NP, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)NP};
    cuuint32_t box[2] = {(cuuint32_t)G8_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(CGPU_ECUDA, "cuTensorMapEncodeTiled failed with %d (rows=%zu NP=%d)", (int)r, rows, NP);
        return CGPU_ECUDA;
    }
    return CGPU_OK;
}

// ---- site patterns: identical columns of L
// two independent 64-bit mixes of a column's bit patterns (candidates only: every match is verified bit for bit)
__global__ void __launch_bounds__(256)
col_hash_kernel(const double *__restrict__ L, int K, int n, unsigned long long *__restrict__ h) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned long long a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full;
    for (int k = 0; k < K; k++) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(L[(size_t)k * n + j]);
        a = (a ^ u) * 0xBF58476D1CE4E5B9ull;
        a ^= a >> 29;
        b = (b + u) * 0x94D049BB133111EBull;
        b ^= b >> 31;
    }
    h[2 * (size_t)j] = a;
    h[2 * (size_t)j + 1] = b;
}
__global__ void __launch_bounds__(256)
col_verify_kernel(const double *__restrict__ L, int K, int n, const unsigned int *__restrict__ pid,
                  const unsigned int *__restrict__ rep, int *__restrict__ bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned int r = rep[pid[j]];
    if (r == (unsigned int)j) return;
    for (int k = 0; k < K; k++)
        if (__double_as_longlong(L[(size_t)k * n + j]) != __double_as_longlong(L[(size_t)k * n + r])) {
            atomicExch(bad, 1);
            return;
        }
}

// at most this many sites share a slot: a slot's count is Poisson(r * multiplicity) and must stay below 256
constexpr int FX_PATTERN_CAP = 32;

static int fx_find_patterns(DevMatrix &m, cudaStream_t st, std::vector<unsigned int> *rep_out) {
    const int n = m.n, K = m.K;
    unsigned long long *d_h = nullptr;
    CGPU_CUDA(dev_alloc(&d_h, (size_t)n * 16));
    col_hash_kernel<<<(n + 255) / 256, 256, 0, st>>>(m.d_L, K, n, d_h);
    CGPU_LAUNCH_CHECK();
    std::vector<unsigned long long> h((size_t)2 * n);
    CGPU_CUDA(cudaMemcpyAsync(h.data(), d_h, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    dev_free(d_h);
    // sites in order; a pattern opens a new slot every FX_PATTERN_CAP members
    struct Slot {
        unsigned int slot, members;
    };
    std::unordered_map<unsigned long long, Slot> open;  // keyed by the first mix; the second one is compared explicitly
    std::vector<unsigned long long> second;             // per slot: second mix of its representative
    std::vector<unsigned int> pid(n), rep;
    open.reserve((size_t)n);
    for (int j = 0; j < n; j++) {
        const unsigned long long a = h[2 * (size_t)j], b = h[2 * (size_t)j + 1];
        auto it = open.find(a);
        if (it != open.end() && second[it->second.slot] == b && it->second.members < (unsigned)FX_PATTERN_CAP) {
            pid[j] = it->second.slot;
            it->second.members++;
        } else {
            const unsigned int sl = (unsigned int)rep.size();
            rep.push_back((unsigned int)j);
            second.push_back(b);
            open[a] = Slot{sl, 1};
            pid[j] = sl;
        }
    }
    if (rep.size() * 8 > (size_t)n * 7) return CGPU_OK;  // less than 1/8 saved: not worth a table lookup per draw
    unsigned int *d_rep = nullptr;
    int *d_bad = nullptr;
    CGPU_CUDA(dev_alloc(&m.d_pid, (size_t)n * 4));
    CGPU_CUDA(dev_alloc(&d_rep, rep.size() * 4));
    CGPU_CUDA(dev_alloc(&d_bad, 4));
    CGPU_CUDA(cudaMemcpyAsync(m.d_pid, pid.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CGPU_CUDA(cudaMemcpyAsync(d_rep, rep.data(), rep.size() * 4, cudaMemcpyHostToDevice, st));
    CGPU_CUDA(cudaMemsetAsync(d_bad, 0, 4, st));
    col_verify_kernel<<<(n + 255) / 256, 256, 0, st>>>(m.d_L, K, n, m.d_pid, d_rep, d_bad);
    CGPU_LAUNCH_CHECK();
    int bad = 0;
    CGPU_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    dev_free(d_rep);
    dev_free(d_bad);
    if (bad) {  // a hash collision: two different columns shared both mixes -- no compression for this matrix
        dev_free(m.d_pid);
        m.d_pid = nullptr;
        return CGPU_OK;
    }
    m.npat = (int)rep.size();
    *rep_out = std::move(rep);
    return CGPU_OK;
}

// Parts of the site axis per tile of the pair GEMM: 1 unless the tile grid leaves most pairs idle
// while the rows are very long (C4: K=100, n=10^7 -> 4 x 5 tiles of 78125 k-blocks for 74 pairs)
static int g2_ksplit(const DevMatrix &m, int nb_pad) {
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int npairs = sms / 2;
    const long long tiles = (long long)(nb_pad / 256) * (long long)((m.fxMpad + 255) / 256);
    const int nkb = m.NPc / G8_BK;
    if (tiles <= 0 || tiles * 2 > npairs || nkb < 512 || G().opt_i8_ksplit == 0) return 1;
    int ks = (int)std::min<long long>({(long long)(2 * npairs / tiles), (long long)(nkb / 128), 32LL});
    if (G().opt_i8_ksplit > 0) ks = G().opt_i8_ksplit;
    return std::max(1, ks);
}

static bool fx_can_fuse(const DevMatrix &m, int max_draws) {
    // S <= 8 dividing 8, a scale 2^-F that is a normal double, and group sums that are exact in fp64
    return m.fxS <= 8 && 8 % m.fxS == 0 && G().opt_i8_fuse != 0 && m.fxF > -1000 && m.fxF < 1000 && max_draws < (1 << 21);
}
bool fx_needs_partial(const DevMatrix &m, int max_draws, int max_nb_pad) {
    static const bool env_one_shot = getenv("CGPU_I8_ONESHOT") != nullptr;
    const bool one_shot = G().opt_i8_oneshot >= 0 ? G().opt_i8_oneshot != 0 : env_one_shot;
    // (the smallest batch of a call, 256 replicates, is the one that may want the split)
    (void)max_nb_pad;
    return one_shot || !fx_can_fuse(m, max_draws) || g2_ksplit(m, 256) > 1 || G().opt_i8_ksplit > 1;
}

int fx_prepare(DevMatrix &m, cudaStream_t st, bool allow_patterns) {
    if (m.d_A8 && (m.npat == 0 || allow_patterns)) return CGPU_OK;
    if (m.d_A8) {  // prepared with pattern slots, but this call cannot use them: slice again, uncompressed
        dev_free(m.d_A8);
        dev_free(m.d_pid);
        m.d_A8 = nullptr;
        m.d_pid = nullptr;
        m.npat = 0;
    }
    if (m.ready) CGPU_CUDA(cudaStreamWaitEvent(st, m.ready, 0));  // the upload may still be in flight on the copy stream
    int *d_rng = nullptr;
    CGPU_CUDA(dev_alloc(&d_rng, 3 * sizeof(int)));
    const int init[3] = {1 << 30, -(1 << 30), 0};
    CGPU_CUDA(cudaMemcpyAsync(d_rng, init, sizeof(init), cudaMemcpyHostToDevice, st));
    fx_scan_kernel<<<G().sm_count * 4, 256, 0, st>>>(m.d_L, (size_t)m.K * m.n, d_rng);
    CGPU_LAUNCH_CHECK();
    int h[3];
    CGPU_CUDA(cudaMemcpyAsync(h, d_rng, sizeof(h), cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    dev_free(d_rng);
    if (h[2]) {
        set_error(CGPU_ERANGE, "CGPU_SUM_INT8 needs finite site log-likelihoods");
        return CGPU_ERANGE;
    }
    if (h[0] > h[1]) { h[0] = 0; h[1] = 0; }  // all-zero matrix
    // magnitude bits + sign + one spare bit so the carry of the balanced digits cannot
    // overflow the top slice
    const int bits = h[1] - h[0] + 3;
    const int S = (bits + 7) / 8;
    if (S > 12) {
        set_error(CGPU_ERANGE, "CGPU_SUM_INT8: the matrix spans %d bits (2^%d .. 2^%d); exact fixed point supports 96", bits, h[0], h[1]);
        return CGPU_ERANGE;
    }
    m.fxS = S < 1 ? 1 : S;
    m.fxF = -h[0];
    m.fxMpad = ((size_t)m.K * m.fxS + G8_BM - 1) / G8_BM * G8_BM;
    std::vector<unsigned int> rep;
    m.npat = 0;
    if (allow_patterns && G().opt_i8_patterns != 0 && m.n >= 256) CGPU_TRY(fx_find_patterns(m, st, &rep));
    m.NPc = m.npat > 0 ? pad_sites(m.npat) : m.NP;
    unsigned int *d_rep = nullptr;
    if (m.npat > 0) {
        CGPU_CUDA(dev_alloc(&d_rep, rep.size() * 4));
        CGPU_CUDA(cudaMemcpyAsync(d_rep, rep.data(), rep.size() * 4, cudaMemcpyHostToDevice, st));
    }
    CGPU_CUDA(dev_alloc(&m.d_A8, (size_t)m.fxMpad * m.NPc));
    CGPU_CUDA(cudaMemsetAsync(m.d_A8, 0, (size_t)m.fxMpad * m.NPc, st));
    fx_slice_kernel<<<G().sm_count * 8, 256, 0, st>>>(m.d_L, m.K, m.n, m.npat > 0 ? m.npat : m.n, d_rep, m.NPc, m.fxS, m.fxF, m.d_A8);
    CGPU_LAUNCH_CHECK();
    if (d_rep) {
        CGPU_CUDA(cudaStreamSynchronize(st));  // `rep` (pageable) has been consumed
        dev_free(d_rep);
    }
    return CGPU_OK;
}

// counts: plain rows [nb_pad][NP] uint8 (nb_pad multiple of 256, padding rows zero)
int launch_contract_i8(const DevMatrix &m, const unsigned char *d_counts_rows, int nb, int nb_pad, int *d_partial,
                       double r, double *d_rep, int ldrep, int col0, cudaStream_t st, PhaseTimer *pt, int max_draws) {
    if (nb <= 0) return CGPU_OK;
    // int32 accumulators: |sum_j c*d| <= 128 * (draws per replicate)
    CUtensorMap mapA, mapB;
    CGPU_TRY(make_map(&mapA, m.d_A8, m.fxMpad, m.NPc, G8_BM));
    CGPU_TRY(make_map(&mapB, d_counts_rows, (size_t)nb_pad, m.NPc, G8_BN));
    static bool attr_done = false;
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G8_SMEM));
        CGPU_CUDA(cudaFuncSetAttribute(gemm_i8_persistent_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G8P<4, 8>::smem));
        CGPU_CUDA(cudaFuncSetAttribute(gemm_i8_persistent_kernel<3, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G8P<3, 12>::smem));
        attr_done = true;
    }
    static const bool env_one_shot = getenv("CGPU_I8_ONESHOT") != nullptr;  // development A/B switch
    const bool one_shot = G().opt_i8_oneshot >= 0 ? G().opt_i8_oneshot != 0 : env_one_shot;
    // the recombination runs in the GEMM epilogue when the digits of a tree stay inside one
    // 32-row epilogue slab ("i8_fuse" 0 keeps the two-kernel form for measurements)
    // (S <= 8, a scale 2^-F that is a normal double, and group sums that are exact in fp64)
    const bool two_cta = !one_shot && (G().opt_i8_2cta == 1 || (G().opt_i8_2cta != 0 && m.NPc / G8_BK > 48)) &&
                         (G().sm_count % 2 == 0 || G().sm_count <= 0);
    const int ksplit = (two_cta && d_partial) ? g2_ksplit(m, nb_pad) : 1;
    const bool fused = !one_shot && fx_can_fuse(m, max_draws) && ksplit == 1;
    if (!fused && !d_partial) {
        set_error(CGPU_EINVAL, "launch_contract_i8: no partial-sum buffer for the unfused recombination");
        return CGPU_EINVAL;
    }
    if (pt) pt->begin(6, st);
    const int m_tiles = (int)(m.fxMpad / G8_BM), n_tiles = nb_pad / G8_BN;
    // cta_group::2 pairs for long tiles; short tiles (C2: 27 k-blocks) are bound by the epilogue's
    // divisions, where the one-CTA kernel with twelve epilogue warps is faster (measured 2.2 vs 3.0 ms)
    if (two_cta) {
        // cta_group::2: replicates on the M side (count rows are operand A), digit rows on N
        CUtensorMap mapC, mapD;
        CGPU_TRY(make_map(&mapC, d_counts_rows, (size_t)nb_pad, m.NPc, 128));
        CGPU_TRY(make_map(&mapD, m.d_A8, m.fxMpad, m.NPc, 128));
        static bool attr2_done = false;
        if (!attr2_done) {
            CGPU_CUDA(cudaFuncSetAttribute(gemm_i8_2cta_kernel<6, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2P<6, 8>::smem));
            attr2_done = true;
        }
        const int c_tiles = nb_pad / 256, d_tiles = (int)((m.fxMpad + 255) / 256);
        const int sms = G().sm_count > 0 ? G().sm_count : 148;
        const int grid = std::min(sms / 2 * 2, 2 * c_tiles * d_tiles * ksplit);
        if (ksplit > 1) CGPU_CUDA(cudaMemsetAsync(d_partial, 0, m.fxMpad * (size_t)nb_pad * sizeof(int), st));  // the parts add into it
        FxEpilogue ep{fused ? d_rep : nullptr, r, {(m.fxF > -1000 && m.fxF < 1000) ? std::ldexp(1.0, -m.fxF) : 0.0, m.fxF},
                      m.K, m.fxS, nb, ldrep, col0, G().opt_i8_dbg > 0 ? G().opt_i8_dbg : 0};
        gemm_i8_2cta_kernel<6, 8><<<grid, G2P<6, 8>::threads, G2P<6, 8>::smem, st>>>(mapC, mapD, m.NPc, d_partial, nb_pad, (int)m.fxMpad,
                                                                                     c_tiles, d_tiles,
                                                                                     G().opt_i8_panel >= 0 ? G().opt_i8_panel : 8, ksplit, ep);
    } else if (one_shot) {
        dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
        gemm_i8_kernel<<<grid, G8_THREADS, G8_SMEM, st>>>(mapA, mapB, m.NPc, d_partial, nb_pad, nb_pad);
    } else {
        const int sms = G().sm_count > 0 ? G().sm_count : 148;
        const int grid = std::min(sms, m_tiles * n_tiles);
        FxEpilogue ep{fused ? d_rep : nullptr, r, {(m.fxF > -1000 && m.fxF < 1000) ? std::ldexp(1.0, -m.fxF) : 0.0, m.fxF},
                      m.K, m.fxS, nb, ldrep, col0, G().opt_i8_dbg > 0 ? G().opt_i8_dbg : 0};
        // short tiles (C2: 27 k-blocks) are epilogue bound, long ones feed bound (measured cross-over ~50)
        if (fused && m.NPc / G8_BK <= 48)
            gemm_i8_persistent_kernel<3, 12><<<grid, G8P<3, 12>::threads, G8P<3, 12>::smem, st>>>(mapA, mapB, m.NPc, d_partial, nb_pad, m_tiles, n_tiles, ep);
        else
            gemm_i8_persistent_kernel<4, 8><<<grid, G8P<4, 8>::threads, G8P<4, 8>::smem, st>>>(mapA, mapB, m.NPc, d_partial, nb_pad, m_tiles, n_tiles, ep);
    }
    CGPU_LAUNCH_CHECK();
    if (pt) pt->end();
    if (fused) return CGPU_OK;
    if (pt) pt->begin(7, st);
    dim3 cgrid((nb + 255) / 256, m.K);
    fx_combine_kernel<<<cgrid, 256, 0, st>>>(d_partial, nb_pad, m.K, m.fxS, m.fxF, nb, r, d_rep, ldrep, col0);
    CGPU_LAUNCH_CHECK();
    if (pt) pt->end();
    return CGPU_OK;
}

}  // namespace cgpu
