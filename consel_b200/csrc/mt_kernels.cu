// MT19937 on the device: polynomial jump-ahead and chunked block regeneration that
// reproduce the reference's process-global uniform stream (src/mt19937.c:57-60,
// 108-127, 182-197; consumed by mrandlist, src/rand.c:25-29) and turn each draw
// into the site index of src/makermt.c:55-60.
//
// Stream model.  x_0..x_623 is the seeded array (init_genrand, src/mt19937.c:63-76),
// x_{t+624} = x_{t+397} ^ twist(x_t, x_{t+1}); draw d (0-based) is temper(x_{624+d}).
// "Window" W_d = (x_d .. x_{d+623}) is everything needed to produce draws d, d+1, ...
//
// Kernels
//   mt_jump_kernel   W_{d+J} = g(F) W_d for g = x^J mod phi: regenerate 19937+624
//                    words from the source window into shared memory, then every
//                    output word j is the XOR of x_{i+j} over the set bits i of g
//                    (walked as a position list).  The list is sliced over gridDim.y
//                    CTAs (atomicXor combine) so early rounds of the doubling tree
//                    still fill the SMs.
//   mt_draw_kernel   227 lanes per chunk of 2^m draws: block regeneration with the
//                    three dependent phases of next_state chained in registers (one
//                    barrier per 624 draws), tempering, site index, coalesced stores.
#include <cmath>

#include "cgpu_internal.h"

namespace cgpu {

__device__ __forceinline__ uint32_t mt_twist(uint32_t hi, uint32_t lo) {
    uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return (y >> 1) ^ ((lo & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// The draw kernel is bound by the integer-logic pipe (LOP3/SHF issue every other cycle per
// scheduler; the multiplier pipe is a separate one with the same rate), so its copies of the
// two functions do their right shifts as high halves of products with a power of two.  The
// multiplier is read from a register the compiler cannot see through (strength reduction
// would turn it back into a shift).
__device__ __forceinline__ uint32_t shr_mul(uint32_t v, uint32_t pow2_32_minus_s) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(pow2_32_minus_s));
    return r;
}

struct DrawConsts { uint32_t p31, p21, p14, mag; };  // 2^(32-1), 2^(32-11), 2^(32-18), 0x9908b0df (opaque multipliers)

__device__ __forceinline__ uint32_t mt_twist_bal(uint32_t hi, uint32_t lo, const DrawConsts &k) {
    uint32_t y;  // (hi & 0x80000000) | (lo & 0x7fffffff): bit select as ONE lop3 (LUT 0xe4 = mask ? a : b)
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xe4;" : "=r"(y) : "r"(hi), "r"(lo));
    const uint32_t y1 = shr_mul(y, k.p31);  // y >> 1 on the multiplier pipe
    uint32_t mag;                           // (lo & 1) * 0x9908b0df with the constant in a register: stays a multiply
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(mag) : "r"(lo & 1u), "r"(k.mag));
    return y1 ^ mag;
}

// tempering as the draw kernel applies it: the plain shift/xor form of mt_temper (moving its left
// shifts to the multiplier pipe as well measured slower -- exactly one multiply per twist is the
// optimum); the hook keeps the call sites of the balanced twist and the
// tempering symmetric
__device__ __forceinline__ uint32_t mt_temper_bal(uint32_t y, const DrawConsts &) { return mt_temper(y); }

// ------------------------------------------------------------------ jump
// The polynomial arrives as the list of its set bit positions, two per 32-bit word
// (pos_a | pos_b << 17), padded to a multiple of 8 with the position of a block of zero
// words.  A CTA copies its slice of the list into shared memory once (streaming it from
// L2 inside the loop stalls every warp on the same miss), then the inner loop is: one
// broadcast 16-byte shared load = 8 positions, per position two address ops, two shared
// loads (each lane owns output words tid and tid+312) and half a 3-input XOR.  That is
// shared-memory bandwidth bound (624 x ~10^4 word reads per jump).
constexpr int JUMP_T = MT_N / 2;                  // 312 lanes, two output words each
constexpr int JUMP_THREADS = 320;
constexpr int JUMP_SEQ = MT_DEG + MT_N + 63;      // words of the regenerated sequence kept in smem
constexpr int JUMP_ZERO = JUMP_SEQ + 1;           // word index of MT_N zero words (list padding target)
constexpr int JUMP_LIST_OCT = 1536;               // list capacity in smem: 1536 x 8 positions per pass
constexpr int JUMP_SMEM_WORDS = JUMP_ZERO + MT_N + 3 /* align */ + JUMP_LIST_OCT * 4;

__global__ void __launch_bounds__(JUMP_THREADS, 2)
mt_jump_kernel(uint32_t *states, const int2 *__restrict__ pairs,
               const uint4 *__restrict__ pos8, int octs_per_slice, int nocts) {
    extern __shared__ __align__(16) uint32_t x[];
    uint4 *plist = reinterpret_cast<uint4 *>(x + ((JUMP_ZERO + MT_N + 3) & ~3));
    const int tid = threadIdx.x;
    const int2 pr = pairs[blockIdx.x];
    const uint32_t *w = states + (size_t)pr.x * MT_N;
    uint32_t *o = states + (size_t)pr.y * MT_N;

    const int q_lo = blockIdx.y * octs_per_slice;
    const int q_hi = min(q_lo + octs_per_slice, nocts);
    if (q_lo >= q_hi) return;
    // positions ascend, so the last real entry of the slice bounds the sequence to regenerate
    int need;
    {
        const uint4 last = pos8[q_hi - 1];
        const uint32_t v[4] = {last.x, last.y, last.z, last.w};
        uint32_t mx = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t a = v[k] & 0x7fffu, b = v[k] >> 17;
            if (a < (uint32_t)JUMP_ZERO) mx = max(mx, a);
            if (b < (uint32_t)JUMP_ZERO) mx = max(mx, b);
        }
        need = (int)mx + 1 + MT_N;  // sequence words x[0..need)
    }
    for (int i = tid; i < MT_N; i += JUMP_THREADS) {
        x[i] = w[i];
        x[JUMP_ZERO + i] = 0;
    }
    const int first_pass = min(q_hi - q_lo, JUMP_LIST_OCT);
    for (int i = tid; i < first_pass; i += JUMP_THREADS) plist[i] = pos8[q_lo + i];
    __syncthreads();
    // x[t+624] depends on x[t+397]: 227 new words per step are independent
    for (int t0 = 0; t0 + MT_N < need; t0 += MT_N - MT_M) {
        int t = t0 + tid;
        if (tid < MT_N - MT_M && t + MT_N < need)
            x[t + MT_N] = x[t + MT_M] ^ mt_twist(x[t], x[t + 1]);
        __syncthreads();
    }
    uint32_t r0 = 0, r1 = 0;
    const char *xb = reinterpret_cast<const char *>(x + (tid < JUMP_T ? tid : 0));
    for (int q0 = q_lo; q0 < q_hi; q0 += JUMP_LIST_OCT) {
        const int cnt = min(q_hi - q0, JUMP_LIST_OCT);
        if (q0 != q_lo) {  // polynomials denser than the smem list: further passes
            __syncthreads();
            for (int i = tid; i < cnt; i += JUMP_THREADS) plist[i] = pos8[q0 + i];
            __syncthreads();
        }
#pragma unroll 2
        for (int q = 0; q < cnt; q++) {
            const uint4 p = plist[q];
            const uint32_t v[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const char *pa = xb + ((v[k] & 0x7fffu) << 2);
                const char *pb = xb + ((v[k] >> 17) << 2);
                const uint32_t a0 = *reinterpret_cast<const uint32_t *>(pa);
                const uint32_t a1 = *reinterpret_cast<const uint32_t *>(pa + JUMP_T * 4);
                const uint32_t b0 = *reinterpret_cast<const uint32_t *>(pb);
                const uint32_t b1 = *reinterpret_cast<const uint32_t *>(pb + JUMP_T * 4);
                r0 ^= a0 ^ b0;
                r1 ^= a1 ^ b1;
            }
        }
    }
    if (tid < JUMP_T) {
        if (gridDim.y == 1) {
            o[tid] = r0;
            o[tid + JUMP_T] = r1;
        } else {
            if (r0) atomicXor(&o[tid], r0);
            if (r1) atomicXor(&o[tid + JUMP_T], r1);
        }
    }
}

// set bit positions of a residue in the packed layout above; returns the number of 8-position groups
int mt_poly_positions(const Gf2Poly *g, uint32_t *out) {
    int np = 0;
    auto put = [&](uint32_t pos) {
        if (np & 1) out[np >> 1] |= pos << 17;
        else out[np >> 1] = pos;
        np++;
    };
    for (int i = 0; i < MT_DEG; i++)
        if (g->w[i >> 6] >> (i & 63) & 1) put((uint32_t)i);
    while (np & 7) put((uint32_t)JUMP_ZERO);
    return np / 8;
}

// destination slots must be zero on entry (sliced launches combine with atomicXor)
int launch_mt_jump(uint32_t *d_states, const int2 *d_pairs, const uint32_t *d_pos, int nocts, int count,
                   cudaStream_t st) {
    if (count <= 0) return CGPU_OK;
    static bool attr_done = false;
    const size_t smem = (size_t)JUMP_SMEM_WORDS * sizeof(uint32_t);
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    if (nocts <= 0) return CGPU_OK;  // the zero polynomial cannot occur (x^J mod phi is never 0)
    // slice the position list over enough CTAs to cover the SMs when few windows jump
    // Slices of the position list per jump: enough CTAs to fill the SMs while the tree is narrow,
    // and for wide levels the count that wastes the least of the last wave (2 CTAs per SM are
    // resident; every extra slice re-generates part of the sequence, ~4 % of a jump).
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int slots = 2 * sms;
    int slices = 1;
    double best = 1e300;
    for (int sl = 1; sl <= 16; sl++) {
        const double waves = std::ceil((double)count * sl / slots) / sl;
        const double cost = waves * (1.0 + 0.04 * (sl - 1));
        if (cost < best - 1e-9) {
            best = cost;
            slices = sl;
        }
    }
    int ops = (nocts + slices - 1) / slices;
    mt_jump_kernel<<<dim3(count, slices), JUMP_THREADS, smem, st>>>(d_states, d_pairs, reinterpret_cast<const uint4 *>(d_pos),
                                                                    ops, nocts);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// ------------------------------------------------------------------ draws

// site index of a draw: (int)((double)y * 2^-32 * n) with the reference's clamp
// (src/mt19937.c:195, src/makermt.c:55-60).  For n < 2^21 the product y*n is exact
// in fp64, so the truncation equals the integer high word of y*n.
// PATH 0: integer, 1: fp64, 2: raw tempered word (cgpu_mt19937_stream)
template <int PATH>
__device__ __forceinline__ uint32_t site_index(uint32_t y, uint32_t n, double dn) {
    if (PATH == 2) return y;
    if (PATH == 0) return __umulhi(y, n);
    double rv = __dmul_rn((double)y, 1.0 / 4294967296.0);
    int k = (int)__dmul_rn(rv, dn);
    if (k < 0) k = 0;
    else if (k >= (int)n) k = (int)n - 1;
    return (uint32_t)k;
}

// Block regeneration with the dependency chain in registers.  Lane t (0..226) of a chunk
// owns words t, t+227, t+454 of the 624-word block.  new[t] needs only old words;
// new[t+227] = new[t] ^ twist(old[t+227], old[t+228]) and new[t+454] = new[t+227] ^
// twist(old[t+454], old[t+455]) chain through the SAME lane, so the three "phases" of
// next_state need no barrier between them; the only cross-lane exception (word 623 wants
// new[0]) is recomputed by lane 169.  The lane keeps its three words in registers from one
// block to the next, reads its four neighbour words from the ping-pong shared buffer, and
// tempers + stores the indices straight from registers: one barrier per block.
//
// CPC chunks are packed 227 lanes apiece into one CTA.  The draws to keep are given as up
// to 32 disjoint stream ranges [lo, hi) (one per scale: the replicate columns this rank
// owns), each mapped to an offset of the compact index buffer; the chunks to run are the
// union of the chunk intervals covering those ranges.
constexpr int DRAW_H = MT_N - MT_M;  // 227
template <int CPC> struct DrawCfg {
    static constexpr int threads = (CPC * DRAW_H + 31) / 32 * 32;
    static constexpr int ctas = 6 / CPC;  // resident CTAs per SM the register budget is held to (6 chunks, <= 40 registers)
};

template <int PATH, typename IdxT>
struct DrawLane {
    uint32_t r0, r1, r2;      // words t, t+227, t+454 of the current block
    unsigned int fast_until;  // blocks below this index take the unconditional store (or skip) path
    int rcur;
    bool inside, third;
    IdxT *outp;
};

// regeneration of one block: registers r0..r2 become words t, t+227, t+454 of the new block
template <int PATH, typename IdxT, int CUR>
__device__ __forceinline__ void draw_regen(DrawLane<PATH, IdxT> &L, uint32_t *bt, int t, const DrawConsts &kc) {
    constexpr int H = DRAW_H;
    const uint32_t *o = bt + CUR * MT_N;  // bt = &buf[c][0][t]; the other half is MT_N words further
    uint32_t *nw = bt + (CUR ^ 1) * MT_N;
    L.r0 = o[MT_M] ^ mt_twist_bal(L.r0, o[1], kc);
    L.r1 = L.r0 ^ mt_twist_bal(L.r1, o[H + 1], kc);
    nw[0] = L.r0;
    nw[H] = L.r1;
    if (L.third) {
        uint32_t nxt;
        if (t == MT_N - 2 * H - 1) nxt = o[MT_M - t] ^ mt_twist_bal(o[-t], o[1 - t], kc);  // new[0]
        else nxt = o[2 * H + 1];
        L.r2 = L.r1 ^ mt_twist_bal(L.r2, nxt, kc);
        nw[2 * H] = L.r2;
    }
}

// the three site indices of a block that lies entirely inside a kept range
template <int PATH, typename IdxT>
__device__ __forceinline__ void draw_emit(const DrawLane<PATH, IdxT> &L, IdxT *out, uint32_t n, double dn, const DrawConsts &kc) {
    constexpr int H = DRAW_H;
    out[0] = (IdxT)site_index<PATH>(mt_temper_bal(L.r0, kc), n, dn);
    out[H] = (IdxT)site_index<PATH>(mt_temper_bal(L.r1, kc), n, dn);
    if (L.third) out[2 * H] = (IdxT)site_index<PATH>(mt_temper_bal(L.r2, kc), n, dn);
}

// one block (624 draws) of one chunk; CUR = which half of the ping-pong buffer holds the old block
template <int PATH, typename IdxT, int CUR>
__device__ __forceinline__ void draw_step(DrawLane<PATH, IdxT> &L, uint32_t *bt, int t, unsigned int it, int c,
                                          const unsigned long long *s_begin, const unsigned long long *s_end,
                                          const DrawPlan &plan, uint32_t n, double dn, IdxT *idx, const DrawConsts &kc) {
    constexpr int H = DRAW_H;
    draw_regen<PATH, IdxT, CUR>(L, bt, t, kc);
    if (it < L.fast_until) {
        if (L.inside) {
            draw_emit<PATH, IdxT>(L, L.outp, n, dn, kc);
            L.outp += MT_N;
        }
        return;
    }
    // slow path (a few times per chunk): block touching a range boundary or the chunk end
    const unsigned long long d_end = s_end[c];
    const unsigned long long d0 = s_begin[c] + (unsigned long long)it * MT_N;
    if (d0 >= d_end) {
        L.fast_until = ~0u;
        L.inside = false;
        return;
    }
    const unsigned long long d1 = d_end - d0 >= (unsigned long long)MT_N ? d0 + MT_N : d_end;
    int rcur = L.rcur;
    while (rcur < plan.nranges && d0 >= plan.hi[rcur]) rcur++;
    for (int r = rcur; r < plan.nranges && plan.lo[r] < d1; r++) {
        const unsigned long long lo = plan.lo[r], hi = plan.hi[r];
        const int i_lo = d0 >= lo ? 0 : (int)(lo - d0);
        const int i_hi = (int)((d1 < hi ? d1 : hi) - d0);
        IdxT *out = idx + plan.off[r] + (long long)(d0 - lo) + t;  // only [i_lo, i_hi) is touched
        if (t >= i_lo && t < i_hi) out[0] = (IdxT)site_index<PATH>(mt_temper_bal(L.r0, kc), n, dn);
        if (t + H >= i_lo && t + H < i_hi) out[H] = (IdxT)site_index<PATH>(mt_temper_bal(L.r1, kc), n, dn);
        if (L.third && t + 2 * H >= i_lo && t + 2 * H < i_hi)
            out[2 * H] = (IdxT)site_index<PATH>(mt_temper_bal(L.r2, kc), n, dn);
    }
    // plan the following blocks: how many of them are entirely inside / outside
    const unsigned long long dn0 = d0 + MT_N;
    while (rcur < plan.nranges && dn0 >= plan.hi[rcur]) rcur++;
    L.fast_until = it + 1;
    L.inside = false;
    if (rcur < plan.nranges && dn0 < d_end) {
        if (plan.lo[rcur] <= dn0) {
            const unsigned long long lim = plan.hi[rcur] < d_end ? plan.hi[rcur] : d_end;
            L.fast_until = it + 1 + (unsigned int)((lim - dn0) / MT_N);
            L.inside = true;
            L.outp = idx + plan.off[rcur] + (dn0 - plan.lo[rcur]) + t;
        } else {
            const unsigned long long lim = plan.lo[rcur] < d_end ? plan.lo[rcur] : d_end;
            L.fast_until = it + 1 + (unsigned int)((lim - dn0) / MT_N);
        }
    } else if (rcur >= plan.nranges) {
        L.fast_until = ~0u;  // nothing left to keep in this chunk
    }
    L.rcur = rcur;
}

template <int PATH, typename IdxT, int CPC>
__global__ void __launch_bounds__(DrawCfg<CPC>::threads, DrawCfg<CPC>::ctas)
mt_draw_kernel(const uint32_t *__restrict__ states, unsigned long long chunk_len, DrawPlan plan, uint32_t n, IdxT *__restrict__ idx,
               DrawConsts kc) {
    __shared__ uint32_t buf[CPC + 1][2][MT_N];  // +1: the padding lanes of the last warp index a dummy slot
    __shared__ unsigned long long s_begin[CPC + 1], s_end[CPC + 1];  // read on the slow path only
    constexpr int H = DRAW_H;
    const int c = threadIdx.x / H, t = threadIdx.x - c * H;
    const double dn = (double)n;
    const unsigned int steps = (unsigned int)((chunk_len + MT_N - 1) / MT_N);

    DrawLane<PATH, IdxT> L;
    L.r0 = L.r1 = L.r2 = 0;
    L.third = t < MT_N - 2 * H;  // lanes 0..169 own a third word
    L.rcur = 0;
    L.fast_until = 0;
    L.inside = false;
    L.outp = idx;
    uint32_t *bt = &buf[c][0][t];
    {
        unsigned long long d_begin = 0, d_end = 0;
        const int ci = blockIdx.x * CPC + c;
        if (c < CPC && ci < plan.cum[plan.nintervals]) {
            int iv = 0;
            while (ci >= plan.cum[iv + 1]) iv++;
            const unsigned long long chunk = plan.first[iv] + (unsigned long long)(ci - plan.cum[iv]);
            d_begin = chunk * chunk_len;
            d_end = d_begin + chunk_len;
            if (d_end > plan.stream_end) d_end = plan.stream_end;
            if (d_end > d_begin) {
                const uint32_t *w = states + (size_t)(plan.slot[iv] + (ci - plan.cum[iv])) * MT_N + t;
                L.r0 = w[0];
                L.r1 = w[H];
                bt[0] = L.r0;
                bt[H] = L.r1;
                if (L.third) {
                    L.r2 = w[2 * H];
                    bt[2 * H] = L.r2;
                }
            } else {
                d_end = d_begin;
            }
        }
        if (t == 0) {
            s_begin[c] = d_begin;
            s_end[c] = d_end;
        }
    }
    __syncthreads();
    // One barrier per block: every read of the old half and write of the new half is done
    // before the next step swaps their roles.  Lanes without a chunk (padding of the last
    // warp, chunks past the end) run the same code on zeros in their own buffer slot: their
    // bounds are empty, so the first slow-path visit turns their stores off for good.
    // Warps take different paths (each chunk has its own ranges), but every path does exactly
    // one barrier after each block, so barrier k always separates block k from block k+1.  Four blocks that all lie
    // inside the current range go through a branch-free body: no per-block range test, stores
    // at immediate offsets from one pointer.
    for (unsigned int it = 0; it < steps;) {
        // (warp-uniform decision: a warp that straddles two chunks must not split over the barriers)
        if (__all_sync(0xffffffffu, L.inside && it + 4 <= L.fast_until && it + 4 <= steps && (it & 1u) == 0)) {
            draw_regen<PATH, IdxT, 0>(L, bt, t, kc);
            draw_emit<PATH, IdxT>(L, L.outp, n, dn, kc);
            __syncthreads();
            draw_regen<PATH, IdxT, 1>(L, bt, t, kc);
            draw_emit<PATH, IdxT>(L, L.outp + MT_N, n, dn, kc);
            __syncthreads();
            draw_regen<PATH, IdxT, 0>(L, bt, t, kc);
            draw_emit<PATH, IdxT>(L, L.outp + 2 * MT_N, n, dn, kc);
            __syncthreads();
            draw_regen<PATH, IdxT, 1>(L, bt, t, kc);
            draw_emit<PATH, IdxT>(L, L.outp + 3 * MT_N, n, dn, kc);
            __syncthreads();
            L.outp += 4 * MT_N;
            it += 4;
        } else {
            if (it & 1u) draw_step<PATH, IdxT, 1>(L, bt, t, it, c, s_begin, s_end, plan, n, dn, idx, kc);
            else draw_step<PATH, IdxT, 0>(L, bt, t, it, c, s_begin, s_end, plan, n, dn, idx, kc);
            __syncthreads();
            it++;
        }
    }
}

// ------------------------------------------------------------------ draws + site histograms, fused
// Same block regeneration as mt_draw_kernel, but the site indices never leave the SM: every
// chunk (227 lanes) keeps two 32-bit histograms of NP sites in shared memory, adds its draws
// with shared-memory atomics, and when a replicate (bn consecutive draws of one kept range)
// completes, packs the counters to bytes and writes the NP-byte row of the count matrix.
// A block of 624 draws holds at most one replicate end because bn >= 2*624 is required, so
// two histograms (the replicate being finished / the one that has begun) are enough; the row
// of a finished replicate is flushed after the block's barrier while the next block already
// counts into the other histogram, and the flushed one is reused two blocks later at the
// earliest.  Replicates cut by a chunk boundary are flushed with 32-bit atomic adds on the
// packed byte row (both halves land in the same zero-initialised row).
struct HistPlan {
    unsigned int bn[32];        // draws per replicate of range r
    unsigned char *row0[32];    // count row of the first kept replicate of range r (rows of NP bytes)
    int NP;
    int *err;                   // set when a counter exceeds 255
};

template <int CPC>
__global__ void __launch_bounds__(DrawCfg<CPC>::threads, DrawCfg<CPC>::ctas)
mt_draw_hist_kernel(const uint32_t *__restrict__ states, unsigned long long chunk_len, DrawPlan plan, HistPlan hp, uint32_t n,
                    DrawConsts kc) {
    extern __shared__ __align__(16) uint32_t dh_smem[];
    constexpr int H = DRAW_H;
    const int NP = hp.NP;
    uint32_t *buf_base = dh_smem;                               // [CPC+1][2][MT_N]
    uint32_t *hist_base = dh_smem + (CPC + 1) * 2 * MT_N;       // [CPC][2][NP]; the padding lanes alias group 0 (they only zero it)
    const int c = threadIdx.x / H, t = threadIdx.x - c * H;
    const bool third = t < MT_N - 2 * H;
    const unsigned int steps = (unsigned int)((chunk_len + MT_N - 1) / MT_N);
    uint32_t *bt = buf_base + (size_t)c * 2 * MT_N + t;
    uint32_t *hist = hist_base + (size_t)(c < CPC ? c : 0) * 2 * NP;

    uint32_t r0 = 0, r1 = 0, r2 = 0;
    unsigned long long d_begin = 0, d_end = 0;
    {
        const int ci = blockIdx.x * CPC + c;
        if (c < CPC && ci < plan.cum[plan.nintervals]) {
            int iv = 0;
            while (ci >= plan.cum[iv + 1]) iv++;
            const unsigned long long chunk = plan.first[iv] + (unsigned long long)(ci - plan.cum[iv]);
            d_begin = chunk * chunk_len;
            d_end = d_begin + chunk_len;
            if (d_end > plan.stream_end) d_end = plan.stream_end;
            if (d_end > d_begin) {
                const uint32_t *w = states + (size_t)(plan.slot[iv] + (ci - plan.cum[iv])) * MT_N + t;
                r0 = w[0];
                r1 = w[H];
                bt[0] = r0;
                bt[H] = r1;
                if (third) {
                    r2 = w[2 * H];
                    bt[2 * H] = r2;
                }
            } else {
                d_end = d_begin;
            }
        }
    }
    for (int i = t; i < 2 * NP; i += H) hist[i] = 0;

    // replicate cursor of this chunk (uniform over the chunk's lanes)
    int rcur = 0;                        // range the cursor is in (or the next one ahead of it)
    unsigned long long seg_lo = 0;       // first kept draw of the current replicate inside this chunk
    unsigned long long rep_end = 0;      // stream position where the current replicate ends
    unsigned char *row = nullptr;        // its count row
    bool partial = false;                // the replicate began before this chunk: flush with atomics
    bool have = false;                   // a current replicate exists
    int hcur = 0;
    auto seek = [&](unsigned long long pos) {  // position the cursor on the replicate containing / following pos
        while (rcur < plan.nranges && pos >= plan.hi[rcur]) rcur++;
        have = rcur < plan.nranges && plan.lo[rcur] < d_end;
        if (!have) return;
        const unsigned long long lo = plan.lo[rcur];
        const unsigned long long p0 = pos > lo ? pos : lo;
        const unsigned long long rep = (p0 - lo) / hp.bn[rcur];
        const unsigned long long rstart = lo + rep * hp.bn[rcur];
        seg_lo = p0;
        rep_end = rstart + hp.bn[rcur];
        row = hp.row0[rcur] + rep * (unsigned long long)NP;
        partial = rstart < p0;
    };
    seek(d_begin);
    // pack one histogram into its byte row and clear it
    auto flush = [&](uint32_t *h, unsigned char *dst, bool atomic) {
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst);
        uint4 *h4 = reinterpret_cast<uint4 *>(h);
        bool over = false;
        for (int w = t; w < NP / 4; w += H) {
            const uint4 v = h4[w];
            h4[w] = make_uint4(0, 0, 0, 0);
            over |= (v.x | v.y | v.z | v.w) > 255u;
            const uint32_t packed = (v.x & 255u) | (v.y & 255u) << 8 | (v.z & 255u) << 16 | (v.w & 255u) << 24;
            if (atomic) {
                if (packed) atomicAdd(&dst32[w], packed);
            } else {
                dst32[w] = packed;
            }
        }
        if (over) atomicExch(hp.err, 1);
    };
    __syncthreads();

    unsigned int fast_until = 0;  // blocks below this index are uniform: all inside the current replicate, or all skipped
    bool counting = false;
    auto plan_fast = [&](unsigned int it_next) {
        fast_until = it_next;
        counting = false;
        if (!have) {
            fast_until = ~0u;  // nothing left to count in this chunk
            return;
        }
        const unsigned long long b0 = d_begin + (unsigned long long)it_next * MT_N;
        const unsigned long long lim = rep_end < d_end ? rep_end : d_end;
        // all differences are below the chunk length (< 2^31): 32-bit divisions by a constant
        if (b0 >= seg_lo) {
            if (lim > b0) {  // full blocks inside [seg_lo, min(rep_end, d_end))
                fast_until = it_next + (unsigned int)(lim - b0) / (unsigned int)MT_N;
                counting = true;
            }
        } else {  // full blocks before the first kept draw
            fast_until = it_next + (unsigned int)(seg_lo - b0) / (unsigned int)MT_N;
        }
    };
    plan_fast(0);

    // regeneration of the next block (registers r0..r2 become its words t, t+227, t+454)
    auto regen = [&](const uint32_t *o, uint32_t *nw) {
        r0 = o[MT_M] ^ mt_twist_bal(r0, o[1], kc);
        r1 = r0 ^ mt_twist_bal(r1, o[H + 1], kc);
        nw[0] = r0;
        nw[H] = r1;
        if (third) {
            uint32_t nxt;
            if (t == MT_N - 2 * H - 1) nxt = o[MT_M - t] ^ mt_twist_bal(o[-t], o[1 - t], kc);  // new[0]
            else nxt = o[2 * H + 1];
            r2 = r1 ^ mt_twist_bal(r2, nxt, kc);
            nw[2 * H] = r2;
        }
    };
    // histogram updates of block `it`, whose words are still in r0..r2.  They are issued at the
    // START of the following step, so the shared-memory atomics drain while the next block is
    // regenerated instead of in front of the barrier.  Returns true when the current replicate
    // ends inside the block (flush after the next barrier).
    auto count = [&](unsigned int it) {
        if (it < fast_until) {
            if (counting) {
                uint32_t *h = hist + hcur * NP;
                atomicAdd(&h[__umulhi(mt_temper(r0), n)], 1u);
                atomicAdd(&h[__umulhi(mt_temper(r1), n)], 1u);
                if (third) atomicAdd(&h[__umulhi(mt_temper(r2), n)], 1u);
            }
            return false;
        }
        // boundary block: draws of the finishing replicate go to hist[hcur], draws of the one that
        // begins (same range, or the next range after a gap) to hist[hcur^1].  Offsets are
        // relative to the block start (32-bit).
        const unsigned long long d0 = d_begin + (unsigned long long)it * MT_N;
        if (d0 >= d_end) return false;
        const int len = d_end - d0 >= (unsigned long long)MT_N ? MT_N : (int)(d_end - d0);
        const int a_lo = seg_lo > d0 ? (seg_lo - d0 < (unsigned long long)len ? (int)(seg_lo - d0) : len) : 0;
        const int a_hi = rep_end - d0 < (unsigned long long)len ? (int)(rep_end - d0) : len;
        int b_lo = len;  // bn >= 2*624: the replicate that begins here cannot end inside this block
        if (a_hi < len) {
            if (rep_end < plan.hi[rcur]) b_lo = a_hi;
            else if (rcur + 1 < plan.nranges && plan.lo[rcur + 1] - d0 < (unsigned long long)len) b_lo = (int)(plan.lo[rcur + 1] - d0);
        }
        uint32_t *ha = hist + hcur * NP, *hb = hist + (hcur ^ 1) * NP;
        const int k0 = t, k1 = t + H, k2 = t + 2 * H;
        if (k0 >= a_lo && k0 < a_hi) atomicAdd(&ha[__umulhi(mt_temper(r0), n)], 1u);
        else if (k0 >= b_lo && k0 < len) atomicAdd(&hb[__umulhi(mt_temper(r0), n)], 1u);
        if (k1 >= a_lo && k1 < a_hi) atomicAdd(&ha[__umulhi(mt_temper(r1), n)], 1u);
        else if (k1 >= b_lo && k1 < len) atomicAdd(&hb[__umulhi(mt_temper(r1), n)], 1u);
        if (third) {
            if (k2 >= a_lo && k2 < a_hi) atomicAdd(&ha[__umulhi(mt_temper(r2), n)], 1u);
            else if (k2 >= b_lo && k2 < len) atomicAdd(&hb[__umulhi(mt_temper(r2), n)], 1u);
        }
        return rep_end <= d0 + (unsigned long long)len;  // the current replicate is complete: flush after the barrier
    };
    auto after = [&](bool done, unsigned int it) {  // bookkeeping once block `it` has been counted and a barrier passed
        if (it < fast_until) return;
        if (done) {
            flush(hist + hcur * NP, row, partial);
            hcur ^= 1;
            partial = false;
            if (rep_end < plan.hi[rcur]) {  // next replicate of the same range
                seg_lo = rep_end;
                rep_end += hp.bn[rcur];
                row += NP;
                have = seg_lo < d_end;
            } else {  // ranges are whole replicates: the next range starts a replicate
                rcur++;
                have = rcur < plan.nranges && plan.lo[rcur] < d_end;
                if (have) {
                    seg_lo = plan.lo[rcur];
                    rep_end = seg_lo + hp.bn[rcur];
                    row = hp.row0[rcur];
                }
            }
        }
        plan_fast(it + 1);
    };

    // step `it`: count block it-1, regenerate block it, barrier, bookkeeping for block it-1
    regen(bt, bt + MT_N);
    __syncthreads();
    for (unsigned int it = 1; it < steps; it += 2) {
        bool done = count(it - 1);
        regen(bt + MT_N, bt);
        __syncthreads();
        after(done, it - 1);
        if (it + 1 < steps) {
            done = count(it);
            regen(bt, bt + MT_N);
            __syncthreads();
            after(done, it);
        }
    }
    {
        const bool done = count(steps - 1);
        __syncthreads();
        after(done, steps - 1);
    }
    // the replicate left open at the end of the chunk continues in the next chunk
    if (have && seg_lo < d_end) flush(hist + hcur * NP, row, true);
}

template <int PATH, typename IdxT>
static void mt_draw_launch(const uint32_t *d_states, unsigned long long chunk_len, const DrawPlan &plan, uint32_t n, IdxT *idx,
                           cudaStream_t st) {
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int nchunks = plan.cum[plan.nintervals];
    const int force = G().opt_draw_nch;
    const DrawConsts kc = {1u << 31, 1u << 21, 1u << 14, 0x9908b0dfu};
    // measured: 3 chunks per CTA beats 2 and 1 whenever there are enough chunks to give every SM a CTA
    if (force == 3 || (force < 0 && nchunks >= 3 * sms))
        mt_draw_kernel<PATH, IdxT, 3><<<(nchunks + 2) / 3, DrawCfg<3>::threads, 0, st>>>(d_states, chunk_len, plan, n, idx, kc);
    else if (force == 2)
        mt_draw_kernel<PATH, IdxT, 2><<<(nchunks + 1) / 2, DrawCfg<2>::threads, 0, st>>>(d_states, chunk_len, plan, n, idx, kc);
    else
        mt_draw_kernel<PATH, IdxT, 1><<<nchunks, DrawCfg<1>::threads, 0, st>>>(d_states, chunk_len, plan, n, idx, kc);
}

// plan for `nranges` disjoint, ascending stream ranges; off[r] = index-buffer offset of range r
int make_draw_plan(int nranges, const unsigned long long *lo, const unsigned long long *hi,
                   const unsigned long long *off, unsigned long long chunk_len, unsigned long long stream_end, DrawPlan *plan) {
    if (nranges < 1 || nranges > 32) {
        set_error(CGPU_EINVAL, "draw plan with %d ranges (1..32)", nranges);
        return CGPU_EINVAL;
    }
    plan->nranges = 0;
    plan->nintervals = 0;
    plan->cum[0] = 0;
    plan->stream_end = stream_end;
    for (int r = 0; r < nranges; r++) {
        if (hi[r] <= lo[r]) continue;
        const int k = plan->nranges++;
        plan->lo[k] = lo[r];
        plan->hi[k] = hi[r];
        plan->off[k] = off[r];
        unsigned long long c0 = lo[r] / chunk_len, c1 = (hi[r] - 1) / chunk_len;
        if (plan->nintervals > 0) {
            const int p = plan->nintervals - 1;
            const unsigned long long last = plan->first[p] + (unsigned long long)(plan->cum[p + 1] - plan->cum[p]) - 1;
            if (c0 <= last) c0 = last + 1;  // the boundary chunk is already scheduled (it stores into every range it touches)
            if (c0 > c1) continue;
            if (c0 == last + 1) {  // extend the previous interval
                plan->cum[p + 1] += (int)(c1 - c0 + 1);
                continue;
            }
        }
        const int p = plan->nintervals++;
        plan->first[p] = c0;
        plan->cum[p + 1] = plan->cum[p] + (int)(c1 - c0 + 1);
    }
    return CGPU_OK;
}

int resolve_draw_slots(const MtPlan &mt, DrawPlan *plan) {
    for (int i = 0; i < plan->nintervals; i++) {
        const int cnt = plan->cum[i + 1] - plan->cum[i];
        const long long a = mt.slot_of(plan->first[i]), b = mt.slot_of(plan->first[i] + (unsigned long long)(cnt - 1));
        if (a < 0 || b != a + cnt - 1) {
            set_error(CGPU_EINVAL, "draw plan asks for chunks [%llu, +%d) the jump tree did not prepare", plan->first[i], cnt);
            return CGPU_EINVAL;
        }
        plan->slot[i] = (int)a;
    }
    return CGPU_OK;
}

int launch_mt_draw(const MtPlan &mt, const DrawPlan &plan, uint32_t n, void *d_idx, cudaStream_t st) {
    if (plan.nranges <= 0 || plan.cum[plan.nintervals] <= 0) return CGPU_OK;
    const uint32_t *d_states = mt.d_states;
    const unsigned long long chunk_len = mt.chunk_len;
    uint32_t *i32 = static_cast<uint32_t *>(d_idx);
    uint16_t *i16 = static_cast<uint16_t *>(d_idx);
    if (n == 0) mt_draw_launch<2, uint32_t>(d_states, chunk_len, plan, n, i32, st);
    else if (idx_is16(n)) mt_draw_launch<0, uint16_t>(d_states, chunk_len, plan, n, i16, st);
    else if (n < (1u << 21)) mt_draw_launch<0, uint32_t>(d_states, chunk_len, plan, n, i32, st);
    else mt_draw_launch<1, uint32_t>(d_states, chunk_len, plan, n, i32, st);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// one contiguous range [lo, hi) stored at idx[0 ..)
int launch_mt_draw_range(const MtPlan &mt, unsigned long long lo, unsigned long long hi, uint32_t n, void *d_idx,
                         cudaStream_t st) {
    if (hi <= lo) return CGPU_OK;
    DrawPlan plan;
    const unsigned long long off = 0;
    CGPU_TRY(make_draw_plan(1, &lo, &hi, &off, mt.chunk_len, hi, &plan));
    CGPU_TRY(resolve_draw_slots(mt, &plan));
    return launch_mt_draw(mt, plan, n, d_idx, st);
}

// fused draws + histograms: counts rows of NP bytes, one zero-initialised matrix per range
size_t mt_draw_hist_smem(int NP, int cpc) { return ((size_t)(cpc + 1) * 2 * MT_N + (size_t)cpc * 2 * NP) * sizeof(uint32_t); }

bool mt_draw_hist_fits(int NP) { return mt_draw_hist_smem(NP, 3) <= (size_t)110 * 1024; }

int launch_mt_draw_hist(const MtPlan &mt, const DrawPlan &plan, const unsigned int *bn, unsigned char *const *row0, int NP,
                        uint32_t n, int *d_err, cudaStream_t st) {
    const int nchunks = plan.cum[plan.nintervals];
    if (plan.nranges <= 0 || nchunks <= 0) return CGPU_OK;
    HistPlan hp;
    for (int r = 0; r < plan.nranges; r++) {
        if (bn[r] < 2u * MT_N) {
            set_error(CGPU_EINVAL, "fused draw+histogram needs at least %d draws per replicate (got %u)", 2 * MT_N, bn[r]);
            return CGPU_EINVAL;
        }
        hp.bn[r] = bn[r];
        hp.row0[r] = row0[r];
    }
    hp.NP = NP;
    hp.err = d_err;
    const DrawConsts kc = {1u << 31, 1u << 21, 1u << 14, 0x9908b0dfu};
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    if (nchunks >= 3 * sms) {
        const size_t smem = mt_draw_hist_smem(NP, 3);
        static bool attr3 = false;
        if (!attr3) {
            CGPU_CUDA(cudaFuncSetAttribute(mt_draw_hist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            attr3 = true;
        }
        mt_draw_hist_kernel<3><<<(nchunks + 2) / 3, DrawCfg<3>::threads, smem, st>>>(mt.d_states, mt.chunk_len, plan, hp, n, kc);
    } else {
        const size_t smem = mt_draw_hist_smem(NP, 1);
        static bool attr1 = false;
        if (!attr1) {
            CGPU_CUDA(cudaFuncSetAttribute(mt_draw_hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            attr1 = true;
        }
        mt_draw_hist_kernel<1><<<nchunks, DrawCfg<1>::threads, smem, st>>>(mt.d_states, mt.chunk_len, plan, hp, n, kc);
    }
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

}  // namespace cgpu
