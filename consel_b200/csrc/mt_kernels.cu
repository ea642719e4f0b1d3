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
//                    output word j is the XOR of x_{i+j} over the set bits i of g.
//                    The bit range of g is sliced over gridDim.y CTAs (atomicXor
//                    combine) so early rounds of the doubling tree still fill SMs.
//   mt_draw_kernel   one CTA per chunk of 2^m draws: ping-pong block regeneration
//                    in shared memory (3 phases of <=227 independent lanes, exactly
//                    the dependency structure of next_state), tempering, site index,
//                    coalesced store of the indices.
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

// ------------------------------------------------------------------ jump
constexpr int JUMP_THREADS = 640;
constexpr int JUMP_SEQ = MT_DEG + MT_N + 63;  // words of the regenerated sequence kept in smem

__global__ void __launch_bounds__(JUMP_THREADS, 2)
mt_jump_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
               const uint32_t *__restrict__ poly, int words_per_slice) {
    extern __shared__ uint32_t sm[];
    uint32_t *x = sm;                 // JUMP_SEQ words
    uint32_t *g = sm + JUMP_SEQ + 1;  // 624 words of the polynomial
    const int tid = threadIdx.x;
    const uint32_t *w = src + (size_t)blockIdx.x * MT_N;
    uint32_t *o = dst + (size_t)blockIdx.x * MT_N;

    const int w_lo = blockIdx.y * words_per_slice;
    const int w_hi = min(w_lo + words_per_slice, MT_POLY_W32);
    if (w_lo >= w_hi) return;
    const int need = min(w_hi * 32, MT_DEG) + MT_N;  // sequence words x[0..need)

    if (tid < MT_N) {
        x[tid] = w[tid];
        g[tid] = poly[tid];
    }
    __syncthreads();
    // x[t+624] depends on x[t+397]: 227 new words per step are independent
    for (int t0 = 0; t0 + MT_N < need; t0 += MT_N - MT_M) {
        int t = t0 + tid;
        if (tid < MT_N - MT_M && t + MT_N < need)
            x[t + MT_N] = x[t + MT_M] ^ mt_twist(x[t], x[t + 1]);
        __syncthreads();
    }
    if (tid < MT_N) {
        uint32_t r = 0;
        const uint32_t *xj = x + tid;
        for (int wi = w_lo; wi < w_hi; wi++) {
            uint32_t gw = g[wi];
            if (wi == MT_POLY_W32 - 1) gw &= 0xffffffffu >> (MT_POLY_W32 * 32 - MT_DEG);  // bits >= 19937 are not part of a residue
            const uint32_t *xb = xj + wi * 32;
#pragma unroll
            for (int b = 0; b < 32; b++)
                if (gw >> b & 1) r ^= xb[b];
        }
        if (gridDim.y == 1) o[tid] = r;
        else if (r) atomicXor(&o[tid], r);
    }
}

int launch_mt_jump(const uint32_t *d_src, uint32_t *d_dst, const uint32_t *d_poly, int count,
                   cudaStream_t st) {
    if (count <= 0) return CGPU_OK;
    static bool attr_done = false;
    const size_t smem = (size_t)(JUMP_SEQ + 1 + MT_N) * sizeof(uint32_t);
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    // slice the polynomial over enough CTAs to cover the SMs when few windows jump
    int sms = G().sm_count > 0 ? G().sm_count : 148;
    int slices = 1;
    while (slices < 16 && count * slices < sms) slices *= 2;
    int wps = (MT_POLY_W32 + slices - 1) / slices;
    if (slices > 1) CGPU_CUDA(cudaMemsetAsync(d_dst, 0, (size_t)count * MT_N * sizeof(uint32_t), st));
    mt_jump_kernel<<<dim3(count, slices), JUMP_THREADS, smem, st>>>(d_src, d_dst, d_poly, wps);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// ------------------------------------------------------------------ draws
constexpr int DRAW_THREADS = 256;

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

// NCH chunks per CTA share the four block barriers of a regeneration step.
// The draws to keep are given as up to 32 disjoint stream ranges [lo, hi) (one per scale:
// the replicate columns this rank owns), each mapped to an offset of the compact index
// buffer; the chunks to run are the union of the chunk intervals covering those ranges.
template <int PATH, typename IdxT, int NCH>
__global__ void __launch_bounds__(DRAW_THREADS)
mt_draw_kernel(const uint32_t *__restrict__ states, int chunk_log2, DrawPlan plan, uint32_t n, IdxT *__restrict__ idx) {
    __shared__ uint32_t buf[NCH][2][MT_N + 1];
    const int tid = threadIdx.x;
    const double dn = (double)n;
    unsigned long long d_begin[NCH], d_end[NCH];
    int rcur[NCH];
    unsigned long long fast_until[NCH];  // steps below this take the branch-free store (or skip) path
    bool inside[NCH];
    IdxT *outp[NCH];
    unsigned long long steps = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int ci = blockIdx.x * NCH + c;
        d_begin[c] = d_end[c] = 0;
        rcur[c] = 0;
        fast_until[c] = 0;
        inside[c] = false;
        outp[c] = idx;
        if (ci < plan.cum[plan.nintervals]) {
            int iv = 0;
            while (ci >= plan.cum[iv + 1]) iv++;
            const unsigned long long chunk = plan.first[iv] + (unsigned long long)(ci - plan.cum[iv]);
            d_begin[c] = chunk << chunk_log2;
            d_end[c] = d_begin[c] + (1ull << chunk_log2);
            if (d_end[c] > plan.stream_end) d_end[c] = plan.stream_end;
            if (d_end[c] > d_begin[c]) {
                const uint32_t *w = states + chunk * MT_N;
                for (int i = tid; i < MT_N; i += DRAW_THREADS) buf[c][0][i] = w[i];
                const unsigned long long st = (d_end[c] - d_begin[c] + MT_N - 1) / MT_N;
                steps = st > steps ? st : steps;
            } else {
                d_end[c] = d_begin[c];
            }
        }
    }
    __syncthreads();
    int cur = 0;
    constexpr int H = MT_N - MT_M;  // 227 words become computable per dependent step
    for (unsigned long long it = 0; it < steps; it++) {
        // phase A: k in [0,227) reads only old words
        if (tid < H) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint32_t *o = buf[c][cur];
                buf[c][cur ^ 1][tid] = o[tid + MT_M] ^ mt_twist(o[tid], o[tid + 1]);
            }
        }
        __syncthreads();
        // phase B: k in [227,454) reads new[k-227]
        if (tid < H) {
            const int k = tid + H;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint32_t *o = buf[c][cur];
                uint32_t *nw = buf[c][cur ^ 1];
                nw[k] = nw[k - H] ^ mt_twist(o[k], o[k + 1]);
            }
        }
        __syncthreads();
        // phase C: k in [454,624); the last word wraps to new[0]
        if (tid < MT_N - 2 * H) {
            const int k = tid + 2 * H;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint32_t *o = buf[c][cur];
                uint32_t *nw = buf[c][cur ^ 1];
                const uint32_t nxt = (k == MT_N - 1) ? nw[0] : o[k + 1];
                nw[k] = nw[k - H] ^ mt_twist(o[k], nxt);
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint32_t *nw = buf[c][cur ^ 1];
            if (it < fast_until[c]) {
                // fast path: the whole block lies inside one range (store all 624) or outside all
                if (inside[c]) {
#pragma unroll
                    for (int i = tid; i < MT_N; i += DRAW_THREADS) outp[c][i] = (IdxT)site_index<PATH>(mt_temper(nw[i]), n, dn);
                    outp[c] += MT_N;
                }
                continue;
            }
            // slow path (a few times per chunk): block touching a range boundary or the chunk end
            const unsigned long long d0 = d_begin[c] + it * MT_N;
            if (d0 >= d_end[c]) {
                fast_until[c] = ~0ull;
                inside[c] = false;
                continue;
            }
            const unsigned long long d1 = d_end[c] - d0 >= (unsigned long long)MT_N ? d0 + MT_N : d_end[c];
            while (rcur[c] < plan.nranges && d0 >= plan.hi[rcur[c]]) rcur[c]++;
            for (int r = rcur[c]; r < plan.nranges && plan.lo[r] < d1; r++) {
                const unsigned long long lo = plan.lo[r], hi = plan.hi[r];
                const int i_lo = d0 >= lo ? 0 : (int)(lo - d0);
                const int i_hi = (int)((d1 < hi ? d1 : hi) - d0);
                IdxT *out = idx + plan.off[r] + (long long)(d0 - lo);  // only [i_lo, i_hi) is touched
#pragma unroll
                for (int i = tid; i < MT_N; i += DRAW_THREADS)
                    if (i >= i_lo && i < i_hi) out[i] = (IdxT)site_index<PATH>(mt_temper(nw[i]), n, dn);
            }
            // plan the following blocks: how many of them are entirely inside / outside
            const unsigned long long dn0 = d0 + MT_N;
            while (rcur[c] < plan.nranges && dn0 >= plan.hi[rcur[c]]) rcur[c]++;
            fast_until[c] = it + 1;
            inside[c] = false;
            if (rcur[c] < plan.nranges && dn0 < d_end[c]) {
                const int r = rcur[c];
                if (plan.lo[r] <= dn0) {
                    const unsigned long long lim = plan.hi[r] < d_end[c] ? plan.hi[r] : d_end[c];
                    fast_until[c] = it + 1 + (lim - dn0) / MT_N;
                    inside[c] = true;
                    outp[c] = idx + plan.off[r] + (dn0 - plan.lo[r]);
                } else {
                    const unsigned long long lim = plan.lo[r] < d_end[c] ? plan.lo[r] : d_end[c];
                    fast_until[c] = it + 1 + (lim - dn0) / MT_N;
                }
            } else if (rcur[c] >= plan.nranges) {
                fast_until[c] = ~0ull;  // nothing left to keep in this chunk
            }
        }
        cur ^= 1;
        // no barrier needed here: the next phase A only writes buf[.][cur^1] (= the old
        // array, last read in phase C before the barrier above) and reads buf[.][cur].
    }
}

template <int PATH, typename IdxT>
static void mt_draw_launch(const uint32_t *d_states, int chunk_log2, const DrawPlan &plan, uint32_t n, IdxT *idx,
                           cudaStream_t st) {
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int nchunks = plan.cum[plan.nintervals];
    const int force = G().opt_draw_nch;
    if (force == 4 || (force < 0 && nchunks >= 8 * sms))
        mt_draw_kernel<PATH, IdxT, 4><<<(nchunks + 3) / 4, DRAW_THREADS, 0, st>>>(d_states, chunk_log2, plan, n, idx);
    else if (force == 2 || (force < 0 && nchunks >= 4 * sms))
        mt_draw_kernel<PATH, IdxT, 2><<<(nchunks + 1) / 2, DRAW_THREADS, 0, st>>>(d_states, chunk_log2, plan, n, idx);
    else
        mt_draw_kernel<PATH, IdxT, 1><<<nchunks, DRAW_THREADS, 0, st>>>(d_states, chunk_log2, plan, n, idx);
}

// plan for `nranges` disjoint, ascending stream ranges; off[r] = index-buffer offset of range r
int make_draw_plan(int nranges, const unsigned long long *lo, const unsigned long long *hi,
                   const unsigned long long *off, int chunk_log2, unsigned long long stream_end, DrawPlan *plan) {
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
        unsigned long long c0 = lo[r] >> chunk_log2, c1 = (hi[r] - 1) >> chunk_log2;
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

int launch_mt_draw(const uint32_t *d_states, int chunk_log2, const DrawPlan &plan, uint32_t n, void *d_idx,
                   cudaStream_t st) {
    if (plan.nranges <= 0 || plan.cum[plan.nintervals] <= 0) return CGPU_OK;
    uint32_t *i32 = static_cast<uint32_t *>(d_idx);
    uint16_t *i16 = static_cast<uint16_t *>(d_idx);
    if (n == 0) mt_draw_launch<2, uint32_t>(d_states, chunk_log2, plan, n, i32, st);
    else if (idx_is16(n)) mt_draw_launch<0, uint16_t>(d_states, chunk_log2, plan, n, i16, st);
    else if (n < (1u << 21)) mt_draw_launch<0, uint32_t>(d_states, chunk_log2, plan, n, i32, st);
    else mt_draw_launch<1, uint32_t>(d_states, chunk_log2, plan, n, i32, st);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// one contiguous range [lo, hi) stored at idx[0 ..)
int launch_mt_draw_range(const uint32_t *d_states, int chunk_log2, unsigned long long lo, unsigned long long hi,
                         uint32_t n, void *d_idx, cudaStream_t st) {
    if (hi <= lo) return CGPU_OK;
    DrawPlan plan;
    const unsigned long long off = 0;
    CGPU_TRY(make_draw_plan(1, &lo, &hi, &off, chunk_log2, hi, &plan));
    return launch_mt_draw(d_states, chunk_log2, plan, n, d_idx, st);
}

}  // namespace cgpu
