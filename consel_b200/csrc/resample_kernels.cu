// Site histograms: c[b, j] = #{draws of replicate b that hit site j}
// (src/makermt.c:53-62), written in the layout the contraction kernels read.
//
// Count layout ("BJ8"): one 64-bit word per (replicate group g = b/8, site j) holding
// the eight byte-counts of replicates 8g..8g+7 (byte t = replicate 8g+t); rows are
// NP = pad_sites(n) words long and zero padded, so a warp of the contraction kernel
// fetches the counts of its 8 replicates at a site with ONE broadcast 8-byte load.
//
// Two ways of producing it:
//   hist_group_kernel  histograms in shared memory.  Default: one CTA per replicate group, one
//                      WARP per replicate with a private byte histogram, fire-and-forget shared
//                      atomics on the packed bytes (hist_algo 1; 0 = match.any de-duplication and
//                      2 = tag arbitration are kept for measurements), row checksum against
//                      overflow, then either the interleaved BJ8 words (fp64 contraction) or plain
//                      byte rows (tensor-core contraction) are written with coalesced stores.
//                      Byte rows of long alignments (up to 225280 sites) let 2, 4 or 8 warps share
//                      one histogram.
//   hist_global_kernel everything else (BJ8 above 28160 sites, rows above 225280): one thread per
//                      draw, 32-bit atomic add on the global word that holds the byte.
// Draw sources: the index buffer written by mt_draw_kernel (reference stream) or a
// counter-based Philox4x32-10 keyed by (seed; scale, replicate) — Salmon, Moraes, Dror,
// Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11.
#include <stdlib.h>

#include <algorithm>

#include "cgpu_internal.h"
#include "philox.cuh"
#include "hist_bucket.cuh"

namespace cgpu {

constexpr int HIST_SMEM_SITES = 28160;  // 8 * 28160 B = 220 KB of shared memory
constexpr int HIST_THREADS = 256;

// one warp adds the site indices its lanes hold to a warp-private byte histogram.
//  HALGO 0  match.any + popcount: the first lane of each group of equal sites adds the
//           group size (no atomics, no extra memory)
//  HALGO 1  shared-memory atomic add on the 32-bit word that holds the byte
//  HALGO 2  claim-and-verify: every pending lane writes its lane id into a tag byte of its
//           site, re-reads it, and the lane that owns the tag increments; lanes that lost
//           a duplicate retry (a site drawn m times in one 32-draw step takes m rounds).
//           Plain byte loads/stores only; needs a second NP-byte array per warp.
template <int HALGO>
__device__ __forceinline__ void warp_hist_add(unsigned char *hist, unsigned char *tag, uint32_t k, bool valid,
                                              int lane, int *overflow) {
    if (HALGO == 0) {
        unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned peers = __match_any_sync(active, k);
            if (lane == __ffs(peers) - 1) {
                unsigned v = hist[k] + __popc(peers);
                if (v > 255u) *overflow = 1;
                hist[k] = (unsigned char)v;
            }
        }
        __syncwarp();
    } else if (HALGO == 1) {
        // fire-and-forget add on the packed byte; a byte that wraps past 255 is caught by the
        // row checksum after the counting phase (a wrap changes the byte sum by -255)
        if (valid) atomicAdd(reinterpret_cast<unsigned *>(hist) + (k >> 2), 1u << ((k & 3u) * 8u));
    } else {
        bool pending = valid;
        while (__any_sync(0xffffffffu, pending)) {
            if (pending) tag[k] = (unsigned char)lane;
            __syncwarp();
            if (pending && tag[k] == (unsigned char)lane) {
                unsigned v = hist[k] + 1u;
                if (v > 255u) *overflow = 1;
                hist[k] = (unsigned char)v;
                pending = false;
            }
            __syncwarp();
        }
    }
}

// SRC 0: uint32 indices from the buffer (replicate r of this launch at idx[r*bn ..));
// SRC 2: the same with uint16 indices;  SRC 1: Philox
// RPC replicates per CTA (8 warps): 8 = one warp per replicate; 4 / 2 / 1 (byte rows only) let
// 2 / 4 / 8 warps share one histogram, so alignments of up to 225280 sites still count in
// shared memory instead of with global atomics.
template <int SRC, int HALGO, bool ROWS, int RPC>
__global__ void __launch_bounds__(HIST_THREADS)
hist_group_kernel(const void *__restrict__ idx_v, int n, int NP, int bn, int nb,
                  unsigned long long *__restrict__ counts, int *__restrict__ err,
                  uint32_t key0, uint32_t key1, uint32_t scale, uint32_t b_first, uint32_t ctr3,
                  const unsigned int *__restrict__ pid) {
    extern __shared__ unsigned char hsm[];  // RPC histograms of NP bytes (+ 8 tag arrays for HALGO 2)
    static_assert(RPC == 8 || (ROWS && HALGO == 1), "shared histograms need the atomic update and the row layout");
    constexpr int WPR = 8 / RPC;  // warps per replicate
    unsigned char *tag = hsm + (size_t)8 * NP + (size_t)(threadIdx.x >> 5) * NP;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int slot = wid / WPR, sub = wid % WPR;
    const int ngroups = (nb + RPC - 1) / RPC;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        for (int i = tid; i < NP * RPC / 4; i += HIST_THREADS) reinterpret_cast<uint32_t *>(hsm)[i] = 0;
        __syncthreads();
        const int r = g * RPC + slot;  // replicate within this launch
        if (r < nb) {
            unsigned char *hist = hsm + (size_t)slot * NP;
            int overflow = 0;
            if (SRC != 1) {
                const uint32_t *src32 = static_cast<const uint32_t *>(idx_v) + (size_t)r * bn;
                const uint16_t *src16 = static_cast<const uint16_t *>(idx_v) + (size_t)r * bn;
                if (SRC == 2 && HALGO == 1) {
                    // uint16 indices, atomic byte histogram: each warp owns a contiguous segment of the
                    // replicate's draws and reads it with 16-byte loads (8 indices per lane), all updates
                    // unconditional; a scalar head (up to the 16-byte boundary) and tail frame the body
                    unsigned *h32 = reinterpret_cast<unsigned *>(hist);
                    auto add = [&](uint32_t k) {
                        if (pid) k = __ldg(pid + k);  // site -> pattern slot
                        atomicAdd(h32 + (k >> 2), 1u << ((k & 3u) * 8u));
                    };
                    const int seg = (bn + WPR - 1) / WPR;
                    int i = sub * seg;
                    const int hi = min(bn, i + seg);
                    const int head = min(hi - i, (int)((16 - (reinterpret_cast<uintptr_t>(src16 + i) & 15)) & 15) / 2);
                    if (lane < head) add(src16[i + lane]);
                    i += max(head, 0);
                    if (i + 256 <= hi) {
                        uint4 v = __ldg(reinterpret_cast<const uint4 *>(src16 + i + lane * 8));
                        for (; i + 256 <= hi; i += 256) {
                            uint4 nx = v;  // the next 8 indices are in flight while these are counted
                            if (i + 512 <= hi) nx = __ldg(reinterpret_cast<const uint4 *>(src16 + i + 256 + lane * 8));
                            add(v.x & 0xffffu); add(v.x >> 16);
                            add(v.y & 0xffffu); add(v.y >> 16);
                            add(v.z & 0xffffu); add(v.z >> 16);
                            add(v.w & 0xffffu); add(v.w >> 16);
                            v = nx;
                        }
                    }
                    for (int j = i + lane; j < hi; j += 32) add(src16[j]);
                } else
                // 8 independent index loads in flight per lane before the histogram updates
                // (the loop is otherwise bound by global-load latency)
                for (int i0 = sub * 256; i0 < bn; i0 += WPR * 256) {
                    uint32_t kk[8];
                    bool vv[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = i0 + u * 32 + lane;
                        vv[u] = i < bn;
                        kk[u] = vv[u] ? (SRC == 2 ? (uint32_t)src16[i] : src32[i]) : 0u;
                        if (pid && vv[u]) kk[u] = __ldg(pid + kk[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (i0 + u * 32 < bn) warp_hist_add<HALGO>(hist, tag, kk[u], vv[u], lane, &overflow);
                    }
                }
            } else {
                // draw i of replicate b is word (i & 3) of Philox(counter = (i >> 2, b, scale, stream id))
                const uint32_t b = b_first + (uint32_t)r;
                for (int i0 = sub * 128; i0 < bn; i0 += WPR * 128) {
                    int q = (i0 >> 2) + lane;
                    Philox4 p = philox4x32_10((uint32_t)q, b, scale, ctr3, key0, key1);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        bool valid = q * 4 + u < bn;
                        uint32_t k = __umulhi(p.v[u], (uint32_t)n);
                        if (pid) k = __ldg(pid + k);
                        warp_hist_add<HALGO>(hist, tag, k, valid, lane, &overflow);
                    }
                }
            }
            if (overflow) atomicExch(err, 1);
        }
        __syncthreads();
        if (HALGO == 1 && sub == 0 && r < nb) {
            // checksum: the bytes of a histogram must add up to the number of draws
            const uint4 *h4 = reinterpret_cast<const uint4 *>(hsm + (size_t)slot * NP);
            unsigned int sum = 0;
            for (int i = lane; i < NP / 16; i += 32) {
                const uint4 v = h4[i];
                sum = __dp4a(v.x, 0x01010101u, sum);
                sum = __dp4a(v.y, 0x01010101u, sum);
                sum = __dp4a(v.z, 0x01010101u, sum);
                sum = __dp4a(v.w, 0x01010101u, sum);
            }
#pragma unroll
            for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0 && sum != (unsigned int)bn) atomicExch(err, 1);
        }
        if (ROWS) {
            // plain rows: replicate RPC*g+slot owns NP consecutive bytes (rows past nb are padding)
            unsigned char *dst = reinterpret_cast<unsigned char *>(counts) + (size_t)g * RPC * NP;
            const uint4 *src4 = reinterpret_cast<const uint4 *>(hsm);
            uint4 *dst4 = reinterpret_cast<uint4 *>(dst);
            for (int i = tid; i < RPC * NP / 16; i += HIST_THREADS) dst4[i] = src4[i];
            __syncthreads();
            continue;
        }
        // interleave: 4 consecutive sites of all 8 histograms -> 4 BJ8 words (32 bytes)
        unsigned long long *row = counts + (size_t)g * NP;
        for (int j4 = tid; j4 < NP / 4; j4 += HIST_THREADS) {
            uint32_t h[8];
#pragma unroll
            for (int w = 0; w < 8; w++) h[w] = reinterpret_cast<const uint32_t *>(hsm + (size_t)w * NP)[j4];
            uint32_t out[8];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                // site j4*4+s: byte s of every h[w]
                const uint32_t sel = (uint32_t)s | ((4u + s) << 4);  // low two bytes = {a.byte[s], b.byte[s]}
                uint32_t lo01 = __byte_perm(h[0], h[1], sel) & 0xffffu, lo23 = __byte_perm(h[2], h[3], sel);
                uint32_t hi01 = __byte_perm(h[4], h[5], sel) & 0xffffu, hi23 = __byte_perm(h[6], h[7], sel);
                out[2 * s] = lo01 | (lo23 << 16);
                out[2 * s + 1] = hi01 | (hi23 << 16);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(row + (size_t)j4 * 4);
            dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
            dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
        }
        __syncthreads();
    }
}

template <int SRC, bool ROWS>
__global__ void __launch_bounds__(256)
hist_global_kernel(const void *__restrict__ idx_v, int n, int NP, int bn, int nb,
                   unsigned int *__restrict__ counts32, uint32_t key0, uint32_t key1, uint32_t scale,
                   uint32_t b_first, uint32_t ctr3, const unsigned int *__restrict__ pid) {
    const unsigned long long total = SRC != 1 ? (unsigned long long)nb * bn
                                              : (unsigned long long)nb * ((bn + 3) / 4);
    for (unsigned long long d = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; d < total;
         d += (unsigned long long)gridDim.x * blockDim.x) {
        if (SRC != 1) {
            uint32_t r = (uint32_t)(d / bn);
            uint32_t k = SRC == 2 ? (uint32_t)static_cast<const uint16_t *>(idx_v)[d] : static_cast<const uint32_t *>(idx_v)[d];
            if (pid) k = __ldg(pid + k);
            if (ROWS) {
                atomicAdd(&counts32[((size_t)r * NP + k) >> 2], 1u << ((k & 3) * 8));
            } else {
                size_t word = ((size_t)(r >> 3) * NP + k) * 2 + ((r & 7) >> 2);
                atomicAdd(&counts32[word], 1u << ((r & 3) * 8));
            }
        } else {
            uint32_t qn = (uint32_t)((bn + 3) / 4);
            uint32_t r = (uint32_t)(d / qn), q = (uint32_t)(d % qn);
            Philox4 p = philox4x32_10(q, b_first + r, scale, ctr3, key0, key1);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if ((int)(q * 4 + u) < bn) {
                    uint32_t k = __umulhi(p.v[u], (uint32_t)n);
                    if (pid) k = __ldg(pid + k);
                    if (ROWS) {
                        atomicAdd(&counts32[((size_t)r * NP + k) >> 2], 1u << ((k & 3) * 8));
                    } else {
                        size_t word = ((size_t)(r >> 3) * NP + k) * 2 + ((r & 7) >> 2);
                        atomicAdd(&counts32[word], 1u << ((r & 3) * 8));
                    }
                }
            }
        }
    }
}

// Long alignments, byte rows: ONE replicate per CTA and TH/32 warps on its single shared
// histogram.  The shared-memory footprint is the histogram alone (NP bytes), so an SM holds
// floor(220 KB / NP) CTAs and the warp count per SM stays at 32-64 whatever the site count is
// (n = 100000: two CTAs of 32 warps; the 8-warp CTAs of hist_group_kernel leave 8 warps per SM
// there, and the kernel is bound by the latency of the index loads and the shared atomics).
// Every warp owns a contiguous segment of the replicate's draws and reads it with 16-byte loads,
// the next one in flight while the current indices are counted.
template <int SRC, int TH>
__global__ void __launch_bounds__(TH)
hist_wide_kernel(const void *__restrict__ idx_v, int n, int NP, int bn, int nb, unsigned char *__restrict__ rows,
                 int *__restrict__ err, uint32_t key0, uint32_t key1, uint32_t scale, uint32_t b_first, uint32_t ctr3,
                 const unsigned int *__restrict__ pid) {
    extern __shared__ __align__(16) unsigned char hsm[];
    __shared__ unsigned int s_sum[TH / 32];
    constexpr int NW = TH / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned *h32 = reinterpret_cast<unsigned *>(hsm);
    auto add = [&](uint32_t k) {
        if (pid) k = __ldg(pid + k);  // site -> pattern slot
        atomicAdd(h32 + (k >> 2), 1u << ((k & 3u) * 8u));
    };
    for (int r = blockIdx.x; r < nb; r += gridDim.x) {
        for (int i = tid; i < NP / 16; i += TH) reinterpret_cast<uint4 *>(hsm)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (SRC == 1) {
            const uint32_t b = b_first + (uint32_t)r;
            const int nquad = (bn + 3) >> 2;
            for (int q = tid; q < nquad; q += TH) {
                Philox4 p = philox4x32_10((uint32_t)q, b, scale, ctr3, key0, key1);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (q * 4 + u < bn) add(__umulhi(p.v[u], (uint32_t)n));
            }
        } else {
            constexpr int PER = SRC == 2 ? 8 : 4;  // indices per 16-byte load
            const uint16_t *src16 = static_cast<const uint16_t *>(idx_v) + (size_t)r * bn;
            const uint32_t *src32 = static_cast<const uint32_t *>(idx_v) + (size_t)r * bn;
            auto at = [&](int i) -> uint32_t { return SRC == 2 ? (uint32_t)src16[i] : src32[i]; };
            const uintptr_t base = SRC == 2 ? reinterpret_cast<uintptr_t>(src16) : reinterpret_cast<uintptr_t>(src32);
            const int seg = (bn + NW - 1) / NW;
            int i = wid * seg;
            const int hi = min(bn, i + seg);
            if (i < hi) {
                const int esz = SRC == 2 ? 2 : 4;
                const int head = min(hi - i, (int)(((16 - ((base + (size_t)i * esz) & 15)) & 15) / esz));
                if (lane < head) add(at(i + lane));
                i += head;
                constexpr int STEP = 32 * PER;
                if (i + STEP <= hi) {
                    const uint4 *p4 = reinterpret_cast<const uint4 *>(base + (size_t)i * esz);
                    uint4 v = __ldg(p4 + lane);
                    for (; i + STEP <= hi; i += STEP, p4 += 32) {
                        uint4 nx = v;
                        if (i + 2 * STEP <= hi) nx = __ldg(p4 + 32 + lane);
                        if (SRC == 2) {
                            add(v.x & 0xffffu); add(v.x >> 16);
                            add(v.y & 0xffffu); add(v.y >> 16);
                            add(v.z & 0xffffu); add(v.z >> 16);
                            add(v.w & 0xffffu); add(v.w >> 16);
                        } else {
                            add(v.x); add(v.y); add(v.z); add(v.w);
                        }
                        v = nx;
                    }
                }
                for (int j = i + lane; j < hi; j += 32) add(at(j));
            }
        }
        __syncthreads();
        // checksum (a byte that wrapped past 255 changes the sum) while the row is copied out
        unsigned int sum = 0;
        uint4 *dst4 = reinterpret_cast<uint4 *>(rows + (size_t)r * NP);
        for (int i = tid; i < NP / 16; i += TH) {
            const uint4 v = reinterpret_cast<const uint4 *>(hsm)[i];
            sum = __dp4a(v.x, 0x01010101u, sum);
            sum = __dp4a(v.y, 0x01010101u, sum);
            sum = __dp4a(v.z, 0x01010101u, sum);
            sum = __dp4a(v.w, 0x01010101u, sum);
            dst4[i] = v;
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) s_sum[wid] = sum;
        __syncthreads();
        if (tid == 0) {
            unsigned int tot = 0;
            for (int w = 0; w < NW; w++) tot += s_sum[w];
            if (tot != (unsigned int)bn) atomicExch(err, 1);
        }
        // (the next iteration's zeroing is ordered behind the copy by the barrier above; s_sum is
        // rewritten only after the next two barriers)
    }
}

template <int SRC, int TH>
static int hist_wide_launch(const void *d_idx, int n, int NP, int bn, int nb, unsigned long long *d_counts, int *d_err,
                            uint32_t k0, uint32_t k1, uint32_t scale, uint32_t b_first, uint32_t ctr3, cudaStream_t st, const unsigned int *pid) {
    auto kern = hist_wide_kernel<SRC, TH>;
    static bool attr_done = false;
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * HIST_SMEM_SITES));
        attr_done = true;
    }
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const int per_sm = std::max(1, std::min(2048 / TH, (int)((size_t)(224 * 1024) / ((size_t)NP + 1024))));
    const int grid = std::min(nb, sms * per_sm);
    kern<<<grid, TH, (size_t)NP, st>>>(d_idx, n, NP, bn, nb, reinterpret_cast<unsigned char *>(d_counts), d_err, k0, k1, scale,
                                       b_first, ctr3, pid);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

// Bucketed histogram of genome-length byte rows (hist_bucket.cuh): four launches per call, the
// partition buffer (2 bytes per draw) and the bucket tables come from the stream-ordered pool.
template <int SRC>
static int hist_bucket_launch(const void *d_idx, int n, int NP, int bn, int nb, unsigned long long *d_counts, int *d_err,
                              uint32_t k0, uint32_t k1, uint32_t scale, uint32_t b_first, uint32_t ctr3, cudaStream_t st) {
    constexpr int WBL = 16;
    const int nbk = (NP + (1 << WBL) - 1) >> WBL;
    const int tiles = (bn + BK_TILE - 1) / BK_TILE;
    const size_t pstride = (size_t)((bn + 7) / 8 * 8) + (size_t)8 * nbk;
    auto k_count = bucket_count_kernel<SRC, WBL>;
    auto k_scatter = bucket_scatter_kernel<SRC, WBL>;
    auto k_hist = bucket_hist_kernel<WBL>;
    static bool attr_done = false;
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_scatter_smem()));
        CGPU_CUDA(cudaFuncSetAttribute(k_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 << WBL));
        attr_done = true;
    }
    unsigned int *d_tab = nullptr;      // cnt | off | cur, nb x nbk each
    unsigned short *d_part = nullptr;   // nb x pstride
    const size_t tab = (size_t)nb * nbk;
    CGPU_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_tab), 3 * tab * sizeof(unsigned int), st));
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&d_part), (size_t)nb * pstride * sizeof(unsigned short), st);
    if (e != cudaSuccess) {
        cudaFreeAsync(d_tab, st);
        set_error(CGPU_ENOMEM, "bucketed histogram: %zu bytes of partition buffer: %s", (size_t)nb * pstride * 2, cudaGetErrorString(e));
        return CGPU_ENOMEM;
    }
    unsigned int *d_cnt = d_tab, *d_off = d_tab + tab, *d_cur = d_tab + 2 * tab;
    int rc = CGPU_OK;
    auto body = [&]() -> int {
        CGPU_CUDA(cudaMemsetAsync(d_cnt, 0, tab * sizeof(unsigned int), st));
        const uint32_t *idx = static_cast<const uint32_t *>(d_idx);
        k_count<<<dim3(tiles, nb), BK_THREADS, 0, st>>>(idx, (uint32_t)n, bn, nbk, d_cnt, k0, k1, scale, b_first, ctr3);
        CGPU_LAUNCH_CHECK();
        bucket_scan_kernel<<<(nb + 127) / 128, 128, 0, st>>>(d_cnt, nb, nbk, d_off, d_cur);
        CGPU_LAUNCH_CHECK();
        k_scatter<<<dim3(tiles, nb), BK_THREADS, bk_scatter_smem(), st>>>(idx, (uint32_t)n, bn, nbk, d_off, d_cur, d_part, pstride, k0, k1,
                                                                          scale, b_first, ctr3);
        CGPU_LAUNCH_CHECK();
        k_hist<<<dim3(nbk, nb), BKH_THREADS, 1 << WBL, st>>>(d_part, pstride, d_cnt, d_off, nbk, NP,
                                                             reinterpret_cast<unsigned char *>(d_counts), d_err);
        CGPU_LAUNCH_CHECK();
        return CGPU_OK;
    };
    rc = body();
    cudaFreeAsync(d_part, st);
    cudaFreeAsync(d_tab, st);
    return rc;
}

// CGPU_HIST_ALGO selects the in-warp conflict scheme (development A/B switch)
static int hist_algo() {
    static int v = [] {
        const char *e = getenv("CGPU_HIST_ALGO");
        return e ? atoi(e) : 1;
    }();
    return G().opt_hist_algo >= 0 ? G().opt_hist_algo : v;
}

template <int SRC, int HALGO, bool ROWS, int RPC = 8>
static int hist_group_launch(const void *d_idx, int n, int NP, int bn, int nb, unsigned long long *d_counts,
                             int *d_err, uint32_t k0, uint32_t k1, uint32_t scale, uint32_t b_first, uint32_t ctr3, cudaStream_t st, const unsigned int *pid) {
    auto kern = hist_group_kernel<SRC, HALGO, ROWS, RPC>;
    static bool attr_done = false;
    if (!attr_done) {
        CGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * HIST_SMEM_SITES));
        attr_done = true;
    }
    const int ngroups = (nb + RPC - 1) / RPC;
    const size_t smem = (size_t)RPC * NP * (HALGO == 2 ? 2 : 1);
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    int per_sm = (int)std::min<size_t>(8, (220 * 1024) / smem);
    if (per_sm < 1) per_sm = 1;
    const int grid = std::min(ngroups, sms * per_sm);
    kern<<<grid, HIST_THREADS, smem, st>>>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, pid);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

static int hist_dispatch(int src, const void *d_idx, int n, int NP, int bn, int nb,
                         unsigned long long *d_counts, int *d_err, uint32_t k0, uint32_t k1, uint32_t scale,
                         uint32_t b_first, uint32_t ctr3, cudaStream_t st, bool rows, const unsigned int *pid) {
    if (nb <= 0) return CGPU_OK;
    if (pid && !rows) {
        set_error(CGPU_EINVAL, "pattern slots are a byte-row (int8 path) feature");
        return CGPU_EINVAL;
    }
    const int ngroups = (nb + 7) / 8;
    if (G().opt_hist_bucket == 2 && rows && !pid && src != 2 && nb <= 65535 && (NP + 65535) / 65536 <= BK_MAXB)
        return src == 1 ? hist_bucket_launch<1>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st)
                        : hist_bucket_launch<0>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st);
    int algo = hist_algo();
    if (algo == 2 && 2 * NP > HIST_SMEM_SITES) algo = 0;  // tag arrays double the footprint
    if (rows && algo == 1 && NP > HIST_SMEM_SITES / 2 && NP <= 8 * HIST_SMEM_SITES && G().opt_hist_wide != 0) {
        // 16 warps per replicate while four or more histograms fit an SM, 32 warps above that
        const bool th512 = (size_t)4 * ((size_t)NP + 1024) <= (size_t)224 * 1024;
#define CGPU_HISTWIDE_CASE(S)                                                                                              \
    if (src == S)                                                                                                          \
        return th512 ? hist_wide_launch<S, 512>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid)    \
                     : hist_wide_launch<S, 1024>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid);
        CGPU_HISTWIDE_CASE(0) CGPU_HISTWIDE_CASE(1) CGPU_HISTWIDE_CASE(2)
#undef CGPU_HISTWIDE_CASE
    }
    if (NP <= HIST_SMEM_SITES) {
#define CGPU_HIST_CASE(S, A)                                                                                     \
    if (src == S && algo == A)                                                                                    \
        return rows ? hist_group_launch<S, A, true>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid) \
                    : hist_group_launch<S, A, false>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid);
        CGPU_HIST_CASE(0, 0) CGPU_HIST_CASE(0, 1) CGPU_HIST_CASE(0, 2)
        CGPU_HIST_CASE(1, 0) CGPU_HIST_CASE(1, 1) CGPU_HIST_CASE(1, 2)
        CGPU_HIST_CASE(2, 0) CGPU_HIST_CASE(2, 1) CGPU_HIST_CASE(2, 2)
#undef CGPU_HIST_CASE
        set_error(CGPU_EINVAL, "CGPU_HIST_ALGO=%d", algo);
        return CGPU_EINVAL;
    } else if (rows && NP <= 8 * HIST_SMEM_SITES) {
        // byte rows of a long alignment: 4, 2 or 1 replicates per CTA share the 8 warps
#define CGPU_HISTW_CASE(S)                                                                                               \
    if (src == S) {                                                                                                       \
        if (NP <= 2 * HIST_SMEM_SITES) return hist_group_launch<S, 1, true, 4>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid); \
        if (NP <= 4 * HIST_SMEM_SITES) return hist_group_launch<S, 1, true, 2>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid); \
        return hist_group_launch<S, 1, true, 1>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st, pid);      \
    }
        CGPU_HISTW_CASE(0) CGPU_HISTW_CASE(1) CGPU_HISTW_CASE(2)
#undef CGPU_HISTW_CASE
        set_error(CGPU_EINVAL, "histogram source %d", src);
        return CGPU_EINVAL;
    } else {
        // byte rows longer than any shared-memory histogram.  "hist_bucket" 1 (opt-in): partition the draws by
        // 65536-site bucket and count every bucket in shared memory (hist_bucket.cuh); 2 takes that path for
        // every uint32-index row (tests).  Default: 32-bit global atomics on the packed words.
        if (rows && !pid && src != 2 && G().opt_hist_bucket >= 1 && nb <= 65535 && (NP + 65535) / 65536 <= BK_MAXB)
            return src == 1 ? hist_bucket_launch<1>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st)
                            : hist_bucket_launch<0>(d_idx, n, NP, bn, nb, d_counts, d_err, k0, k1, scale, b_first, ctr3, st);
        CGPU_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)ngroups * NP * 8, st));
        int sms = G().sm_count > 0 ? G().sm_count : 148;
        int grid = sms * 16;
        unsigned int *c32 = reinterpret_cast<unsigned int *>(d_counts);
#define CGPU_HISTG_CASE(S, R)                                                                                 \
    if (src == S && rows == R)                                                                                 \
        hist_global_kernel<S, R><<<grid, 256, 0, st>>>(d_idx, n, NP, bn, nb, c32, k0, k1, scale, b_first, ctr3, pid);
        CGPU_HISTG_CASE(0, false) CGPU_HISTG_CASE(0, true) CGPU_HISTG_CASE(1, false) CGPU_HISTG_CASE(1, true)
        CGPU_HISTG_CASE(2, false) CGPU_HISTG_CASE(2, true)
#undef CGPU_HISTG_CASE
        CGPU_LAUNCH_CHECK();
        if (rows && d_err) {  // a byte that wrapped past 255 changes the row's byte sum (hist_bucket.cuh)
            row_checksum_kernel<<<nb, RCS_THREADS, 0, st>>>(reinterpret_cast<const unsigned char *>(d_counts), NP, bn, d_err);
            CGPU_LAUNCH_CHECK();
        }
    }
    return CGPU_OK;
}

int launch_hist_from_idx(const void *d_idx, int n, int NP, int bn, int nb,
                         unsigned long long *d_counts, int *d_err, cudaStream_t st, bool rows, const unsigned int *pid) {
    return hist_dispatch(idx_is16((uint32_t)n) ? 2 : 0, d_idx, n, NP, bn, nb, d_counts, d_err, 0, 0, 0, 0, 0, st, rows, pid);
}

int launch_hist_philox(unsigned long long seed, int scale, int b_first, int n, int NP, int bn, int nb,
                       unsigned long long *d_counts, int *d_err, cudaStream_t st, bool rows, unsigned long long stream_offset,
                       const unsigned int *pid) {
    // The counter-based generator has no stream position; calls that continue one logical stream
    // (makermt -g: one call per gene file, stream_offset = draws consumed so far) must still draw
    // independent replicates, so the offset is hashed (splitmix64 finaliser) into the fourth
    // counter word.  Offset 0 keeps the word 0.
    unsigned long long z = stream_offset;
    if (z) {
        z += 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        z = (z & 0xffffffffull) ? z : 1;
    }
    return hist_dispatch(1, nullptr, n, NP, bn, nb, d_counts, d_err, (uint32_t)seed, (uint32_t)(seed >> 32),
                         (uint32_t)scale, (uint32_t)b_first, (uint32_t)z, st, rows, pid);
}

// BJ8 -> plain counts[b*n + j] bytes (test seam: cgpu_resample_counts)
__global__ void counts_unpack_kernel(const unsigned long long *__restrict__ counts, int n, int NP, int nb,
                                     unsigned char *__restrict__ out) {
    size_t total = (size_t)nb * n;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(e / n), j = (int)(e % n);
        unsigned long long w = counts[(size_t)(b >> 3) * NP + j];
        out[e] = (unsigned char)(w >> ((b & 7) * 8));
    }
}

int launch_counts_unpack(const unsigned long long *d_counts, int n, int NP, int nb, unsigned char *d_out,
                         cudaStream_t st) {
    if (nb <= 0) return CGPU_OK;
    counts_unpack_kernel<<<1024, 256, 0, st>>>(d_counts, n, NP, nb, d_out);
    CGPU_LAUNCH_CHECK();
    return CGPU_OK;
}

}  // namespace cgpu
