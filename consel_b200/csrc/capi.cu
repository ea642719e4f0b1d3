// C ABI of libconselgpu.so (include/consel_gpu.h): lifecycle, the makermt-side
// pipeline (MT19937 jump-ahead -> draws -> site histograms -> contraction) and its
// batching over device memory.  The consel-side entry points live in capi_stats.cu.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <map>
#include <mutex>
#include <vector>

#include "cgpu_internal.h"

namespace cgpu {

// one process-wide message: the AU driver issues count queries from a helper thread, and the
// program that reports the failure (`#! Terminated: ...`) runs on the main thread
static char t_err[512] = "no error";
static std::mutex g_err_mu;
std::atomic<long long> g_launches{0};

void set_error(int code, const char *fmt, ...) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    va_list ap;
    va_start(ap, fmt);
    int off = snprintf(t_err, sizeof(t_err), "cgpu error %d: ", code);
    vsnprintf(t_err + off, sizeof(t_err) - off, fmt, ap);
    va_end(ap);
}

Globals &G() {
    static Globals g;
    return g;
}

int ensure_init() {
    if (!G().inited) return cgpu_init(-1);
    (void)cudaGetLastError();  // every entry point starts from a clean error state
    return CGPU_OK;
}

PhaseTimer::PhaseTimer(cudaStream_t s) : st(s) {
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaEventRecord(t0, st);
}
PhaseTimer::~PhaseTimer() {
    for (int i = 0; i < nspans; i++) {
        cudaEventDestroy(spans[i].a);
        cudaEventDestroy(spans[i].b);
    }
    free(spans);
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
}
void PhaseTimer::begin(int phase, cudaStream_t on) {
    cur = on ? on : st;
    if (nspans == cap) {
        cap = cap ? 2 * cap : 64;
        spans = (Span *)realloc(spans, sizeof(Span) * cap);
    }
    Span &sp = spans[nspans++];
    sp.phase = phase;
    sp.ended = false;
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, cur);
}
void PhaseTimer::end() {
    cudaEventRecord(spans[nspans - 1].b, cur);
    spans[nspans - 1].ended = true;
}
void PhaseTimer::finish(double *ms) {
    for (int i = 0; i < NPHASE; i++) ms[i] = 0;
    cudaEventRecord(t1, st);
    cudaEventSynchronize(t1);
    float tot = 0;
    if (cudaEventElapsedTime(&tot, t0, t1) == cudaSuccess) ms[5] = tot;
    static const bool trace = getenv("CGPU_TRACE") != nullptr;  // development: every span with its start time
    for (int i = 0; i < nspans; i++) {
        Span &sp = spans[i];
        float v = 0;
        if (trace && sp.ended) {
            float a = 0, b = 0;
            if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&a, t0, sp.a) == cudaSuccess &&
                cudaEventElapsedTime(&b, t0, sp.b) == cudaSuccess)
                fprintf(stderr, "[cgpu trace] phase %d: %8.3f -> %8.3f ms\n", sp.phase, a, b);
        }
        // a span left open by an early error return has no end event: skip it (querying it
        // would leave a stale "invalid resource handle" for the next launch check)
        if (sp.ended && cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&v, sp.a, sp.b) == cudaSuccess)
            ms[sp.phase] += v;
    }
    ms[1] += ms[8] + ms[9];
    // slots 10 / 11: how many tensor-core GEMM / histogram launches the call made (bench rooflines)
    ms[10] = ms[11] = 0;
    for (int i = 0; i < nspans; i++) {
        if (spans[i].phase == 6) ms[10] += 1;
        if (spans[i].phase == 2) ms[11] += 1;
    }
    (void)cudaGetLastError();
}

// ---- per-kernel timing registry (KT)
struct KSpan {
    std::string name;
    cudaEvent_t a, b;
    double work;
};
static std::vector<KSpan> &kspans() {
    static std::vector<KSpan> v;
    return v;
}
KT::KT(const char *name, cudaStream_t s, double work) : st(s) {
    if (G().opt_ktime != 1) return;
    KSpan sp;
    sp.name = name;
    sp.work = work;
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, st);
    kspans().push_back(sp);
    slot = (int)kspans().size() - 1;
}
KT::~KT() {
    if (slot >= 0 && slot < (int)kspans().size()) cudaEventRecord(kspans()[slot].b, st);
}

static void shard_range(int bb, int *b0, int *b1) {
    const Globals &g = G();
    *b0 = (int)((long long)bb * g.shard_rank / g.shard_world);
    *b1 = (int)((long long)bb * (g.shard_rank + 1) / g.shard_world);
}

// scratch arena (slots: 0 index stream, 1/2 count buffers, 3 int32 partial sums); stream-ordered on
// the library stream, valid until the next request for the same slot
static cudaError_t scratch_get(int slot, size_t bytes, void **out) {
    Globals &g = G();
    if (g.scratch_cap[slot] < bytes) {
        dev_free(g.scratch[slot]);
        g.scratch[slot] = nullptr;
        g.scratch_cap[slot] = 0;
        const size_t want = bytes + bytes / 8;
        cudaError_t e = dev_alloc(&g.scratch[slot], want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            e = dev_alloc(&g.scratch[slot], bytes);
            if (e != cudaSuccess) return e;
            g.scratch_cap[slot] = bytes;
        } else {
            g.scratch_cap[slot] = want;
        }
    }
    *out = g.scratch[slot];
    return cudaSuccess;
}

int lt_prepare(DevMatrix &m, cudaStream_t st);

static void free_reps(DevReps *r) {
    if (!r) return;
    if (r->d_rep)
        for (int s = 0; s < r->S; s++) dev_free(r->d_rep[s]);
    if (r->d_stat)
        for (int s = 0; s < r->S; s++) dev_free(r->d_stat[s]);
    dev_free(r->d_edges);
    dev_free(r->d_ass_ptr);
    dev_free(r->d_ass_idx);
    free(r->d_rep);
    free(r->d_stat);
    free(r->bb);
    free(r->b0);
    free(r->nb);
    delete r;
}

static int alloc_reps(int K, int S, const int *bb, DevReps **out) {
    DevReps *r = new DevReps();
    r->K = K;
    r->S = S;
    r->bb = (int *)calloc(S, sizeof(int));
    r->b0 = (int *)calloc(S, sizeof(int));
    r->nb = (int *)calloc(S, sizeof(int));
    r->d_rep = (double **)calloc(S, sizeof(double *));
    for (int s = 0; s < S; s++) {
        int b0, b1;
        shard_range(bb[s], &b0, &b1);
        r->bb[s] = bb[s];
        r->b0[s] = b0;
        r->nb[s] = b1 - b0;
        size_t bytes = (size_t)K * std::max(1, b1 - b0) * sizeof(double);
        cudaError_t e = dev_alloc(&r->d_rep[s], bytes);
        if (e != cudaSuccess) {
            set_error(CGPU_ENOMEM, "device allocation (%zu bytes) for replicate block %d: %s", bytes, s, cudaGetErrorString(e));
            free_reps(r);
            return CGPU_ENOMEM;
        }
    }
    *out = r;
    return CGPU_OK;
}

// Chunk length (draws) of the MT19937 jump tree, from a cost model measured on B200: a jump
// costs ~0.75 us of whole-GPU time (126 us on one SM), a tree level ~20 us of latency, and
// a draw CTA slot (6 chunks resident per SM) produces one draw per ~1.0 ns; the draw launch
// runs in whole waves of 6 x SMs chunks.  Candidates: the powers of two (their jump
// polynomials x^(2^i) are tabulated at init) and, per wave count, the EQUAL-FILL length that
// makes the chunks of this call just fill the slots (774 chunks of 2^22 draws leave 13 % of
// the slots of a C2 launch idle for its whole duration); its polynomials x^(len*2^r) are
// computed on first use and cached.
static unsigned long long pick_chunk_len(unsigned long long needed_draws, unsigned long long stream_end, int nranges,
                                         const unsigned long long *lo, const unsigned long long *hi) {
    if (G().opt_chunk_log2 >= 10 && G().opt_chunk_log2 <= 30) return 1ull << G().opt_chunk_log2;
    const int sms = G().sm_count > 0 ? G().sm_count : 148;
    const double slots = 6.0 * sms;
    unsigned long long best = 1ull << 12;
    double best_cost = 1e300;
    auto levels = [&](unsigned long long len) {
        int top = 0;
        while ((stream_end + len - 1) / len > (1ull << top)) top++;
        return top;
    };
    for (int m = 12; m <= 24; m++) {
        const double chunk = (double)(1ull << m);
        const double pn = std::ceil((double)needed_draws / chunk) + nranges;
        const int top = levels(1ull << m);
        if (m + top > MT_MAXLOG) continue;
        const double waves = std::ceil(pn / slots);
        const double cost = pn * 0.75e-6 + top * 20e-6 + waves * chunk * 1.0e-9;
        if (cost < best_cost) {
            best_cost = cost;
            best = 1ull << m;
        }
    }
    const bool force_fill = G().opt_chunk_fill == 2;  // tests: equal-fill whatever the model says
    // (needed_draws == 0: a rank whose shard of every scale is empty -- fewer replicates than ranks; the forced
    // equal-fill length would be zero)
    if (G().opt_chunk_fill == 0 || nranges > 32 || needed_draws == 0 || (!force_fill && needed_draws < (1ull << 22))) return best;
    constexpr unsigned long long ALIGN = 4 * MT_N;  // whole groups of four blocks (the branch-free body of the draw kernel)
    std::vector<unsigned long long> off(nranges, 0);
    unsigned long long pick = best;
    for (int w = 1; w <= 4; w++) {
        const long long cap = (long long)(w * slots);
        if (cap <= nranges + 1) continue;
        unsigned long long len = (needed_draws + (unsigned long long)(cap - nranges) - 1) / (unsigned long long)(cap - nranges);
        len = (len + ALIGN - 1) / ALIGN * ALIGN;
        int count = 0;
        for (int tries = 0; tries < 8; tries++) {  // the estimate ignores how the ranges sit on the chunk grid
            DrawPlan dp;
            if (make_draw_plan(nranges, lo, hi, off.data(), len, stream_end, &dp) != CGPU_OK) return best;
            count = dp.cum[dp.nintervals];
            if (count <= cap) break;
            len = (len + len / 128 + ALIGN - 1) / ALIGN * ALIGN;
        }
        if (count > cap || (len < 4096 && !force_fill)) continue;
        const int top = levels(len);
        if (top > 40 || (len >> (62 - top)) != 0) continue;  // len * 2^top must stay a 64-bit exponent
        const double cost = count * 0.75e-6 + top * 20e-6 + w * (double)len * 1.0e-9;
        if (cost < 0.98 * best_cost || (force_fill && pick == best)) {
            best_cost = cost;
            pick = len;
        }
    }
    return pick;
}

// Position lists of the jump polynomials x^(len * 2^r) mod phi, r = 0 .. levels-1, for a chunk
// length that is not a power of two: built on the host on first use (a few squarings each),
// kept on the device.  Entries are never freed while a plan may still reference them; the
// cache is bounded by rebuilding from scratch (after a device synchronize) when it is full.
struct JumpPolys {
    uint32_t *d_pos = nullptr;
    int levels = 0;
    int nquads[64];
};
static std::map<unsigned long long, JumpPolys> &jump_cache() {
    static std::map<unsigned long long, JumpPolys> c;  // (node-based: addresses of live entries are stable)
    return c;
}
static std::vector<uint32_t *> &jump_graveyard() {
    static std::vector<uint32_t *> g;
    return g;
}
static void jump_cache_clear() {
    for (auto &kv : jump_cache()) cudaFree(kv.second.d_pos);
    for (uint32_t *p : jump_graveyard()) cudaFree(p);
    jump_cache().clear();
    jump_graveyard().clear();
}
// called at the start of a resampling call, when no plan references the cache
static int jump_cache_trim() {
    if (jump_cache().size() < 20 && jump_graveyard().empty()) return CGPU_OK;
    CGPU_CUDA(cudaDeviceSynchronize());
    if (jump_cache().size() >= 20) jump_cache_clear();
    for (uint32_t *p : jump_graveyard()) cudaFree(p);
    jump_graveyard().clear();
    return CGPU_OK;
}
static int jump_polys_for(unsigned long long len, int levels, const uint32_t **d_pos, const int **nquads) {
    auto &cache = jump_cache();
    auto it = cache.find(len);
    if (it != cache.end() && it->second.levels >= levels) {
        *d_pos = it->second.d_pos;
        *nquads = it->second.nquads;
        return CGPU_OK;
    }
    if (it != cache.end()) {  // too few levels: an earlier plan of this call may still use the old lists
        jump_graveyard().push_back(it->second.d_pos);
        cache.erase(it);
    }
    JumpPolys jp;
    jp.levels = std::max(levels, 12);
    while (jp.levels > levels && (len >> (62 - jp.levels)) != 0) jp.levels--;
    std::vector<uint32_t> pos((size_t)jp.levels * MT_POS_STRIDE, 0u);
    for (int r = 0; r < jp.levels; r++) {
        Gf2Poly g;
        mt_poly_xpow(len << r, &g);
        jp.nquads[r] = mt_poly_positions(&g, pos.data() + (size_t)r * MT_POS_STRIDE);
    }
    CGPU_CUDA(cudaMalloc(&jp.d_pos, pos.size() * sizeof(uint32_t)));
    CGPU_CUDA(cudaMemcpy(jp.d_pos, pos.data(), pos.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    auto ins = cache.emplace(len, jp).first;
    *d_pos = ins->second.d_pos;
    *nquads = ins->second.nquads;
    return CGPU_OK;
}

static unsigned long long g_last_chunk_len = 0;

// Chunk start states W_{offset + c*2^m} for the chunks c that cover the kept ranges, by a
// binary jump tree walked from the top bit down: at bit r every live node a (a multiple of
// 2^(r+1)) hands its state to a + 2^r through the jump x^(2^(m+r)) if any wanted chunk lies
// in [a + 2^r, a + 2^(r+1)).  One launch per level; the jumps of a level are independent.
static int mt_plan(unsigned long seed, unsigned long long stream_offset, unsigned long long stream_end, int nranges,
                   const unsigned long long *lo, const unsigned long long *hi, MtPlan *plan) {
    unsigned long long needed = 0;
    for (int r = 0; r < nranges; r++) needed += hi[r] > lo[r] ? hi[r] - lo[r] : 0;
    plan->chunk_len = pick_chunk_len(needed, stream_end, nranges, lo, hi);
    g_last_chunk_len = plan->chunk_len;
    const unsigned long long clen = plan->chunk_len;
    const bool pow2 = (clen & (clen - 1)) == 0;
    int m = 0;  // log2 of a power-of-two chunk length
    while (pow2 && (1ull << m) < clen) m++;
    DrawPlan dp;
    {
        std::vector<unsigned long long> off(nranges, 0);
        CGPU_TRY(make_draw_plan(nranges, lo, hi, off.data(), clen, stream_end, &dp));
    }
    plan->nintervals = dp.nintervals;
    for (int i = 0; i < dp.nintervals; i++) {
        plan->first[i] = dp.first[i];
        plan->count[i] = dp.cum[i + 1] - dp.cum[i];
        plan->slot[i] = dp.cum[i];
    }
    int nslots = dp.cum[dp.nintervals];
    if (nslots == 0) return CGPU_OK;
    const unsigned long long maxchunk = plan->first[plan->nintervals - 1] + (unsigned long long)(plan->count[plan->nintervals - 1] - 1);
    int top = 0;
    while (top < 63 && (maxchunk >> top) != 0) top++;
    if (top > 0 && (pow2 ? m + top - 1 > MT_MAXLOG : (top > 40 || (clen >> (62 - top)) != 0))) {
        set_error(CGPU_EINVAL, "stream of %llu draws exceeds the jump table", stream_end);
        return CGPU_EINVAL;
    }
    if (pow2) {
        plan->d_pos = G().d_jump_table + (size_t)m * MT_POS_STRIDE;
        plan->nquads = G().jump_nquads + m;
    } else {
        CGPU_TRY(jump_polys_for(clen, top, &plan->d_pos, &plan->nquads));
    }
    auto wanted = [&](unsigned long long a, unsigned long long b) {  // any wanted chunk in [a, b]?
        for (int i = 0; i < plan->nintervals; i++)
            if (plan->first[i] <= b && plan->first[i] + (unsigned long long)(plan->count[i] - 1) >= a) return true;
        return false;
    };
    std::vector<std::pair<unsigned long long, int>> extra;  // inner nodes that are not wanted chunks themselves
    auto slot_get = [&](unsigned long long a) {
        long long s = plan->slot_of(a);
        if (s >= 0) return (int)s;
        for (auto &e : extra)
            if (e.first == a) return e.second;
        extra.emplace_back(a, nslots);
        return nslots++;
    };
    std::vector<std::vector<int2>> levels(top);
    std::vector<unsigned long long> cur{0}, next;
    const int root = slot_get(0);
    for (int r = top - 1; r >= 0; r--) {
        const unsigned long long step = 1ull << r;
        next.clear();
        for (unsigned long long a : cur) {
            if (wanted(a, a + step - 1)) next.push_back(a);
            if (wanted(a + step, a + 2 * step - 1)) {
                levels[r].push_back(make_int2(slot_get(a), slot_get(a + step)));
                next.push_back(a + step);
            }
        }
        cur.swap(next);
    }
    size_t npairs = 0;
    for (auto &l : levels) npairs += l.size();
    plan->h_stage.assign(MT_N + 2 * npairs, 0u);
    uint32_t *w0 = plan->h_stage.data();
    mt_seed_window((uint32_t)(seed & 0xffffffffUL), w0);
    if (stream_offset) {
        Gf2Poly gp;
        mt_poly_xpow(stream_offset, &gp);
        uint32_t w1[MT_N];
        mt_jump_host(w0, &gp, w1);
        memcpy(w0, w1, sizeof(w1));
    }
    {
        int2 *hp = reinterpret_cast<int2 *>(plan->h_stage.data() + MT_N);
        for (int r = top - 1; r >= 0; r--)
            for (auto &pr : levels[r]) *hp++ = pr;
    }
    plan->nslots = nslots;
    plan->root = root;
    plan->npairs = npairs;
    plan->level_cnt.assign(top, 0);
    for (int r = 0; r < top; r++) plan->level_cnt[r] = (int)levels[r].size();
    CGPU_CUDA(dev_alloc(&plan->d_states, (size_t)nslots * MT_N * sizeof(uint32_t)));
    if (npairs) CGPU_CUDA(dev_alloc(&plan->d_pairs, npairs * sizeof(int2)));
    return CGPU_OK;
}

// second half of the preparation: copies and the jump launches, on any stream that is ordered
// after the allocations above (made on the library stream)
static int mt_launch(MtPlan *plan, cudaStream_t st) {
    Globals &g = G();
    if (plan->nslots == 0) return CGPU_OK;
    CGPU_CUDA(cudaMemsetAsync(plan->d_states, 0, (size_t)plan->nslots * MT_N * sizeof(uint32_t), st));
    CGPU_CUDA(cudaMemcpyAsync(plan->d_states + (size_t)plan->root * MT_N, plan->h_stage.data(), MT_N * sizeof(uint32_t),
                              cudaMemcpyHostToDevice, st));
    if (plan->npairs)
        CGPU_CUDA(cudaMemcpyAsync(plan->d_pairs, plan->h_stage.data() + MT_N, plan->npairs * sizeof(int2), cudaMemcpyHostToDevice, st));
    size_t done = 0;
    for (int r = (int)plan->level_cnt.size() - 1; r >= 0; r--) {
        const int cnt = plan->level_cnt[r];
        if (!cnt) continue;
        CGPU_TRY(launch_mt_jump(plan->d_states, plan->d_pairs + done, plan->d_pos + (size_t)r * MT_POS_STRIDE,
                                plan->nquads[r], cnt, st));
        done += cnt;
    }
    return CGPU_OK;
}

static int mt_prepare(unsigned long seed, unsigned long long stream_offset, unsigned long long stream_end, int nranges,
                      const unsigned long long *lo, const unsigned long long *hi, MtPlan *plan, cudaStream_t st) {
    CGPU_TRY(mt_plan(seed, stream_offset, stream_end, nranges, lo, hi, plan));
    return mt_launch(plan, st);
}

struct Workspace {
    bool borrowed = false;  // index / count buffers belong to the scratch arena
    unsigned char *d_idx = nullptr;  // uint16 or uint32 site indices (idx_bytes(n))
    unsigned long long *d_counts[2] = {nullptr, nullptr};
    int *d_err = nullptr;
    cudaEvent_t hist_done[2] = {nullptr, nullptr}, contract_done[2] = {nullptr, nullptr}, idx_ready = nullptr;
    Workspace() {
        for (int i = 0; i < 2; i++) {
            cudaEventCreateWithFlags(&hist_done[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&contract_done[i], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&idx_ready, cudaEventDisableTiming);
    }
    ~Workspace() {
        if (!borrowed) {
            dev_free(d_idx);
            dev_free(d_counts[0]);
            dev_free(d_counts[1]);
        }
        dev_free(d_err);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(hist_done[i]);
            cudaEventDestroy(contract_done[i]);
        }
        cudaEventDestroy(idx_ready);
    }
};

// host_out (optional): rank-local compact blocks K x nb[s]; the copy of scale s is queued on
// the copy stream as soon as its last contraction is done, so the D2H of one scale hides
// behind the compute of the next.
static int scaleboot_core(DevMatrix &m, int S, const int *bn, const int *bb, unsigned long seed,
                          unsigned long long stream_offset, int rng_mode, int sum_mode, DevReps *reps,
                          PhaseTimer &pt, cudaStream_t st, double *const *host_out = nullptr) {
    const bool i8 = sum_mode == CGPU_SUM_INT8;
    if (sum_mode != CGPU_SUM_STRICT && sum_mode != CGPU_SUM_FAST && !i8) {
        set_error(CGPU_EINVAL, "unknown sum_mode %d", sum_mode);
        return CGPU_EINVAL;
    }
    if (rng_mode != CGPU_RNG_MT19937 && rng_mode != CGPU_RNG_PHILOX) {
        set_error(CGPU_EINVAL, "unknown rng_mode %d", rng_mode);
        return CGPU_EINVAL;
    }
    CGPU_TRY(jump_cache_trim());  // no jump plan is alive here
    for (int s = 0; s < S; s++)
        if (bn[s] <= 0 || bb[s] < 0) {
            set_error(CGPU_EINVAL, "scale %d: bn=%d bb=%d", s, bn[s], bb[s]);
            return CGPU_EINVAL;
        }
    const int n = m.n;
    int NP = m.NP;  // length of a count row (the int8 path may shorten it to the pattern slots, below)
    // batch size: bound the index and count buffers (device memory is 180 GB; keep the
    // workspace near 8 GB so several processes can share a GPU in tests)
    const size_t budget = (size_t)4 << 30;
    int max_bn = 1;
    for (int s = 0; s < S; s++) max_bn = std::max(max_bn, bn[s]);
    // (the index term only matters when the indices are produced batch by batch: an MT19937 call
    // whose whole index stream does not fit the workspace, see idx_whole_run below)
    unsigned long long total_draws = 0;
    for (int s = 0; s < S; s++) total_draws += (unsigned long long)reps->nb[s] * (unsigned long long)bn[s];
    const bool idx_per_batch = rng_mode == CGPU_RNG_MT19937 &&
                               (S > 32 || total_draws * idx_bytes((uint32_t)n) > ((unsigned long long)64 << 30));
    long long cap = (long long)(budget / (size_t)NP);
    if (idx_per_batch) cap = std::min<long long>(cap, (long long)(budget / ((size_t)max_bn * 4)));
    cap = std::max<long long>(64, cap / 64 * 64);
    int max_nb = 0;
    for (int s = 0; s < S; s++) max_nb = std::max(max_nb, reps->nb[s]);
    int *d_partial = nullptr;  // int32 partial sums of the tensor-core path

    if (i8) {
        if ((long long)max_bn * 128 >= (1ll << 31)) {
            set_error(CGPU_ERANGE, "CGPU_SUM_INT8: %d draws per replicate can overflow the int32 accumulators", max_bn);
            return CGPU_ERANGE;
        }
        cap = std::max<long long>(256, cap / 256 * 256);
    } else {
        CGPU_TRY(lt_prepare(m, st));
    }
    const int align = i8 ? 256 : 64;
    if (G().opt_batch > 0) cap = std::max<long long>(align, (long long)G().opt_batch / align * align);  // tests: force several batches per scale
    const int batch = (int)std::min<long long>(cap, std::max(align, (max_nb + align - 1) / align * align));
    // The fixed-point slices of L (and the partial-sum buffer of an unfused recombination) are prepared
    // right before the FIRST GEMM when the matrix is still on its way to the device (a large host-buffer
    // call): the jump tree, the draws and the first histogram do not read L, so they run under the
    // upload.  Otherwise (resident matrix, or a small one) they are prepared up front, and identical
    // columns of L are looked for: with pattern slots the count rows and the GEMM get shorter.
    bool fx_ready = !i8;
    const bool fx_early = i8 && (!host_out || (size_t)m.K * (size_t)n <= ((size_t)8 << 20));
    const bool allow_patterns = fx_early && (double)max_bn <= 4.0 * (double)n && G().opt_draw_fuse != 1 && G().opt_i8_patterns != 0;
    auto ensure_fx = [&]() -> int {
        if (fx_ready) return CGPU_OK;
        pt.begin(0);
        CGPU_TRY(fx_prepare(m, st, allow_patterns));
        pt.end();
        if (fx_needs_partial(m, max_bn, batch))
            CGPU_CUDA(scratch_get(3, m.fxMpad * (size_t)batch * sizeof(int), reinterpret_cast<void **>(&d_partial)));
        fx_ready = true;
        return CGPU_OK;
    };
    if (fx_early) {
        CGPU_TRY(ensure_fx());
        NP = m.NPc;
    } else if (i8 && m.d_A8 && m.npat > 0) {
        CGPU_TRY(ensure_fx());  // (re-slices uncompressed: this call cannot use the slots)
    }
    const unsigned int *pid = (i8 && m.npat > 0) ? m.d_pid : nullptr;

    Workspace ws;
    ws.borrowed = true;
    CGPU_CUDA(dev_alloc(&ws.d_err, sizeof(int)));
    CGPU_CUDA(cudaMemsetAsync(ws.d_err, 0, sizeof(int), st));

    MtPlan plan;
    struct PlanGuard {
        MtPlan &p;
        ~PlanGuard() {
            dev_free(p.d_states);
            dev_free(p.d_pairs);
        }
    } guard{plan};
    // staged index production (see below): one jump tree per stage
    constexpr int MAX_STAGES = 4;
    MtPlan plans[MAX_STAGES];
    PlanGuard sguard0{plans[0]}, sguard1{plans[1]}, sguard2{plans[2]}, sguard3{plans[3]};
    int nstages = 1, stage_cut[MAX_STAGES + 1] = {0, 0, 0, 0, 0};
    cudaEvent_t stage_tree[MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    struct TreeEvents {
        cudaEvent_t *e;
        ~TreeEvents() {
            cudaStreamSynchronize(G().stream2);  // an error path may leave the aux stream working on the plans
            for (int i = 0; i < MAX_STAGES; i++)
                if (e[i]) cudaEventDestroy(e[i]);
        }
    } tguard{stage_tree};
    std::vector<unsigned long long> st_lo, st_hi;  // kept draw ranges per scale (filled below)
    std::vector<unsigned long long> base(S + 1, 0);
    for (int s = 0; s < S; s++) base[s + 1] = base[s] + (unsigned long long)bb[s] * (unsigned long long)bn[s];
    // When this rank's share of the index stream fits the workspace, ONE draw launch covers
    // all of its scales (fills the machine; a per-scale launch has only ~100-200 chunk
    // CTAs): the kept draws are S disjoint stream ranges packed into a compact buffer.
    // Otherwise indices are produced batch by batch.
    bool idx_whole_run = false;
    std::vector<unsigned long long> idx_off(S + 1, 0);
    // int8 mode with a site count that fits shared memory: one fused draw+histogram launch
    // writes the count rows of every scale (no index stream at all)
    bool fuse_hist = false;
    unsigned char *d_cnt_all = nullptr;
    struct CntGuard {
        unsigned char *&p;
        ~CntGuard() { dev_free(p); }
    } cguard{d_cnt_all};
    std::vector<size_t> cnt_off(S + 1, 0);
    // (opt-in, "draw_fuse" 1: measured 8.7 ms against 4.8 + 3.1 ms for the two kernels at C2 — a
    // replicate ends every 5.5 blocks and the per-lane boundary bookkeeping plus the row flush
    // in front of a CTA-wide barrier cost more than the index round trip through HBM saves)
    if (rng_mode == CGPU_RNG_MT19937 && i8 && S <= 32 && G().opt_draw_fuse == 1 && mt_draw_hist_fits(NP)) {
        int min_bn = 1 << 30;
        for (int s = 0; s < S; s++) {
            min_bn = std::min(min_bn, bn[s]);
            cnt_off[s + 1] = cnt_off[s] + (size_t)((reps->nb[s] + 255) / 256 * 256) * NP;
        }
        fuse_hist = min_bn >= 2 * MT_N && cnt_off[S] > 0 && cnt_off[S] <= ((size_t)24 << 30);
        for (int s = 0; s < S; s++) fuse_hist = fuse_hist && reps->nb[s] <= batch;  // one GEMM batch per scale
        if (fuse_hist) {
            CGPU_CUDA(dev_alloc(&d_cnt_all, cnt_off[S]));
            CGPU_CUDA(cudaMemsetAsync(d_cnt_all, 0, cnt_off[S], st));
        }
    }
    if (!fuse_hist)
        for (int i = 0; i < 2; i++) {
            CGPU_CUDA(scratch_get(1 + i, (size_t)(batch / 8) * NP * 8, reinterpret_cast<void **>(&ws.d_counts[i])));
            // fp64 paths: padding sites and replicates must read as zero counts.  The byte rows of
            // the tensor-core path need no clearing: the histogram kernels write whole rows
            // (padding sites included) and the GEMM never stores the columns of padding rows.
            if (!i8) CGPU_CUDA(cudaMemsetAsync(ws.d_counts[i], 0, (size_t)(batch / 8) * NP * 8, st));
        }
    if (rng_mode == CGPU_RNG_MT19937) {
        for (int s = 0; s < S; s++) idx_off[s + 1] = idx_off[s] + (unsigned long long)reps->nb[s] * (unsigned long long)bn[s];
        // index stream of the whole call in one buffer when it fits 64 GB and half of the free
        // device memory (C3: 38 GB of uint32 indices); otherwise batch by batch
        unsigned long long idx_budget = (unsigned long long)64 << 30;
        // (the query takes a driver lock -- a concurrent NVML poller stretches it to milliseconds -- so it is
        // only made when the index buffer would have to GROW past 16 GB; a buffer that already exists fits)
        if (idx_off[S] * idx_bytes((uint32_t)n) > ((unsigned long long)16 << 30) &&
            idx_off[S] * idx_bytes((uint32_t)n) > G().scratch_cap[0]) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                idx_budget = std::min<unsigned long long>(idx_budget, (free_b + G().scratch_cap[0]) / 2);
        }
        idx_whole_run = !fuse_hist && S <= 32 && idx_off[S] > 0 && idx_off[S] * idx_bytes((uint32_t)n) <= idx_budget;
        const size_t idx_elems = fuse_hist ? 1 : idx_whole_run ? (size_t)idx_off[S] : (size_t)batch * max_bn;
        CGPU_CUDA(scratch_get(0, idx_elems * idx_bytes((uint32_t)n), reinterpret_cast<void **>(&ws.d_idx)));
        std::vector<unsigned long long> lo(S), hi(S);
        for (int s = 0; s < S; s++) {
            lo[s] = base[s] + (unsigned long long)reps->b0[s] * (unsigned long long)bn[s];
            hi[s] = lo[s] + (unsigned long long)reps->nb[s] * (unsigned long long)bn[s];
        }
        // Host-output calls in int8 mode are bound by the device-to-host copy of the replicate
        // blocks (PCIe, 15 ms at C2 against 10 ms of compute), which can only start when the
        // first scale is finished.  The index stream is therefore produced in up to four
        // stages of growing size (about 4 %, 16 %, 45 % and 100 % of the draws, each with its
        // own jump tree and chunk size), each stage's draw launch queued right before its
        // first scale: the copy engine starts after ~1 ms and stays busy, at the price of
        // ~1 ms of extra jumps.  Everything stays on the main stream (the tensor-core GEMM
        // needs whole SMs, so overlapping long draw CTAs with it only delays it).
        if (idx_whole_run && i8 && host_out && S >= 4 && S <= 32 && G().opt_stage != 0 &&
            (G().opt_stage == 1 || idx_off[S] >= (1ull << 28))) {
            double frac[MAX_STAGES - 1] = {0.04, 0.16, 0.45};
            if (const char *e = getenv("CGPU_STAGE_FRACS")) sscanf(e, "%lf,%lf,%lf", &frac[0], &frac[1], &frac[2]);  // measurement
            nstages = 0;
            stage_cut[0] = 0;
            for (int f = 0; f < MAX_STAGES - 1; f++) {
                int c = stage_cut[nstages] + 1;
                while (c < S && (double)idx_off[c] < frac[f] * (double)idx_off[S]) c++;
                if (c >= S) break;
                stage_cut[++nstages] = c;
            }
            stage_cut[++nstages] = S;
            for (int g = 0; g < nstages; g++)
                CGPU_TRY(mt_plan(seed, stream_offset, base[S], stage_cut[g + 1] - stage_cut[g], lo.data() + stage_cut[g],
                                 hi.data() + stage_cut[g], &plans[g]));
            st_lo = lo;
            st_hi = hi;
            for (int g = 1; g < nstages; g++) CGPU_CUDA(cudaEventCreateWithFlags(&stage_tree[g], cudaEventDisableTiming));
            CGPU_CUDA(cudaEventRecord(ws.idx_ready, st));  // the plans' allocations are ordered on the main stream
            CGPU_CUDA(cudaStreamWaitEvent(G().stream2, ws.idx_ready, 0));
        } else {
        pt.begin(8);
        if (S > 32) {  // more scales than plan ranges: prepare the hull of this rank's draws
            lo[0] = *std::min_element(lo.begin(), lo.end());
            hi[0] = *std::max_element(hi.begin(), hi.end());
        }
        CGPU_TRY(mt_prepare(seed, stream_offset, base[S], S > 32 ? 1 : S, lo.data(), hi.data(), &plan, st));
        pt.end();
        pt.begin(9);
        if (fuse_hist) {
            // draws and histograms in one kernel: the site indices never reach HBM
            std::vector<unsigned long long> flo, fhi, foff;
            std::vector<unsigned int> fbn;
            std::vector<unsigned char *> frow;
            for (int s = 0; s < S; s++) {
                if (hi[s] <= lo[s]) continue;
                flo.push_back(lo[s]);
                fhi.push_back(hi[s]);
                foff.push_back(0);
                fbn.push_back((unsigned int)bn[s]);
                frow.push_back(d_cnt_all + cnt_off[s]);
            }
            if (!flo.empty()) {
                DrawPlan dp;
                CGPU_TRY(make_draw_plan((int)flo.size(), flo.data(), fhi.data(), foff.data(), plan.chunk_len, base[S], &dp));
                CGPU_TRY(resolve_draw_slots(plan, &dp));
                CGPU_TRY(launch_mt_draw_hist(plan, dp, fbn.data(), frow.data(), NP, (uint32_t)n, ws.d_err, st));
            }
        } else if (idx_whole_run) {
            DrawPlan dp;
            CGPU_TRY(make_draw_plan(S, lo.data(), hi.data(), idx_off.data(), plan.chunk_len, base[S], &dp));
            CGPU_TRY(resolve_draw_slots(plan, &dp));
            CGPU_TRY(launch_mt_draw(plan, dp, (uint32_t)n, ws.d_idx, st));
        }
        pt.end();
        }
    }

    // Two-stream software pipeline over (scale, batch) items: the histogram of item i+1
    // (shared-memory bound) runs on the aux stream while the contraction of item i (FP64
    // bound) runs on the main stream; the two count buffers alternate.  With per-batch
    // index generation (sharded or very long streams) everything stays on one stream.
    // (the tensor-core GEMM keeps ~200 KB of shared memory per SM, which starves a
    // concurrent histogram kernel, so the int8 mode runs its stages back to back)
    static const bool no_overlap = getenv("CGPU_NO_OVERLAP") != nullptr;
    const int want = G().opt_overlap >= 0 ? G().opt_overlap : (i8 ? 0 : 1);
    const bool overlap = !no_overlap && want && (rng_mode == CGPU_RNG_PHILOX || idx_whole_run);
    cudaStream_t aux = overlap ? G().stream2 : st;
    if (overlap) {
        CGPU_CUDA(cudaEventRecord(ws.idx_ready, st));  // memsets, jump tree and index stream are queued on st
        CGPU_CUDA(cudaStreamWaitEvent(aux, ws.idx_ready, 0));
    }
    int item = 0;
    for (int s = 0; s < S; s++) {
        const double r = (double)bn[s] / (double)n;  // src/makermt.c:48
        if (nstages > 1)
            for (int g = 0; g < nstages; g++)
                if (s == stage_cut[g]) {  // jump tree and draws of the stage that begins with this scale
                    if (g == 0) {
                        pt.begin(8);
                        CGPU_TRY(mt_launch(&plans[0], st));
                        pt.end();
                    } else {
                        CGPU_CUDA(cudaStreamWaitEvent(st, stage_tree[g], 0));  // built on the aux stream meanwhile
                    }
                    pt.begin(9);
                    DrawPlan dp;
                    CGPU_TRY(make_draw_plan(stage_cut[g + 1] - stage_cut[g], st_lo.data() + stage_cut[g], st_hi.data() + stage_cut[g],
                                            idx_off.data() + stage_cut[g], plans[g].chunk_len, base[S], &dp));
                    CGPU_TRY(resolve_draw_slots(plans[g], &dp));
                    CGPU_TRY(launch_mt_draw(plans[g], dp, (uint32_t)n, ws.d_idx, st));
                    pt.end();
                    if (g + 1 < nstages) {
                        // the next stage's jump tree (latency bound while it is narrow) runs on the aux
                        // stream under this stage's draws, histograms and GEMMs
                        cudaStream_t aux2 = G().stream2;
                        pt.begin(8, aux2);
                        CGPU_TRY(mt_launch(&plans[g + 1], aux2));
                        pt.end();
                        CGPU_CUDA(cudaEventRecord(stage_tree[g + 1], aux2));
                    }
                }
        for (int a = 0; a < reps->nb[s]; a += batch, item++) {
            const int nbat = std::min(batch, reps->nb[s] - a);
            const int bA = reps->b0[s] + a;  // global replicate index of the batch
            const int buf = item & 1;
            if (overlap && item >= 2) CGPU_CUDA(cudaStreamWaitEvent(aux, ws.contract_done[buf], 0));
            if (fuse_hist) {
                // counts of this scale were written by the fused launch
            } else if (rng_mode == CGPU_RNG_MT19937) {
                const unsigned long long lo = base[s] + (unsigned long long)bA * bn[s];
                const unsigned long long hi = lo + (unsigned long long)nbat * bn[s];
                const unsigned char *idx_src = ws.d_idx;
                if (idx_whole_run) {
                    idx_src = ws.d_idx + (idx_off[s] + (unsigned long long)a * bn[s]) * idx_bytes((uint32_t)n);
                } else {
                    pt.begin(9);
                    CGPU_TRY(launch_mt_draw_range(plan, lo, hi, (uint32_t)n, ws.d_idx, st));
                    pt.end();
                }
                pt.begin(2, aux);
                CGPU_TRY(launch_hist_from_idx(idx_src, n, NP, bn[s], nbat, ws.d_counts[buf], ws.d_err, aux, i8, pid));
                pt.end();
            } else {
                pt.begin(2, aux);
                CGPU_TRY(launch_hist_philox((unsigned long long)seed, s, bA, n, NP, bn[s], nbat, ws.d_counts[buf], ws.d_err, aux, i8, stream_offset, pid));
                pt.end();
            }
            if (overlap) {
                CGPU_CUDA(cudaEventRecord(ws.hist_done[buf], aux));
                CGPU_CUDA(cudaStreamWaitEvent(st, ws.hist_done[buf], 0));
            }
            CGPU_TRY(ensure_fx());
            pt.begin(3);
            if (i8)
                CGPU_TRY(launch_contract_i8(m, fuse_hist ? d_cnt_all + cnt_off[s] : reinterpret_cast<const unsigned char *>(ws.d_counts[buf]), nbat,
                                            (nbat + 255) / 256 * 256, d_partial, r, reps->d_rep[s], reps->nb[s], a, st, &pt, bn[s]));
            else
                CGPU_TRY(launch_contract(m, ws.d_counts[buf], nbat, r, reps->d_rep[s], reps->nb[s], a, sum_mode, st));
            pt.end();
            if (overlap) CGPU_CUDA(cudaEventRecord(ws.contract_done[buf], st));
        }
        if (host_out && reps->nb[s] > 0) {
            cudaStream_t cp = G().stream3;
            cudaEvent_t done;
            CGPU_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
            CGPU_CUDA(cudaEventRecord(done, st));
            CGPU_CUDA(cudaStreamWaitEvent(cp, done, 0));
            cudaEventDestroy(done);
            pt.begin(4, cp);
            CGPU_CUDA(cudaMemcpyAsync(host_out[s], reps->d_rep[s], (size_t)reps->K * reps->nb[s] * sizeof(double),
                                      cudaMemcpyDeviceToHost, cp));
            pt.end();
        }
    }
    if (host_out) {
        // the main stream (and with it every later free) must not run ahead of the copies
        CGPU_CUDA(cudaEventRecord(ws.idx_ready, G().stream3));
        CGPU_CUDA(cudaStreamWaitEvent(st, ws.idx_ready, 0));
    }
    int h_err = 0;
    CGPU_CUDA(cudaMemcpyAsync(&h_err, ws.d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        set_error(CGPU_ERANGE, "a site was drawn more than 255 times in one replicate (n=%d too small for int8 counts)", n);
        return CGPU_ERANGE;
    }
    return CGPU_OK;
}

// d_L <- L on stream `st`, no host synchronisation (the caller keeps L alive until the stream has
// passed m->ready).  With a communicator and `collective` set every rank copies only its share of
// the rows from the host and the ranks all-gather over NVLink: N ranks of one box then pull K*n*8
// bytes through the host's memory system once, not N times.  The fp64 tile layout (LT) and the
// fixed-point slices are derived lazily by the modes that need them (lt_prepare / fx_prepare).
static int matrix_upload(const double *L, int K, int n, DevMatrix **out, cudaStream_t st, bool collective = false) {
    if (!L || K <= 0 || n <= 0) {
        set_error(CGPU_EINVAL, "matrix_upload: K=%d n=%d", K, n);
        return CGPU_EINVAL;
    }
    DevMatrix *m = new DevMatrix();
    m->K = K;
    m->n = n;
    m->NP = pad_sites(n);
    m->TK = pick_tk(K);
    m->KT = 32 * m->TK;
    m->ktiles = (K + m->KT - 1) / m->KT;
    const int world = collective && G().opt_dist_upload != 0 ? comm_world() : 1, rank = comm_rank();
    const int rows_per = (K + world - 1) / world;
    int rc = CGPU_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        set_error(CGPU_ECUDA, "matrix_upload %s: %s", what, cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? CGPU_ENOMEM : CGPU_ECUDA;
    };
    cudaError_t e;
    if ((e = dev_alloc(&m->d_L, (size_t)rows_per * world * n * sizeof(double))) != cudaSuccess) fail(e, "alloc L");
    if (!rc && (e = cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
    if (!rc && st != G().stream) {  // the allocation is ordered on the library stream
        cudaEvent_t ev;
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) fail(e, "event");
        if (!rc) {
            cudaEventRecord(ev, G().stream);
            cudaStreamWaitEvent(st, ev, 0);
            cudaEventDestroy(ev);
        }
    }
    if (!rc) {
        if (world == 1) {
            if ((e = cudaMemcpyAsync(m->d_L, L, (size_t)K * n * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) fail(e, "H2D");
        } else {
            const int r0 = std::min(K, rank * rows_per), r1 = std::min(K, r0 + rows_per);
            double *mine = m->d_L + (size_t)rank * rows_per * n;
            if (r1 > r0 && (e = cudaMemcpyAsync(mine, L + (size_t)r0 * n, (size_t)(r1 - r0) * n * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess)
                fail(e, "H2D");
            if (!rc) rc = comm_allgather_f64(mine, m->d_L, (size_t)rows_per * n, st);  // in place: rank p's rows land at p*rows_per
        }
    }
    if (!rc && (e = cudaEventRecord(m->ready, st)) != cudaSuccess) fail(e, "record");
    if (rc) {
        dev_free(m->d_L);
        if (m->ready) cudaEventDestroy(m->ready);
        delete m;
        return rc;
    }
    *out = m;
    return CGPU_OK;
}

// the fp64 contraction's tile layout of L and the "every entry is finite" flag, on first use
int lt_prepare(DevMatrix &m, cudaStream_t st) {
    if (m.d_LT) return CGPU_OK;
    if (m.ready) CGPU_CUDA(cudaStreamWaitEvent(st, m.ready, 0));
    int *d_flag = nullptr;
    CGPU_CUDA(dev_alloc(&m.d_LT, (size_t)m.ktiles * m.NP * m.KT * sizeof(double)));
    CGPU_CUDA(dev_alloc(&d_flag, sizeof(int)));
    CGPU_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    CGPU_TRY(launch_transpose_L(m.d_L, m.K, m.n, m.NP, m.TK, m.d_LT, d_flag, st));
    int h_flag = 0;
    CGPU_CUDA(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    dev_free(d_flag);
    m.finite = !h_flag;
    return CGPU_OK;
}

static int reps_download(const DevReps *r, double *const *rep_out, cudaStream_t st) {
    for (int s = 0; s < r->S; s++) {
        if (r->nb[s] <= 0) continue;
        // rank-local compact block K x nb[s] (== the whole K x bb[s] block when unsharded)
        CGPU_CUDA(cudaMemcpyAsync(rep_out[s], r->d_rep[s], (size_t)r->K * r->nb[s] * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
    }
    CGPU_CUDA(cudaStreamSynchronize(st));
    return CGPU_OK;
}

}  // namespace cgpu

using namespace cgpu;

extern "C" {

int cgpu_version(void) { return 100; }

void cgpu_host_rowsums(const double *L, int K, int n, double *datvec) {
    if (!L || !datvec || K <= 0) return;
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) hw = std::max(1u, hw / (unsigned)std::max(1, atoi(e)));  // one process per GPU shares the host
    int nt = (int)std::min<unsigned>(hw ? hw : 1, (unsigned)K);
    if ((long long)K * n < (1 << 20)) nt = 1;
    auto work = [&](int t) {
        for (int i = t; i < K; i += nt) {
            const double *row = L + (size_t)i * n;
            double x = 0;
            for (int j = 0; j < n; j++) x += row[j];  // -ffp-contract=off, no reassociation: one rounded add per site
            datvec[i] = x;
        }
    };
    if (nt == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
}

const char *cgpu_last_error(void) { return t_err; }

int cgpu_init(int device) {
    Globals &g = G();
    if (g.inited) return CGPU_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error(CGPU_ENODEV, "no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        return CGPU_ENODEV;
    }
    if (device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % count : 0;
    }
    if (device >= count) {
        set_error(CGPU_EINVAL, "device %d of %d", device, count);
        return CGPU_EINVAL;
    }
    CGPU_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CGPU_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(CGPU_ENODEV, "device %d is sm_%d%d; libconselgpu is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
        return CGPU_ENODEV;
    }
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    {
        // the main stream outranks the aux stream: when both have blocks pending (staged index
        // production), the scale that is next to be copied out is served first
        int pri_lo = 0, pri_hi = 0;
        CGPU_CUDA(cudaDeviceGetStreamPriorityRange(&pri_lo, &pri_hi));
        CGPU_CUDA(cudaStreamCreateWithPriority(&g.stream, cudaStreamNonBlocking, pri_hi));
        CGPU_CUDA(cudaStreamCreateWithPriority(&g.stream2, cudaStreamNonBlocking, pri_lo));
    }
    CGPU_CUDA(cudaStreamCreateWithFlags(&g.stream3, cudaStreamNonBlocking));
    {
        // keep freed blocks in the default stream-ordered pool: the per-call workspaces
        // (index/count buffers, replicate blocks) are then recycled without a cudaMalloc
        cudaMemPool_t pool;
        CGPU_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long thresh = ~0ull;
        CGPU_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    }
    const Gf2Poly *T = mt_jump_table();
    {
        std::vector<uint32_t> pos((size_t)(MT_MAXLOG + 1) * MT_POS_STRIDE, 0u);
        for (int i = 0; i <= MT_MAXLOG; i++) g.jump_nquads[i] = mt_poly_positions(&T[i], pos.data() + (size_t)i * MT_POS_STRIDE);
        CGPU_CUDA(dev_alloc(&g.d_jump_table, pos.size() * sizeof(uint32_t)));
        CGPU_CUDA(cudaMemcpy(g.d_jump_table, pos.data(), pos.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    g.inited = true;
    return CGPU_OK;
}

void cgpu_shutdown(void) {
    Globals &g = G();
    if (!g.inited) return;
    cgpu_comm_destroy();  // (every rank of a job should get here: NCCL tears a communicator down together)
    cudaStreamSynchronize(g.stream);
    dev_free(g.d_jump_table);
    jump_cache_clear();
    for (int i = 0; i < 4; i++) {
        dev_free(g.scratch[i]);
        g.scratch[i] = nullptr;
        g.scratch_cap[i] = 0;
    }
    cudaStreamSynchronize(g.stream);
    cudaStreamDestroy(g.stream);
    cudaStreamDestroy(g.stream2);
    cudaStreamDestroy(g.stream3);
    g.stream2 = g.stream3 = nullptr;
    g.d_jump_table = nullptr;
    g.stream = nullptr;
    g.inited = false;
}

int cgpu_set_shard(int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) {
        set_error(CGPU_EINVAL, "shard %d/%d", rank, world);
        return CGPU_EINVAL;
    }
    G().shard_rank = rank;
    G().shard_world = world;
    return CGPU_OK;
}

int cgpu_shard_range(int bb, int *b_begin, int *b_end) {
    if (bb < 0 || !b_begin || !b_end) return CGPU_EINVAL;
    shard_range(bb, b_begin, b_end);
    return CGPU_OK;
}

long long cgpu_launch_count(void) { return g_launches.load(); }

int cgpu_set_option(const char *name, int value) {
    Globals &g = G();
    if (!name) return CGPU_EINVAL;
    if (!strcmp(name, "overlap")) g.opt_overlap = value;
    else if (!strcmp(name, "hist_algo")) g.opt_hist_algo = value;
    else if (!strcmp(name, "contract_algo")) g.opt_contract_algo = value;
    else if (!strcmp(name, "i8_oneshot")) g.opt_i8_oneshot = value;
    else if (!strcmp(name, "draw_nch")) g.opt_draw_nch = value;
    else if (!strcmp(name, "i8_fuse")) g.opt_i8_fuse = value;
    else if (!strcmp(name, "draw_fuse")) g.opt_draw_fuse = value;
    else if (!strcmp(name, "stage")) g.opt_stage = value;
    else if (!strcmp(name, "i8_dbg")) g.opt_i8_dbg = value;
    else if (!strcmp(name, "mc_smooth")) g.opt_mc_smooth = value;
    else if (!strcmp(name, "chunk_log2")) g.opt_chunk_log2 = value;
    else if (!strcmp(name, "chunk_fill")) g.opt_chunk_fill = value;
    else if (!strcmp(name, "ktime")) g.opt_ktime = value;
    else if (!strcmp(name, "hist_wide")) g.opt_hist_wide = value;
    else if (!strcmp(name, "i8_2cta")) g.opt_i8_2cta = value;
    else if (!strcmp(name, "batch")) g.opt_batch = value;
    else if (!strcmp(name, "i8_panel")) g.opt_i8_panel = value;
    else if (!strcmp(name, "i8_patterns")) g.opt_i8_patterns = value;
    else if (!strcmp(name, "i8_ksplit")) g.opt_i8_ksplit = value;
    else if (!strcmp(name, "dist_upload")) g.opt_dist_upload = value;
    else if (!strcmp(name, "hist_bucket")) g.opt_hist_bucket = value;
    else {
        set_error(CGPU_EINVAL, "unknown option %s", name);
        return CGPU_EINVAL;
    }
    return CGPU_OK;
}

int cgpu_kernel_times(char *buf, int len) {
    if (!buf || len <= 0) return CGPU_EINVAL;
    cudaDeviceSynchronize();
    struct Agg {
        int first = 0;
        double second = 0, work = 0;
    };
    std::map<std::string, Agg> agg;
    std::vector<std::string> order;
    for (auto &sp : kspans()) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            if (!agg.count(sp.name)) order.push_back(sp.name);
            agg[sp.name].first++;
            agg[sp.name].second += ms;
            agg[sp.name].work += sp.work;
        }
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    kspans().clear();
    (void)cudaGetLastError();
    int off = 0;
    buf[0] = 0;
    for (auto &n : order) {
        int w = snprintf(buf + off, (size_t)(len - off), "%s %d %.6f %.6e\n", n.c_str(), agg[n].first, agg[n].second, agg[n].work);
        if (w < 0 || w >= len - off) break;
        off += w;
    }
    return CGPU_OK;
}

int cgpu_matrix_info(const cgpu_matrix *mm, int *K, int *n, int *NP, int *slices, int *frac_bits, long long *mpad, int *patterns) {
    const DevMatrix *m = reinterpret_cast<const DevMatrix *>(mm);
    if (!m) return CGPU_EINVAL;
    if (K) *K = m->K;
    if (n) *n = m->n;
    if (NP) *NP = m->NP;
    if (slices) *slices = m->fxS;
    if (frac_bits) *frac_bits = m->fxF;
    if (mpad) *mpad = (long long)m->fxMpad;
    if (patterns) *patterns = m->npat;
    return CGPU_OK;
}

int cgpu_last_timing(double *ms8) {
    if (!ms8) return CGPU_EINVAL;
    memcpy(ms8, G().last_ms, sizeof(double) * 8);
    return CGPU_OK;
}

int cgpu_last_timing_ex(double *ms, int n) {
    if (!ms || n < 0) return CGPU_EINVAL;
    for (int i = 0; i < n; i++) ms[i] = i < PhaseTimer::NPHASE ? G().last_ms[i] : 0.0;
    return CGPU_OK;
}

int cgpu_matrix_upload(const double *L, int K, int n, cgpu_matrix **out) {
    CGPU_TRY(ensure_init());
    if (!out) return CGPU_EINVAL;
    DevMatrix *m = nullptr;
    CGPU_TRY(matrix_upload(L, K, n, &m, G().stream));
    CGPU_CUDA(cudaStreamSynchronize(G().stream));  // L belongs to the caller again when this returns
    *out = reinterpret_cast<cgpu_matrix *>(m);
    return CGPU_OK;
}

void cgpu_matrix_free(cgpu_matrix *mm) {
    DevMatrix *m = reinterpret_cast<DevMatrix *>(mm);
    if (!m) return;
    dev_free(m->d_L);
    dev_free(m->d_LT);
    dev_free(m->d_A8);
    dev_free(m->d_pid);
    if (m->ready) cudaEventDestroy(m->ready);
    delete m;
}

int cgpu_scaleboot_dev(const cgpu_matrix *mm, int S, const int *bn, const int *bb, unsigned long seed,
                       unsigned long long stream_offset, int rng_mode, int sum_mode, cgpu_reps **out) {
    CGPU_TRY(ensure_init());
    DevMatrix *m = const_cast<DevMatrix *>(reinterpret_cast<const DevMatrix *>(mm));  // lazily caches the int8 slices
    if (!m || S <= 0 || !bn || !bb || !out) {
        set_error(CGPU_EINVAL, "scaleboot_dev: bad arguments");
        return CGPU_EINVAL;
    }
    DevReps *r = nullptr;
    CGPU_TRY(alloc_reps(m->K, S, bb, &r));
    PhaseTimer pt(G().stream);
    int rc = scaleboot_core(*m, S, bn, bb, seed, stream_offset, rng_mode, sum_mode, r, pt, G().stream);
    cudaStreamSynchronize(G().stream);
    pt.finish(G().last_ms);
    if (rc != CGPU_OK) {
        free_reps(r);
        return rc;
    }
    *out = reinterpret_cast<cgpu_reps *>(r);
    return CGPU_OK;
}

int cgpu_reps_download(const cgpu_reps *rr, double *const *rep_out) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !rep_out) return CGPU_EINVAL;
    return reps_download(r, rep_out, G().stream);
}

int cgpu_reps_download_cols(const cgpu_reps *rr, int s, int col0, int ncols, double *out) {
    CGPU_TRY(ensure_init());
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || !out || s < 0 || s >= r->S || !r->d_rep || col0 < 0 || ncols < 0 || col0 + ncols > r->nb[s]) {
        set_error(CGPU_EINVAL, "reps_download_cols: bad arguments");
        return CGPU_EINVAL;
    }
    if (ncols == 0) return CGPU_OK;
    CGPU_CUDA(cudaMemcpy2DAsync(out, (size_t)ncols * 8, r->d_rep[s] + col0, (size_t)r->nb[s] * 8, (size_t)ncols * 8, r->K,
                                cudaMemcpyDeviceToHost, G().stream));
    CGPU_CUDA(cudaStreamSynchronize(G().stream));
    return CGPU_OK;
}

unsigned long long cgpu_last_chunk_len(void) { return g_last_chunk_len; }

void cgpu_reps_free(cgpu_reps *rr) { free_reps(reinterpret_cast<DevReps *>(rr)); }

int cgpu_reps_devptr(const cgpu_reps *rr, int s, void **ptr, int *rows, int *cols, int *ld) {
    const DevReps *r = reinterpret_cast<const DevReps *>(rr);
    if (!r || s < 0 || s >= r->S) return CGPU_EINVAL;
    if (ptr) *ptr = r->d_rep[s];
    if (rows) *rows = r->K;
    if (cols) *cols = r->nb[s];
    if (ld) *ld = r->nb[s];
    return CGPU_OK;
}

int cgpu_scaleboot_all(const double *L, int K, int n, int S, const int *bn, const int *bb, unsigned long seed,
                       unsigned long long stream_offset, int rng_mode, int sum_mode, double *const *rep_out) {
    CGPU_TRY(ensure_init());
    if (!L || !bn || !bb || !rep_out || S <= 0) {
        set_error(CGPU_EINVAL, "scaleboot_all: bad arguments");
        return CGPU_EINVAL;
    }
    cudaStream_t st = G().stream;
    PhaseTimer pt(st);
    DevMatrix *m = nullptr;
    // int8 mode: the upload runs on the copy stream under the jump tree, the draws and the first
    // histogram (none of them reads L); with a communicator the ranks share the upload (collective)
    cudaStream_t up = sum_mode == CGPU_SUM_INT8 ? G().stream3 : st;
    pt.begin(0, up);
    int rc = matrix_upload(L, K, n, &m, up, true);
    pt.end();
    DevReps *r = nullptr;
    if (rc == CGPU_OK) rc = alloc_reps(K, S, bb, &r);
    if (rc == CGPU_OK) rc = scaleboot_core(*m, S, bn, bb, seed, stream_offset, rng_mode, sum_mode, r, pt, st, rep_out);
    cudaStreamSynchronize(st);
    // every exit path: the upload and the device-to-host copies of the copy stream are finished
    // before the blocks they touch are freed and before the caller gets its buffers back
    cudaStreamSynchronize(G().stream3);
    pt.finish(G().last_ms);
    free_reps(r);
    cgpu_matrix_free(reinterpret_cast<cgpu_matrix *>(m));
    return rc;
}

int cgpu_resample_counts(int n, int S, const int *bn, const int *bb, unsigned long seed,
                         unsigned long long stream_offset, int rng_mode, int s, unsigned char *counts) {
    CGPU_TRY(ensure_init());
    if (n <= 0 || S <= 0 || !bn || !bb || s < 0 || s >= S || !counts) {
        set_error(CGPU_EINVAL, "resample_counts: bad arguments");
        return CGPU_EINVAL;
    }
    cudaStream_t st = G().stream;
    int b0, b1;
    shard_range(bb[s], &b0, &b1);
    const int nb = b1 - b0;
    if (nb <= 0) return CGPU_OK;
    const int NP = pad_sites(n);
    Workspace ws;
    const size_t cbytes = (size_t)((nb + 7) / 8) * NP * 8;
    CGPU_CUDA(dev_alloc(&ws.d_counts[0], cbytes));
    CGPU_CUDA(cudaMemsetAsync(ws.d_counts[0], 0, cbytes, st));
    CGPU_CUDA(dev_alloc(&ws.d_err, sizeof(int)));
    CGPU_CUDA(cudaMemsetAsync(ws.d_err, 0, sizeof(int), st));
    MtPlan plan;
    struct PlanGuard {
        MtPlan &p;
        ~PlanGuard() {
            dev_free(p.d_states);
            dev_free(p.d_pairs);
        }
    } guard{plan};
    if (rng_mode == CGPU_RNG_MT19937) {
        unsigned long long base = 0, total = 0;
        for (int t = 0; t < S; t++) {
            if (t < s) base += (unsigned long long)bb[t] * bn[t];
            total += (unsigned long long)bb[t] * bn[t];
        }
        CGPU_CUDA(dev_alloc(&ws.d_idx, (size_t)nb * bn[s] * idx_bytes((uint32_t)n)));
        const unsigned long long lo = base + (unsigned long long)b0 * bn[s], hi = lo + (unsigned long long)nb * bn[s];
        CGPU_TRY(mt_prepare(seed, stream_offset, total, 1, &lo, &hi, &plan, st));
        CGPU_TRY(launch_mt_draw_range(plan, lo, hi, (uint32_t)n, ws.d_idx, st));
        CGPU_TRY(launch_hist_from_idx(ws.d_idx, n, NP, bn[s], nb, ws.d_counts[0], ws.d_err, st));
    } else if (rng_mode == CGPU_RNG_PHILOX) {
        CGPU_TRY(launch_hist_philox((unsigned long long)seed, s, b0, n, NP, bn[s], nb, ws.d_counts[0], ws.d_err, st, false, stream_offset));
    } else {
        set_error(CGPU_EINVAL, "unknown rng_mode %d", rng_mode);
        return CGPU_EINVAL;
    }
    unsigned char *d_out = nullptr;
    CGPU_CUDA(dev_alloc(&d_out, (size_t)nb * n));
    int rc = launch_counts_unpack(ws.d_counts[0], n, NP, nb, d_out, st);
    int h_err = 0;
    if (rc == CGPU_OK) {
        cudaMemcpyAsync(counts, d_out, (size_t)nb * n, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&h_err, ws.d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error(CGPU_ECUDA, "resample_counts: %s", cudaGetErrorString(e));
            rc = CGPU_ECUDA;
        }
    }
    dev_free(d_out);
    if (rc == CGPU_OK && h_err) {
        set_error(CGPU_ERANGE, "a site was drawn more than 255 times in one replicate");
        rc = CGPU_ERANGE;
    }
    return rc;
}

int cgpu_mt19937_stream(unsigned long seed, unsigned long long skip, unsigned long long count, unsigned int *out) {
    CGPU_TRY(ensure_init());
    if (!out && count) return CGPU_EINVAL;
    if (count == 0) return CGPU_OK;
    cudaStream_t st = G().stream;
    MtPlan plan;
    struct PlanGuard {
        MtPlan &p;
        ~PlanGuard() {
            dev_free(p.d_states);
            dev_free(p.d_pairs);
        }
    } guard{plan};
    const unsigned long long lo = skip, hi = skip + count;
    CGPU_TRY(mt_prepare(seed, 0, skip + count, 1, &lo, &hi, &plan, st));
    uint32_t *d_idx = nullptr;
    CGPU_CUDA(dev_alloc(&d_idx, (size_t)count * sizeof(uint32_t)));
    // n = 0 asks for the raw tempered words
    int rc = launch_mt_draw_range(plan, lo, hi, 0u, d_idx, st);
    if (rc == CGPU_OK) {
        cudaMemcpyAsync(out, d_idx, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error(CGPU_ECUDA, "mt19937_stream: %s", cudaGetErrorString(e));
            rc = CGPU_ECUDA;
        }
    }
    dev_free(d_idx);
    return rc;
}

}  // extern "C"
