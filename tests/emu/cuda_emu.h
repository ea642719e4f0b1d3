// CPU execution of plain CUDA kernels, for tests (test infrastructure only: nothing of the
// product includes this file).
//
// A kernel is compiled as ordinary C++ (g++ -DCGPU_EMU) and launched with emu::launch(grid, block,
// dynamic_smem_bytes, kernel, args...): ONE OS THREAD PER CUDA THREAD of a CTA, the CTAs of the
// grid run one after another on the same threads.  threadIdx / blockIdx / blockDim / gridDim are
// thread-local, __syncthreads is a real barrier over the CTA's threads, __shfl_*_sync exchange
// through a per-warp mailbox behind warp barriers, atomics are the compiler's __atomic builtins.
// Function-scope `__shared__` arrays become function-local statics (shared by the CTA's threads;
// CTAs never overlap in time), dynamic shared memory is one buffer per launch (CGPU_DYN_SMEM).
// Because every access is a real memory access of a real thread, ThreadSanitizer sees a missing
// barrier or an unsynchronised shared-memory update as a data race.
//
// Two execution modes, chosen at compile time:
//   default            one OS thread per CUDA thread, CTAs one after another (what the sanitizers see)
//   -DCGPU_EMU_FIBERS  the threads of a CTA are coroutines (ucontext) of ONE OS thread that switch at barriers
//                      and warp collectives; several CTAs run at once on the host's cores.  ~50x faster: for
//                      end-to-end runs of the emulated library.  `__shared__` arrays are then thread_local
//                      statics (per OS thread = per CTA).
// A thread that returns from the kernel counts as arrived at every later barrier (as on the GPU) in fiber mode;
// the OS-thread mode requires all threads of a CTA (warp) to reach every barrier (collective).
//
// Supported: what consel_b200/csrc/{hist_bucket.cuh,philox.cuh,stats_kernels.cu,resample_kernels.cu} use.  Warp-collectives must be
// called by all 32 lanes of a warp (as CUDA requires with a full mask).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef CGPU_EMU_FIBERS
#include <sys/mman.h>
#if !defined(__x86_64__) || defined(CGPU_EMU_UCONTEXT)
#include <ucontext.h>
#endif

#include <atomic>
#endif
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) uint4 {  // 16-byte aligned like CUDA's: -fsanitize=alignment flags a misaligned vector access
    unsigned x, y, z, w;
};
struct alignas(16) double2 {
    double x, y;
};
struct alignas(16) double4 {
    double x, y, z, w;
};
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace emu {
#ifndef CGPU_EMU_FIBERS
class Barrier {  // reusable barrier (mutex + condition variable: visible to ThreadSanitizer)
  public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const unsigned gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            gen_++;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }
    void reset(int n) { n_ = n; count_ = 0; }
    void drop() {}

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0;
    unsigned gen_ = 0;
};
#else
void yield();              // back to the CTA's scheduler (next runnable coroutine of the CTA)
void yield_in_warp(int w);  // straight to the next live lane of warp w: a warp collective costs 32 switches, not a CTA round
class Barrier {  // all participants are coroutines of one OS thread: plain counters
  public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        const unsigned gen = gen_;
        if (++count_ >= n_) {
            count_ = 0;
            gen_++;
        } else if (warp_ >= 0) {
            while (gen_ == gen) yield_in_warp(warp_);
        } else {
            while (gen_ == gen) yield();
        }
    }
    void set_warp(int w) { warp_ = w; }
    void reset(int n) { n_ = n; count_ = 0; }
    void drop() {  // a participant returned from the kernel: it counts as arrived from now on
        n_--;
        if (n_ > 0 && count_ >= n_) {
            count_ = 0;
            gen_++;
        }
    }

  private:
    int n_, count_ = 0, warp_ = -1;
    unsigned gen_ = 0;
};
#endif
struct Warp {
    Barrier bar{32};
    // mailboxes of the warp collectives, double buffered: collective k uses buffer k & 1, so ONE barrier per
    // collective suffices (a lane can be at most one collective ahead of the slowest lane of its warp: it cannot
    // pass barrier k+1 before every lane has read buffer k & 1 and arrived there)
    unsigned long long mail[2][32];
    unsigned seq[32] = {0};
    int lanes;
    explicit Warp(int n) : bar(n), lanes(n) {}
};
struct Cta {
    Barrier *bar = nullptr;
    std::vector<Warp *> warps;
    unsigned char *dyn = nullptr;
};
inline thread_local Cta *t_cta = nullptr;
inline thread_local int t_lane = 0, t_warp = 0;
inline unsigned char *dyn_smem() { return t_cta->dyn; }
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#ifdef CGPU_EMU_FIBERS
#define __shared__ static thread_local
#else
#define __shared__ static
#endif
#define CGPU_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem())

static inline void __syncthreads() { emu::t_cta->bar->wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_cta->warps[emu::t_warp]->bar.wait(); }
// warp collectives: every lane deposits 64 bits in the warp's mailbox, barrier, read
static inline unsigned long long emu_exchange64(unsigned long long v, int src_lane) {
    emu::Warp *w = emu::t_cta->warps[emu::t_warp];
    const unsigned k = w->seq[emu::t_lane]++ & 1u;
    w->mail[k][emu::t_lane] = v;
    w->bar.wait();
    return w->mail[k][src_lane & 31];
}
template <class T>
static inline T emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    bits = emu_exchange64(bits, src_lane);
    T r;
    memcpy(&r, &bits, sizeof(T));
    return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int off) { return emu_shfl(v, emu::t_lane ^ off); }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_shfl(v, emu::t_lane >= d ? emu::t_lane - d : emu::t_lane); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_shfl(v, emu::t_lane + d < 32 ? emu::t_lane + d : emu::t_lane); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int lane) { return emu_shfl(v, lane); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Warp *w = emu::t_cta->warps[emu::t_warp];
    const unsigned k = w->seq[emu::t_lane]++ & 1u;
    w->mail[k][emu::t_lane] = pred ? 1ull : 0ull;
    w->bar.wait();
    unsigned m = 0;
    for (int l = 0; l < w->lanes; l++) m |= (unsigned)w->mail[k][l] << l;
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
    return __ballot_sync(mask, !pred) == 0;
}
// lanes of `mask` whose value equals mine (called by exactly the lanes of `mask`: the others must still
// take part in the mailbox exchange, so this emulation requires a full warp to call it -- as the kernels do)
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    emu::Warp *w = emu::t_cta->warps[emu::t_warp];
    const unsigned k = w->seq[emu::t_lane]++ & 1u;
    w->mail[k][emu::t_lane] = v;
    w->bar.wait();
    unsigned m = 0;
    for (int l = 0; l < w->lanes; l++)
        if (((mask >> l) & 1u) && (unsigned)w->mail[k][l] == v) m |= 1u << l;
    return m;
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicXor(unsigned *p, unsigned v) { return __atomic_fetch_xor(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
// IEEE round-to-nearest fp64 operations, one rounding each (compile with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return __builtin_sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline double __uint2double_rn(unsigned x) { return (double)x; }
static inline size_t __cvta_generic_to_shared(const void *p) { return reinterpret_cast<size_t>(p); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const unsigned long long ab = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s4 = (sel >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(ab >> (8 * (s4 & 7u))) & 0xffu;
        if (s4 & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;  // replicate the sign bit
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;

#ifdef CGPU_EMU_FIBERS
// ---- coroutine switch.  x86-64: a dozen instructions (callee-saved registers + stack pointer); elsewhere
// ucontext (whose swapcontext makes two signal-mask system calls per switch: ~50x slower).
#if defined(__x86_64__) && !defined(CGPU_EMU_UCONTEXT)
#define CGPU_EMU_ASM_SWITCH 1
extern "C" void cgpu_emu_switch(void **save_sp, void *load_sp);
__asm__(
    ".text\n"
    ".weak cgpu_emu_switch\n"
    ".type cgpu_emu_switch,@function\n"
    "cgpu_emu_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size cgpu_emu_switch,.-cgpu_emu_switch\n");
#endif
namespace emu {
struct Fiber {
#ifdef CGPU_EMU_ASM_SWITCH
    void *sp = nullptr;
#else
    ucontext_t ctx;
#endif
    bool done = false;
};
struct Worker {  // one OS thread: runs CTAs one at a time, the CTA's threads as coroutines
#ifdef CGPU_EMU_ASM_SWITCH
    void *sched_sp = nullptr;
#else
    ucontext_t sched;
#endif
    std::vector<Fiber> fibers;
    std::vector<void *> stacks;
    int cur = 0;
    const std::function<void()> *body = nullptr;
};
inline thread_local Worker *t_worker = nullptr;
inline void yield() {
    Worker *w = t_worker;
#ifdef CGPU_EMU_ASM_SWITCH
    cgpu_emu_switch(&w->fibers[w->cur].sp, w->sched_sp);
#else
    swapcontext(&w->fibers[w->cur].ctx, &w->sched);
#endif
}
struct WorkerDims {
    dim3 block;
    int nthreads = 0;
};
inline thread_local WorkerDims t_dims;
inline void become(int t) {  // the identity of coroutine t in the thread-locals the kernels read
    t_worker->cur = t;
    t_lane = t & 31;
    t_warp = t >> 5;
    const dim3 &b = t_dims.block;
    threadIdx = dim3((unsigned)t % b.x, ((unsigned)t / b.x) % b.y, (unsigned)t / (b.x * b.y));
}
inline void yield_in_warp(int wp) {
    Worker *w = t_worker;
    const int me = w->cur, base = wp * 32, lanes = std::min(32, t_dims.nthreads - base);
    for (int d = 1; d < lanes; d++) {
        const int t = base + ((me - base) + d) % lanes;
        if (w->fibers[t].done) continue;
#ifdef CGPU_EMU_ASM_SWITCH
        if (!w->fibers[t].sp) continue;
        become(t);
        cgpu_emu_switch(&w->fibers[me].sp, w->fibers[t].sp);
        become(me);  // (whoever switched back to us restored nothing: re-assert our identity)
        return;
#else
        become(t);
        swapcontext(&w->fibers[me].ctx, &w->fibers[t].ctx);
        become(me);
        return;
#endif
    }
    yield();  // no other live lane in this warp
}
inline void fiber_entry() {
    Worker *w = t_worker;
    (*w->body)();
    Fiber &f = w->fibers[w->cur];
    f.done = true;
    t_cta->bar->drop();
    t_cta->warps[t_warp]->bar.drop();
#ifdef CGPU_EMU_ASM_SWITCH
    cgpu_emu_switch(&f.sp, w->sched_sp);  // never resumed
    __builtin_trap();
#endif
    // (ucontext: returns to the scheduler through uc_link)
}
constexpr size_t kFiberStack = 256 * 1024;
// coroutine stacks are recycled across launches (a launch needs workers x threads of them)
inline std::mutex g_stack_mu;
inline std::vector<void *> g_stack_pool;
inline void stack_get_many(std::vector<void *> &out) {
    size_t have = 0;
    {
        std::lock_guard<std::mutex> lk(g_stack_mu);
        while (have < out.size() && !g_stack_pool.empty()) {
            out[have++] = g_stack_pool.back();
            g_stack_pool.pop_back();
        }
    }
    for (; have < out.size(); have++)
        out[have] = mmap(nullptr, kFiberStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
}
inline void stack_put_many(const std::vector<void *> &v) {
    std::lock_guard<std::mutex> lk(g_stack_mu);
    g_stack_pool.insert(g_stack_pool.end(), v.begin(), v.end());
}

template <class K, class... A>
void launch(dim3 grid, dim3 block, size_t smem, K kernel, A... args) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    const long long nctas = (long long)grid.x * grid.y * grid.z;
    if (nctas <= 0 || nthreads <= 0) return;
    const std::function<void()> body = [&] { kernel(args...); };
    std::atomic<long long> next{0};
    const int nworkers = (int)std::min<long long>(nctas, std::max(1u, std::thread::hardware_concurrency()));
    auto work = [&] {
        Worker wk;
        wk.body = &body;
        wk.fibers.resize(nthreads);
        wk.stacks.resize(nthreads);
        stack_get_many(wk.stacks);
        t_worker = &wk;
        Cta cta;
        Barrier bar(nthreads);
        cta.bar = &bar;
        for (int w = 0; w < nwarps; w++) {
            cta.warps.push_back(new Warp(std::min(32, nthreads - 32 * w)));
            cta.warps.back()->bar.set_warp(w);
        }
        unsigned char *dynbuf = static_cast<unsigned char *>(aligned_alloc(128, std::max<size_t>(128, (smem + 127) / 128 * 128)));
        cta.dyn = dynbuf;
        t_cta = &cta;
        t_dims.block = block;
        t_dims.nthreads = nthreads;
        blockDim = block;
        gridDim = grid;
        for (;;) {
            const long long c = next.fetch_add(1);
            if (c >= nctas) break;
            blockIdx = dim3((unsigned)(c % grid.x), (unsigned)((c / grid.x) % grid.y), (unsigned)(c / ((long long)grid.x * grid.y)));
            bar.reset(nthreads);
            for (int w = 0; w < nwarps; w++) {
                cta.warps[w]->bar.reset(cta.warps[w]->lanes);
                memset(cta.warps[w]->seq, 0, sizeof(cta.warps[w]->seq));
            }
            for (int t = 0; t < nthreads; t++) {
                Fiber &f = wk.fibers[t];
                f.done = false;
#ifdef CGPU_EMU_ASM_SWITCH
                // initial frame: six zeroed callee-saved registers, then the entry point as return address; after the
                // `ret` the stack pointer is 8 mod 16, as at any function entry
                uintptr_t top = (reinterpret_cast<uintptr_t>(wk.stacks[t]) + kFiberStack) & ~(uintptr_t)15;
                void **frame = reinterpret_cast<void **>(top - 64);
                for (int i = 0; i < 6; i++) frame[i] = nullptr;
                frame[6] = reinterpret_cast<void *>(&fiber_entry);
                frame[7] = nullptr;
                f.sp = frame;
#else
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = wk.stacks[t];
                f.ctx.uc_stack.ss_size = kFiberStack;
                f.ctx.uc_link = &wk.sched;
                makecontext(&f.ctx, fiber_entry, 0);
#endif
            }
            int alive = nthreads;
            while (alive > 0) {
                alive = 0;
                for (int t = 0; t < nthreads; t++) {
                    Fiber &f = wk.fibers[t];
                    if (f.done) continue;
                    become(t);
#ifdef CGPU_EMU_ASM_SWITCH
                    cgpu_emu_switch(&wk.sched_sp, f.sp);
#else
                    swapcontext(&wk.sched, &f.ctx);
#endif
                    if (!f.done) alive++;
                }
            }
        }
        for (Warp *w : cta.warps) delete w;
        free(dynbuf);
        stack_put_many(wk.stacks);
        t_worker = nullptr;
        t_cta = nullptr;
    };
    if (nworkers == 1) {
        std::thread(work).join();  // a fresh OS thread: the caller's thread-locals stay untouched
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nworkers; i++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
}
}  // namespace emu
#else
namespace emu {
// launch<<<grid, block, smem>>>: CTAs in order (x fastest), the block's threads persist over the CTAs
template <class K, class... A>
void launch(dim3 grid, dim3 block, size_t smem, K kernel, A... args) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    Cta cta;
    Barrier bar(nthreads);
    cta.bar = &bar;
    for (int w = 0; w < nwarps; w++) cta.warps.push_back(new Warp(std::min(32, nthreads - 32 * w)));
    // exactly `smem` bytes, 16-byte aligned like the hardware's window: AddressSanitizer sees an overrun
    unsigned char *dynbuf = static_cast<unsigned char *>(aligned_alloc(128, std::max<size_t>(128, (smem + 127) / 128 * 128)));
    cta.dyn = dynbuf;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t] {
            t_cta = &cta;
            t_lane = t & 31;
            t_warp = t >> 5;
            blockDim = block;
            gridDim = grid;
            threadIdx = dim3((unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        blockIdx = dim3(bx, by, bz);
                        cta.warps[t_warp]->seq[t_lane] = 0;
                        kernel(args...);
                        bar.wait();  // the next CTA reuses the shared memory
                    }
        });
    for (auto &t : th) t.join();
    for (Warp *w : cta.warps) delete w;
    free(dynbuf);
}
}  // namespace emu
#endif  // CGPU_EMU_FIBERS
