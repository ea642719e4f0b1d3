// CPU stand-ins for the PTX helpers of consel_b200/csrc/contract_kernels.cu (mbarrier + cp.async.bulk), used
// by tests/emu/build_emu_lib.py in place of the inline-PTX block of that file.
//
// An mbarrier is one 64-bit word in shared memory: phase bit | arrival count per phase | pending arrivals |
// outstanding transaction bytes.  A phase completes when both pending arrivals and outstanding bytes reach
// zero (PTX ISA, "mbarrier": arrive, expect-tx, complete-tx); try_wait.parity(P) succeeds once the phase
// with parity P is complete, i.e. the barrier's current phase bit differs from P.  The bulk copy is a memcpy
// followed by complete_tx.  All updates are compare-and-swap loops, so the same code serves the coroutine
// and the OS-thread mode of cuda_emu.h.
#pragma once
#include <stdint.h>
#include <string.h>

#include <thread>

namespace cgpu {
namespace mbar_emu {
struct State {
    uint32_t phase;    // 0 / 1
    uint32_t count;    // arrivals a phase expects
    int32_t pending;   // arrivals still missing in the current phase
    int64_t tx;        // bytes announced (expect_tx) minus bytes delivered (complete_tx)
};
inline State unpack(uint64_t w) {
    State s;
    s.phase = (uint32_t)(w >> 63);
    s.count = (uint32_t)((w >> 52) & 0x7ffu);
    s.pending = (int32_t)((w >> 40) & 0xfffu);
    if (s.pending & 0x800) s.pending -= 0x1000;
    s.tx = (int64_t)(w & 0xffffffffffull);
    if (s.tx & 0x8000000000ll) s.tx -= 0x10000000000ll;
    return s;
}
inline uint64_t pack(const State &s) {
    return ((uint64_t)(s.phase & 1u) << 63) | ((uint64_t)(s.count & 0x7ffu) << 52) | ((uint64_t)((uint32_t)s.pending & 0xfffu) << 40) |
           ((uint64_t)s.tx & 0xffffffffffull);
}
template <class F>
inline void update(uint64_t *bar, F f) {
    uint64_t old = __atomic_load_n(bar, __ATOMIC_ACQUIRE);
    for (;;) {
        State s = unpack(old);
        f(s);
        if (s.pending == 0 && s.tx == 0) {  // phase complete: next phase, arrivals re-armed
            s.phase ^= 1u;
            s.pending = (int32_t)s.count;
        }
        if (__atomic_compare_exchange_n(bar, &old, pack(s), false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) return;
    }
}
inline void relax() {
#ifdef CGPU_EMU_FIBERS
    emu::yield();
#else
    std::this_thread::yield();
#endif
}
}  // namespace mbar_emu

static inline void mbar_init(uint64_t *bar, uint32_t count) {
    mbar_emu::State s{0u, count, (int32_t)count, 0};
    __atomic_store_n(bar, mbar_emu::pack(s), __ATOMIC_RELEASE);
}
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {  // arrive + expect_tx
    mbar_emu::update(bar, [&](mbar_emu::State &s) { s.tx += bytes; s.pending -= 1; });
}
static inline void mbar_arrive(uint64_t *bar) {
    mbar_emu::update(bar, [&](mbar_emu::State &s) { s.pending -= 1; });
}
static inline void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (mbar_emu::unpack(__atomic_load_n(bar, __ATOMIC_ACQUIRE)).phase == (parity & 1u)) mbar_emu::relax();
}
static inline void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    memcpy(dst_smem, src_gmem, bytes);
    mbar_emu::update(bar, [&](mbar_emu::State &s) { s.tx -= bytes; });
}
}  // namespace cgpu
