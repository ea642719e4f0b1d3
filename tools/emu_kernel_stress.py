#!/usr/bin/env python
"""Randomised parity sweep of the product's CUDA sources EXECUTED ON THE CPU (tests/emu/: the real kernels and the
real C ABI, only the tensor-core MMA instruction replaced by a plain int32 product) against the oracle -- the
CPU twin of tools/stress.py, for odd shapes and option combinations the fixed tests do not hold:

  makermt side  strict == oracle makermt bit for bit; site counts == the oracle's; int8 == the correctly rounded
                exact sum (math.fsum over the drawn multiset); Philox counts are valid multinomials; shards ==
                column slices; options (chunking, staging, histogram variants, batching, pattern slots, fusion
                switches) never change a bit
  consel side   statistics == oracle repminmaxs for random associations; count queries == oracle cntdist on the
                sorted rows (integer and smoothed), order statistics, KH / SH / wKH / wSH == oracle do_mctest

    python tools/emu_kernel_stress.py [SEED=1] [CASES=40] [builddir=/tmp/emu_kstress]
"""
import ctypes as C
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "emu"))
import consel_b200 as cg  # noqa: E402
from consel_b200 import api, pipeline  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def bind(build: Path):
    import build_emu_lib
    lib = C.CDLL(str(build_emu_lib.build(build)))
    lib.cgpu_last_error.restype = C.c_char_p
    lib.cgpu_launch_count.restype = C.c_longlong
    lib.cgpu_smooth_count.restype = C.c_double
    api._lib = pipeline._host = lib
    cg.init()
    cg.set_shard(0, 1)


def exact(Lrow, c, r):
    return math.fsum(np.repeat(Lrow, c)) / np.float64(r)


def makermt_case(rng, case):
    big = rng.random() < 0.12
    if big:   # rows longer than every shared-memory histogram: global atomics / bucketed
        n = int(rng.integers(226000, 420000))
        S, K = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        bb = [int(rng.integers(1, 5)) for _ in range(S)]
    else:
        n = int(rng.choice([rng.integers(1, 40), rng.integers(40, 700), rng.integers(700, 9000), rng.integers(28000, 70000)]))
        S = int(rng.integers(1, 5))
        bb = [int(rng.integers(1, 90)) for _ in range(S)]
        while n * sum(bb) > 1.2e6:
            bb = [max(1, b // 2) for b in bb]
        K = int(rng.integers(1, max(2, min(70, int(4e6 / (n * sum(bb)))))))
    if os.environ.get("EMU_KSTRESS_WIDE") and not big:  # many trees on few sites: several tree tiles of the fp64 contraction, several digit-row tiles
        n = int(rng.integers(5, 400))
        K = int(rng.integers(70, 420))
        bb = [int(rng.integers(1, 40)) for _ in range(S)]
    rr = [float(x) for x in np.round(rng.uniform(0.3, 1.7, size=S), 3)]
    seed = int(rng.integers(1, 2 ** 31))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = base[None, :] + rng.choice([0.0, 0.106, 2.0]) * rng.standard_normal((K, n))
    if rng.random() < 0.3 and n > 8:   # repeated site patterns (pattern slots of the int8 path)
        src = rng.integers(0, n, size=n // 2)
        L[:, rng.integers(0, n, size=n // 2)] = L[:, src]
    if rng.random() < 0.15:
        L = np.round(L, 2)
    _, bn, rr1 = orc.scale_setup(L, rr)
    if min(bn) < 1:
        return "skipped (a scale with no draws)"
    opts = {"chunk_log2": int(rng.choice([-1, -1, 10, 12, 14, 17])), "stage": int(rng.choice([-1, 0, 1])),
            "i8_fuse": int(rng.choice([-1, 0])), "draw_fuse": int(rng.choice([-1, -1, 1])), "chunk_fill": int(rng.choice([-1, 0, 2])),
            "hist_wide": int(rng.choice([-1, 0, 1])), "hist_bucket": int(rng.choice([-1, 1, 2])), "i8_patterns": int(rng.choice([-1, 0])),
            "batch": int(rng.choice([-1, -1, 256])), "overlap": int(rng.choice([-1, 0, 1])), "i8_ksplit": int(rng.choice([-1, -1, 2]))}
    print(f"  (makermt case {case}: K={K} n={n} rr={rr} bb={bb} seed={seed} {opts})", flush=True)  # a crash names its case
    ref = orc.makermt(L, bn, bb, seed)
    counts = orc.scaleboot_counts(n, bn, bb, seed)
    base_i8 = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_INT8)
    base_ph = cg.scaleboot_all(L, bn, bb, seed, rng_mode=cg.RNG_PHILOX, sum_mode=cg.SUM_INT8)
    try:
        for k, v in opts.items():
            cg.set_option(k, v)
        strict = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_STRICT)
        i8 = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_INT8)
        fast = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_FAST)
        ph = cg.scaleboot_all(L, bn, bb, seed, rng_mode=cg.RNG_PHILOX, sum_mode=cg.SUM_INT8)
        for s in range(S):  # the counter-based resampler through the byte-row histograms (bucketed one included)
            assert np.array_equal(ph[s], base_ph[s]), ("options changed a Philox int8 bit", s)
        for s in range(S):
            assert np.array_equal(strict[s], ref[s]), ("strict != oracle", s)
            assert np.array_equal(i8[s], base_i8[s]), ("options changed an int8 bit", s)
            assert np.allclose(fast[s], ref[s], rtol=1e-11, atol=0), ("fast", s)
            got = cg.resample_counts(n, bn, bb, seed, s)
            assert np.array_equal(got, counts[s]), ("counts != oracle", s)
            for b in {0, bb[s] - 1}:
                for k in {0, K - 1}:
                    assert i8[s][k, b] == exact(L[k], counts[s][b].astype(np.int64), rr1[s]), ("int8 != exact", s, k, b)
            ph = cg.resample_counts(n, bn, bb, seed, s, rng_mode=cg.RNG_PHILOX)
            assert ph.shape == counts[s].shape and np.all(ph.sum(axis=1) == bn[s]), ("philox row sums", s)
        world = int(rng.integers(1, 4))
        if world > 1:
            rank = int(rng.integers(0, world))
            cg.set_shard(rank, world)
            loc8 = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_INT8)
            locs = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_STRICT)
            cg.set_shard(0, 1)
            for s in range(S):
                b0, b1 = int(bb[s]) * rank // world, int(bb[s]) * (rank + 1) // world
                assert np.array_equal(loc8[s], i8[s][:, b0:b1]), ("shard int8", rank, world, s)
                assert np.array_equal(locs[s], strict[s][:, b0:b1]), ("shard strict", rank, world, s)
    except AssertionError as e:
        raise AssertionError(f"makermt case {case}: K={K} n={n} rr={rr} bb={bb} seed={seed} {opts}: {e}") from None
    finally:
        cg.set_shard(0, 1)
        for k in opts:
            cg.set_option(k, -1)
    return f"K={K} n={n} S={S} bb={bb} {opts}"


def consel_case(rng, case):
    K = int(rng.choice([rng.integers(2, 12), rng.integers(12, 80), rng.integers(80, 300)]))
    S = int(rng.integers(1, 5))
    bb = [int(rng.choice([rng.integers(1, 40), rng.integers(40, 700), rng.integers(700, 3000)])) for _ in range(S)]
    while K * sum(bb) > 4e5:
        bb = [max(1, b // 2) for b in bb]
    reps = []
    centre = rng.normal(size=K) * rng.choice([0.0, 1.0, 30.0])
    for s in range(S):
        x = centre[:, None] + rng.normal(size=(K, bb[s])) * rng.choice([1e-3, 1.0, 50.0])
        if rng.random() < 0.3:
            x = np.round(x, 1)            # ties between replicates and between trees
        if rng.random() < 0.2 and K > 2:
            x[int(rng.integers(1, K))] = x[0]
        reps.append(np.ascontiguousarray(x))
    datvec = centre * 10 + rng.normal(size=K)
    if rng.random() < 0.5:
        ptr, idx = orc.identity_assoc(K)
    else:
        cm = int(rng.integers(1, min(K, 12) + 1))
        sets = [np.sort(rng.choice(K, size=int(rng.integers(1, K)), replace=False)) for _ in range(cm)]
        ptr = np.cumsum([0] + [len(x) for x in sets]).astype(np.int32)
        idx = np.concatenate(sets).astype(np.int32)
    cm = len(ptr) - 1
    try:
        R = cg.Reps.upload(reps)
        R.stats_build(ptr, idx)
        stats = R.stats_download(sorted_rows=False)
        srt = R.stats_download(sorted_rows=True)
        want = [orc.repminmaxs(reps[s], ptr, idx) for s in range(S)]
        for s in range(S):
            assert np.array_equal(stats[s], want[s]), ("statistics", s)
            assert np.array_equal(srt[s], np.sort(want[s], axis=1)), ("sorted statistics", s)
        nq = int(rng.integers(1, 5))
        thr = np.where(rng.random((cm, nq)) < 0.3, want[0][:, :1] if bb[0] else 0.0, rng.normal(size=(cm, nq)) * 3)
        q = R.count_query(thr)
        for i in range(cm):
            for s in range(S):
                row = np.sort(want[s][i])
                for a in range(nq):
                    assert q["cnt"][i, s, a] == orc.cntdist(row, thr[i, a], 2), ("count", i, s, a)
                    if bb[s] >= 2:  # (cntdist3 reads vec[1] and vec[bb-2]: with one replicate the reference itself is undefined)
                        sm = cg.smooth_count(q[i, s, a], bb[s], thr[i, a])
                        assert sm == orc.cntdist(row, thr[i, a], 3), ("smoothed count", i, s, a, sm)
        s0 = int(rng.integers(0, S))
        rank = rng.integers(0, bb[s0], size=cm).astype(np.int32)
        val = R.order_stat(s0, rank)
        assert np.array_equal(val, np.sort(want[s0], axis=1)[np.arange(cm), rank]), "order statistic"
        i0 = int(rng.integers(0, S))
        if bb[i0] >= 2:
            got = R.mctest(i0, datvec, ptr, idx)
            exp = orc.mctest(reps[i0], datvec, ptr, idx)
            for name, a, b in zip(("kh", "sh", "wkh", "wsh"), got, exp):
                assert np.array_equal(a, b, equal_nan=True), (name, a, b)
        R.free()
    except AssertionError as e:
        raise AssertionError(f"consel case {case}: K={K} bb={bb} cm={cm}: {e}") from None
    return f"K={K} S={S} bb={bb} cm={cm}"


def main() -> int:
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    bind(Path(sys.argv[3] if len(sys.argv) > 3 else "/tmp/emu_kstress"))
    rng = np.random.default_rng(seed)
    t0 = time.time()
    bad = 0
    for c in range(cases):
        for fn in (makermt_case, consel_case):
            try:
                print(f"case {c} {fn.__name__}: {fn(rng, c)}", flush=True)
            except (AssertionError, cg.CgpuError) as e:
                bad += 1
                print(f"case {c} {fn.__name__} FAILED: {type(e).__name__}: {e}", flush=True)
    print(f"emu_kernel_stress seed {seed}: {2 * cases - bad} of {2 * cases} checks passed in {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
