"""Randomised parity sweep on a GPU box (development aid; the pytest suite holds the fixed cases).
strict == oracle makermt bit for bit; int8 == big-integer exact sum; shards == slices; options
(chunk size, staging, fusion switches) never change a bit."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
from oracle import oracle as orc

cg.init()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 40
t0 = time.time()
for case in range(ncase):
    n = int(rng.choice([rng.integers(20, 400), rng.integers(400, 5000), rng.integers(5000, 40000)]))
    S = int(rng.integers(1, 6))
    rr = list(np.round(rng.uniform(0.3, 1.6, size=S), 3))
    bb = [int(rng.integers(1, 300)) for _ in range(S)]
    budget = 1.5e7  # Python big-integer multiply-adds of the exact oracle per case
    while n * sum(bb) > budget / 2:
        bb = [max(1, b // 2) for b in bb]
    K = int(rng.integers(1, max(2, min(140, int(budget / (n * sum(bb)))))))
    seed = int(rng.integers(1, 2 ** 31))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = base[None, :] + 0.106 * rng.standard_normal((K, n))
    _, bn, rr1 = orc.scale_setup(L, rr)
    if min(bn) < 1:
        continue
    opts = {"chunk_log2": int(rng.choice([-1, 12, 14, 17, 20])), "stage": int(rng.choice([-1, 0, 1])),
            "i8_fuse": int(rng.choice([-1, 0])), "draw_fuse": int(rng.choice([-1, 1])),
            "chunk_fill": int(rng.choice([-1, 0, 2, 2]))}
    for k, v in opts.items():
        cg.set_option(k, v)
    world = int(rng.integers(1, 4))
    rank = int(rng.integers(0, world))
    cg.set_shard(0, 1)
    ref = orc.makermt(L, bn, bb, seed)
    strict = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_STRICT)
    i8 = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_INT8)
    for s in range(S):
        assert np.array_equal(strict[s], ref[s]), ("strict", case, K, n, s, opts)
        counts = cg.resample_counts(n, bn, bb, seed, s)
        want = orc.exact_rounded_contract(L, counts, rr1[s])
        assert np.array_equal(i8[s], want), ("int8", case, K, n, s, opts)
    if world > 1:
        cg.set_shard(rank, world)
        loc8 = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_INT8)
        locs = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_STRICT)
        for s in range(S):
            b0, b1 = cg.shard_range(bb[s])
            assert np.array_equal(loc8[s], i8[s][:, b0:b1]), ("shard int8", case, rank, world, s, opts)
            assert np.array_equal(locs[s], strict[s][:, b0:b1]), ("shard strict", case, rank, world, s, opts)
        cg.set_shard(0, 1)
    print(f"case {case}: K={K} n={n} S={S} bb={bb} world={world} {opts} ok", flush=True)
print(f"{ncase} random cases passed in {time.time()-t0:.1f} s")
