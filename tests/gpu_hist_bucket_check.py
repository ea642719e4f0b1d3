"""GPU check of the bucketed site histogram (cgpu_set_option("hist_bucket", ...), consel_b200/csrc/
hist_bucket.cuh), run as its own process by tests/test_zz_gpu_hist_bucket.py.

The replicate sums of CGPU_SUM_INT8 are exact functions of the site counts, so bit-equal sums against
the default histogram kernels (global atomics above 225280 sites, one-replicate-per-CTA shared
histograms below) and against the correctly rounded exact sums of the ORACLE's counts mean the
bucketed counts are right.  Exit status 0 = everything equal."""
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def synth(K, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    return base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n


CASES_GPU = ((4, 300000, [0.6, 1.3], [6, 5], 1), (3, 1600000, [1.4], [3], 1), (5, 100000, [0.5, 1.2], [300, 7], 2))


def check(cg, orc, cases=CASES_GPU) -> list:
    """(K, n, scales, replicates, option value): 1 = the opt-in path above 225280 sites, 2 = forced below.
    Returns the list of mismatches (empty = the bucketed path returns the bits of the default kernels)."""
    bad = []
    for K, n, rr, bb, opt in cases:
        L = synth(K, n, 11 + K)
        _, bn, rr1 = orc.scale_setup(L, rr)
        for rng_mode in (cg.RNG_MT19937, cg.RNG_PHILOX):
            cg.set_option("hist_bucket", 0)
            base = cg.scaleboot_all(L, bn, bb, 5, rng_mode=rng_mode, sum_mode=cg.SUM_INT8)
            cg.set_option("hist_bucket", opt)
            try:
                got = cg.scaleboot_all(L, bn, bb, 5, rng_mode=rng_mode, sum_mode=cg.SUM_INT8)
            finally:
                cg.set_option("hist_bucket", -1)
            for s in range(len(rr)):
                if not np.array_equal(base[s], got[s]):
                    bad.append(f"n={n} rng={rng_mode} scale {s}: bucketed sums differ from the default kernels")
            if rng_mode == cg.RNG_MT19937:
                nrep = [min(2, b) for b in bb]
                counts = orc.scaleboot_counts(n, bn, bb, 5)
                for s in range(len(rr)):
                    for b in range(nrep[s]):
                        c = counts[s][b].astype(np.int64)
                        want = math.fsum(np.repeat(L[0], c)) / np.float64(rr1[s])
                        if got[s][0, b] != want:
                            bad.append(f"n={n} scale {s} replicate {b}: not the exact sum over the reference's draws")
    return bad


def main() -> int:
    import consel_b200 as cg
    from oracle import oracle as orc

    cg.init(0)
    cg.set_shard(0, 1)
    bad = check(cg, orc)
    for line in bad:
        print("MISMATCH", line)
    print("hist_bucket check:", "ok" if not bad else f"{len(bad)} mismatches")
    return 0 if not bad else 1


if __name__ == "__main__":
    sys.exit(main())
