"""A/B of the pair GEMM inside one process (same GPU, same power state): tile order (i8_panel 0 / 4 / 8 / 16),
one-CTA vs cta_group::2; GEMM phase time per step (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg

def synth(K, n, seed=20261019):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = np.empty((K, n))
    for k0 in range(0, K, 100):
        k1 = min(K, k0 + 100)
        L[k0:k1] = base[None, :] + 0.106 * rng.standard_normal((k1 - k0, n)) - 0.5 * np.arange(k0, k1)[:, None] / n
    return L

cg.init()
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
cfgs = {"C3": (1000, 100000, 10000), "C5": (10000, 50000, 10000)}
for name in sys.argv[1:]:
    K, n, B = cfgs[name]
    L = synth(K, n)
    _, bn, _ = cg.scale_setup(L, RR)
    m = cg.Matrix(L)
    m.scaleboot(bn, [B] * 10, 7, rng_mode=1, sum_mode=2).free()
    for rnd in range(2):
        for label, opts in (("2cta panel8", {"i8_2cta": 1, "i8_panel": 8}), ("2cta plain", {"i8_2cta": 1, "i8_panel": 0}),
                            ("2cta panel4", {"i8_2cta": 1, "i8_panel": 4}), ("2cta panel16", {"i8_2cta": 1, "i8_panel": 16}),
                            ("1cta", {"i8_2cta": 0})):
            for k, v in opts.items():
                cg.set_option(k, v)
            ts = []
            for it in range(3):
                r = m.scaleboot(bn, [B] * 10, 7, rng_mode=1, sum_mode=2)
                ts.append(cg.last_timing()["gemm_i8"])
                r.free()
            print(f"{name} round {rnd} {label:14s} gemm ms {min(ts):8.2f} (all {[round(t,1) for t in ts]})", flush=True)
    cg.set_option("i8_2cta", -1); cg.set_option("i8_panel", -1)
    m.free()
