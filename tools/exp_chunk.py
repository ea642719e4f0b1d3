import sys, numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
rng = np.random.Generator(np.random.PCG64(1))
K, n, B = 105, 3414, 100000
L = -2.3 - 4.07 * rng.exponential(size=n)[None, :] + 0.1 * rng.standard_normal((K, n))
cg.init()
m = cg.Matrix(L)
_, bn, _ = cg.scale_setup(L, [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4])
for log2 in (-1, 20, 21, 22, 23):
    cg.set_option("chunk_log2", log2)
    for it in range(3):
        r = m.scaleboot(bn, [B] * 10, 7, rng_mode=0, sum_mode=2); t = cg.last_timing(); r.free()
    print(f"chunk 2^{log2}: rng {t['rng']:.2f} ms total {t['total']:.2f}", flush=True)
