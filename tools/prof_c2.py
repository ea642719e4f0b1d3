"""Full C2 makermt-side pass (10 scales x 100000) for ncu launch lists (development aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg

rng_mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sum_mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps_n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
K, n, B = 105, 3414, 100000
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
cg.init()
m = cg.Matrix(L)
_, bn, _ = cg.scale_setup(L, [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4])
for it in range(reps_n):
    r = m.scaleboot(bn, [B] * 10, 7, rng_mode=rng_mode, sum_mode=sum_mode)
    t = cg.last_timing()
    r.free()
print(t)
