"""One small makermt-side pass for ncu captures (development aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg

K = int(sys.argv[1]) if len(sys.argv) > 1 else 105
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3414
B = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
rng_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
cg.init()
m = cg.Matrix(L)
_, bn, _ = cg.scale_setup(L, [0.5, 1.0, 1.4])
modes = [int(x) for x in sys.argv[5].split(',')] if len(sys.argv) > 5 else [0, 1]
for sm in modes:
    for it in range(2):
        r = m.scaleboot(bn, [B] * 3, 7, rng_mode=rng_mode, sum_mode=sm)
        t = cg.last_timing()
        r.free()
    print("sum_mode", sm, t)
