"""Driver for the round-2 ncu captures (development aid):
    python tools/prof_r02.py C3|C5|C2 makermt|consel [B]
makermt: two scales (r = 0.5, 1.4) of the configuration's matrix, MT19937 + int8, after one warm-up call.
consel : two scales of int8 replicates, then chost_consel_run (KH/SH tests, statistics, AU) on them."""
import sys
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline

cfgs = {"C2": (105, 3414, 100000), "C3": (1000, 100000, 10000), "C5": (10000, 50000, 10000)}
name, what = sys.argv[1], sys.argv[2]
K, n, B = cfgs[name]
if len(sys.argv) > 3:
    B = int(sys.argv[3])
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = np.empty((K, n))
for k0 in range(0, K, 100):
    k1 = min(K, k0 + 100)
    L[k0:k1] = base[None, :] + 0.106 * rng.standard_normal((k1 - k0, n)) - 0.5 * np.arange(k0, k1)[:, None] / n
cg.init()
m = cg.Matrix(L)
if what == "makermt":
    RR = [0.5, 1.4]
    _, bn, _ = cg.scale_setup(L, RR)
    for it in range(2):
        r = m.scaleboot(bn, [B, B], 7, rng_mode=0, sum_mode=2)
        print(cg.last_timing())
        r.free()
else:
    RR = [0.9, 1.0]
    res = pipeline.run(L, RR, [B, B], 7, sum_mode=2, matrix=m)
    print(res["timing"], res["pv"][:2])
