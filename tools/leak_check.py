import sys, numpy as np, torch
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline
cg.init()
rng = np.random.Generator(np.random.PCG64(3))
K, n, B = 30, 2000, 3000
L = -2.3 - 4.07 * rng.exponential(size=n)[None, :] + 0.1 * rng.standard_normal((K, n))
RR = [0.5, 0.7, 0.9, 1.0, 1.1, 1.3]
def used():
    f, t = torch.cuda.mem_get_info(); return (t - f) / 1e6
for it in range(61):
    res = pipeline.run(L, RR, [B] * 6, 7 + it % 3, sum_mode=[0, 2][it % 2])
    outs = cg.scaleboot_all(L, [int(r * n) for r in RR], [B] * 6, 7, sum_mode=2)
    if it % 20 == 0: print(it, f"{used():.0f} MB in use", flush=True)
