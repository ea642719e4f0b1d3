import sys, os, numpy as np, torch
sys.path.insert(0, ".")
import consel_b200 as cg
K, n, B = 105, 3414, 100000
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
cg.init()
_, bn, _ = cg.scale_setup(L, [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4])
Lh = torch.empty((K, n), dtype=torch.float64, pin_memory=True); Lh.numpy()[:] = L
outs = [torch.empty((K, B), dtype=torch.float64, pin_memory=True).numpy() for _ in range(10)]
for it in range(3):
    if it == 2: print("---- traced call", file=sys.stderr, flush=True)
    cg.scaleboot_all(Lh.numpy(), bn, [B] * 10, 7, sum_mode=cg.SUM_INT8, out=outs)
