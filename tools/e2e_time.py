"""e2e (host buffers) timing of the C2 makermt path with the phase breakdown (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import consel_b200 as cg
K, n, B = 105, 3414, 100000
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
cg.init()
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
_, bn, _ = cg.scale_setup(L, RR)
bb = [B] * 10
Lh = torch.empty((K, n), dtype=torch.float64, pin_memory=True); Lh.numpy()[:] = L
outs = [torch.empty((K, B), dtype=torch.float64, pin_memory=True).numpy() for _ in bb]
for stage in (0, 1, -1):
    cg.set_option("stage", stage)
    for it in range(4):
        t0 = time.perf_counter()
        cg.scaleboot_all(Lh.numpy(), bn, bb, 7, sum_mode=cg.SUM_INT8, out=outs)
        dt = (time.perf_counter() - t0) * 1e3
    print(f"stage={stage}: wall {dt:.2f} ms", {k: round(v, 2) for k, v in cg.last_timing().items()}, flush=True)
