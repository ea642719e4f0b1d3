"""CGPU_TRACE=1 python tools/trace_c3.py: span timeline of several C3 resampling calls (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg

rng = np.random.Generator(np.random.PCG64(20261019))
K, n, B = 1000, 100000, 10000
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
cg.init()
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
_, bn, _ = cg.scale_setup(L, RR)
m = cg.Matrix(L)
for it in range(6):
    t = time.perf_counter()
    r = m.scaleboot(bn, [B] * 10, 7, rng_mode=0, sum_mode=2)
    wall = time.perf_counter() - t
    tm = cg.last_timing()
    print(f"iter {it}: wall {wall*1e3:.1f} ms total {tm['total']:.1f} sum {tm['rng']+tm['resample']+tm['gemm_i8']:.1f}", flush=True)
    sys.stderr.flush()
    r.free()
