"""Ad-hoc timing of the makermt-side pipeline (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
import ctypes as C

def synth(K, n, seed=20261019):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    return base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n

cg.init()
tf = C.c_double()
cg.load().cgpu_bench_fp64_peak(C.byref(tf))
print("fp64 DFMA peak TFLOP/s:", tf.value, flush=True)
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
cfgs = [("C1", 15, 3414, 10000), ("C2", 105, 3414, 100000), ("C3/10", 1000, 100000, 1000), ("C3", 1000, 100000, 10000), ("K300", 300, 3414, 20000)]
cfgs.append(("C5", 10000, 50000, 10000))
if len(sys.argv) > 1:
    cfgs = [c for c in cfgs if c[0] in sys.argv[1:]]
import os
MODES = [(0, 0), (1, 1), (0, 2), (1, 2)] if not os.environ.get("I8_ONLY") else [(0, 2), (1, 2)]
for name, K, n, B in cfgs:
    L = synth(K, n)
    _, bn, _ = cg.scale_setup(L, RR)
    bb = [B] * 10
    m = cg.Matrix(L)
    acc = K * n * sum(bb)
    for rng, sm in MODES:
        for it in range(2 if K * n * B > 5e11 else 3):
            t = time.time()
            r = m.scaleboot(bn, bb, 7, rng_mode=rng, sum_mode=sm)
            wall = time.time() - t
            tm = cg.last_timing()
            r.free()
        print(f"{name} K={K} n={n} B={B} rng={rng} sum={sm}: wall {wall*1e3:.2f} ms  dev {tm}  acc/s {acc/(tm['total']*1e-3):.3e}", flush=True)
    m.free()
