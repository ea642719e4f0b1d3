"""makermt+consel in-process wall time (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline

def synth(K, n, seed=20261019):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    return base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n

cg.init()
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
cfgs = {"C1": (15, 3414, 10000), "C2": (105, 3414, 100000), "C3": (1000, 100000, 10000), "C3s": (1000, 100000, 2000), "C5": (10000, 50000, 10000), "C5s": (10000, 50000, 1000)}
for name in sys.argv[1:]:
    K, n, B = cfgs[name]
    L = synth(K, n)
    mat = cg.Matrix(L)
    for sm in (2,):
        for it in range(int(__import__("os").environ.get("PIPE_ITERS", "2"))):
            t = time.perf_counter()
            res = pipeline.run(L, RR, [B] * 10, 7, sum_mode=sm, matrix=mat)
            wall = time.perf_counter() - t
            print("   iter", it, f"{wall*1e3:.1f} ms", flush=True)
        print(name, "sum_mode", sm, f"wall {wall*1e3:.1f} ms", {k: (round(v*1e3, 1) if isinstance(v, float) else v) for k, v in res["timing"].items() if k != "device_ms"},
              "dev", {k: round(v, 1) for k, v in res["timing"]["device_ms"].items() if v > 0.05}, flush=True)
        print("   au[:3]", res["pv"][:3, 6], "bp[:3]", res["pv"][:3, 0])
    mat.free()
