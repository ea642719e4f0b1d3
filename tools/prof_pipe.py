"""C2 makermt+consel in one process, once warm + once more (for ncu launch lists; development aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline
K, n, B = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (105, 3414, 100000)
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
cg.init()
mat = cg.Matrix(L)
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
for it in range(2):
    res = pipeline.run(L, RR, [B] * 10, 7, sum_mode=2, matrix=mat)
print({k: round(v * 1e3, 1) for k, v in res["timing"].items() if isinstance(v, float)})
