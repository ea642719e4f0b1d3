set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_h.log 2>&1; tail -15 gpurun_out/r02_gputests_h.log
python - <<'PY' > gpurun_out/r02_pw.log 2>&1
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline
cg.init()
rng = np.random.Generator(np.random.PCG64(20261019))
K, n, B = 10000, 50000, 2000
base = -2.3 - 4.07 * rng.exponential(size=n)
L = np.empty((K, n))
for k0 in range(0, K, 100):
    L[k0:k0+100] = base[None, :] + 0.106 * rng.standard_normal((100, n)) - 0.5 * np.arange(k0, k0+100)[:, None] / n
m = cg.Matrix(L)
cg.set_option("ktime", 1)
res = pipeline.run(L, [0.9, 1.0], [B, B], 7, sum_mode=2, matrix=m, do_au=False)
for k, v in cg.kernel_times().items(): print(k, v)
PY
cat gpurun_out/r02_pw.log
