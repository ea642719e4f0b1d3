"""Wall time of the command-line path on a synthetic C2-shaped input (development aid)."""
import subprocess, sys, time, os, tempfile
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as orc
K, n, B = 105, 3414, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
td = tempfile.mkdtemp(prefix="cli_")
orc.write_mt(os.path.join(td, "c2.mt"), L)
orc.write_pa(os.path.join(td, "c2.pa"), [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4], [B] * 10)
bin_ = os.path.abspath("consel_b200/bin")
for cmd in ([f"{bin_}/makermt", "-s", "7", "-p", "c2", "--sum", "int8", "c2"], [f"{bin_}/consel", "c2"], [f"{bin_}/catpv", "c2"]):
    t = time.perf_counter()
    r = subprocess.run(cmd, cwd=td, capture_output=True, text=True)
    dt = time.perf_counter() - t
    tail = [l for l in r.stdout.splitlines() if l.strip()][-3:]
    print(os.path.basename(cmd[0]), f"rc={r.returncode} {dt*1e3:.0f} ms", "|", " / ".join(tail)[:200], flush=True)
print("files:", {f: os.path.getsize(os.path.join(td, f)) for f in sorted(os.listdir(td))})
