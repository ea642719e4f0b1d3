#!/usr/bin/env python
"""Input path: parse time of a ~1 GB `.mt` and a ~1 GB PAML `.lnf` by the product's readers
(consel_b200/bin/parse_bench -> refio::read_mat / read_mat_paml) and by the UNMODIFIED reference
(oracle/_ref/makermt with one replicate per scale: its wall time minus the bootstrap time it prints).
The values parsed are compared bit for bit with the ones written (fnv1a over the doubles).

    python tools/parse_time.py [GB=1.0] [workdir=/tmp/parse_time]
"""
import json
import re
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GB = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
WORK = Path(sys.argv[2] if len(sys.argv) > 2 else "/tmp/parse_time")
WORK.mkdir(parents=True, exist_ok=True)


def c_fnv():
    import ctypes
    src = WORK / "fnv.c"
    so = WORK / "fnv.so"
    src.write_text("#include <stdint.h>\n#include <stddef.h>\nuint64_t fnv(const uint64_t*a,size_t n){uint64_t h=1469598103934665603ull;"
                   "for(size_t i=0;i<n;i++)h=(h^a[i])*1099511628211ull;return h;}\n")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.fnv.restype = ctypes.c_uint64
    lib.fnv.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    return lambda a: f"{lib.fnv(a.ctypes.data, a.size):016x}"


def write_mt(path: Path, L: np.ndarray):
    """the .mt text format (src/misc.c:385-404): `m n` then the values; 17 significant digits round-trip"""
    with open(path, "w") as f:
        f.write(f"{L.shape[0]} {L.shape[1]}\n")
        for row in L:
            np.savetxt(f, row.reshape(-1, 8), fmt="%.17g")


def write_lnf(path: Path, L: np.ndarray):
    """PAML .lnf (src/freadmat.c:110-154): `m n p`, per tree an id line + p pattern lines `idx count lnL ...`"""
    m, n = L.shape
    with open(path, "w") as f:
        f.write(f"{m} {n} {n}\n\n")
        idx = np.arange(1, n + 1)
        for i in range(m):
            f.write(f"{i + 1}\n\n")
            np.savetxt(f, np.column_stack([idx, np.ones(n), L[i], np.exp(L[i])]), fmt=["%6d", "%6d", "%.17g", "%.6g"])
            f.write("\n")


def run_ours(fmt, path):
    out = subprocess.run([str(ROOT / "consel_b200" / "bin" / "parse_bench"), fmt, str(path), "2"], capture_output=True, text=True, check=True)
    rows = [json.loads(t) for t in out.stdout.splitlines() if t.startswith("{")]
    return min(rows, key=lambda r: r["seconds"])


def run_reference(fmt, path):
    ref = ROOT / "oracle" / "_ref" / "makermt"
    if not ref.exists():
        return None
    stem = path.with_suffix("")
    cmd = [str(ref), "-s", "1", "-b", "0.0001"] + (["--paml"] if fmt == "paml" else []) + [str(stem), str(WORK / "ref_out")]
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    mt = re.search(r"bootstrap t=([0-9.eE+-]+) sec", out.stdout)
    if out.returncode != 0 or not mt:
        return {"error": out.stdout[-300:]}
    return {"wall_s": wall, "bootstrap_s": float(mt.group(1)), "parse_s": wall - float(mt.group(1))}


def main():
    h = c_fnv()
    n = 100000
    rng = np.random.Generator(np.random.PCG64(5))
    report = []
    for fmt, per_value in (("mt", 24.0), ("paml", 50.0)):
        m = max(2, int(GB * 1e9 / per_value / n))
        L = -2.3 - 4.07 * rng.exponential(size=n)[None, :] + 0.106 * rng.standard_normal((m, n))
        path = WORK / ("big.mt" if fmt == "mt" else "big.lnf")
        (write_mt if fmt == "mt" else write_lnf)(path, L)
        ours = run_ours(fmt, path)
        ours["values_identical_to_those_written"] = ours["fnv1a_of_values"] == h(np.ascontiguousarray(L))
        ref = run_reference(fmt, path)
        row = {"format": fmt, "file_GB": path.stat().st_size / 1e9, "trees": m, "sites": n, "ours": ours, "reference_makermt": ref}
        if ref and "parse_s" in ref:
            row["speedup"] = ref["parse_s"] / ours["seconds"]
        print(json.dumps(row), flush=True)
        report.append(row)
        path.unlink()
    return 0


if __name__ == "__main__":
    sys.exit(main())
