#!/usr/bin/env python
"""The sharded makermt -> consel pipeline on the EMULATED library at world sizes the CPU suite does not hold
(tests/test_emu_dist.py runs 2 and 3): every rank of a 4- and an 8-process run returns the single-process table
bit for bit -- item ranges with fewer items than ranks * 2, reduce-scatter padding, the collective AU rounds.

    python tools/emu_dist_worlds.py [worlds=4,8] [builddir=/tmp/emu_dist_worlds] [SEED COUNT: random shapes / associations instead]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "emu"):
    sys.path.insert(0, str(p))
import build_emu_lib  # noqa: E402
import test_emu_dist as t  # noqa: E402


def main() -> int:
    worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4,8").split(",")]
    build = Path(sys.argv[2] if len(sys.argv) > 2 else "/tmp/emu_dist_worlds")
    lib = build_emu_lib.build(build)
    out = build / "out"
    out.mkdir(exist_ok=True)
    modes = ("identity", "edges")
    if len(sys.argv) > 3:  # random cases: python tools/emu_dist_worlds.py 2,3,5 /tmp/x SEED COUNT
        seed, count = int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 5
        modes = tuple(f"rand{seed * 1000 + c}" for c in range(count))
    for mode in modes:
        try:
            one = t._run_ranks(lib, out / f"one_{mode}", 1, mode)[0]
        except AssertionError as e:
            if "#! Terminated" not in str(e):
                raise
            print(f"{mode}: skipped (the fit ends in the reference's error() for this input)", flush=True)
            continue
        for world in worlds:
            many = t._run_ranks(lib, out / f"w{world}_{mode}", world, mode)
            for rank, got in enumerate(many):
                for key in ("item", "stat", "pv", "se", "cnt", "ax", "ci", "ret", "datvec"):
                    assert np.array_equal(got[key], one[key], equal_nan=True), (mode, world, rank, key)
            print(f"{mode}: world {world} identical to one process", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
