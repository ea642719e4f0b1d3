"""CPU-side checks of the drop-in boundary: the shared library loads, exports every
symbol include/consel_gpu.h declares, and refuses to compute without a GPU (there is
no CPU fallback in the product)."""
import ctypes as C

import numpy as np
import pytest

import consel_b200 as cg


def test_library_exports_every_declared_symbol():
    lib = cg.load()
    names = cg.exported_symbols()
    assert len(names) >= 20 and "cgpu_scaleboot_all" in names and "cgpu_count_query" in names
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/consel_gpu.h but not exported"
    assert lib.cgpu_version() >= 100


def test_product_does_not_reach_into_the_oracle():
    # the oracle is test infrastructure: nothing under consel_b200/ or include/ may name it
    import re
    from pathlib import Path
    root = Path(cg.__file__).resolve().parent
    for p in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cpp")) + \
            list(root.rglob("*.h")) + list(root.rglob("*.c")) + list((root.parent / "include").glob("*.h")):
        if "build" in p.parts:
            continue
        txt = p.read_text(errors="ignore")
        assert not re.search(r"(from|import)\s+oracle|liboracle|oracle/_ref|orc_[a-z]+\(", txt), p


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback_without_a_gpu():
    with pytest.raises(cg.CgpuError) as e:
        cg.init()
    assert e.value.code == -2
    L = np.zeros((3, 8))
    with pytest.raises(cg.CgpuError):
        cg.scaleboot_all(L, [4], [2], 1)
    with pytest.raises(cg.CgpuError):
        cg.Reps.upload([np.zeros((3, 4))])


def test_scale_setup_matches_reference_rounding(orc, gold):
    # bn = (int)(rr*n), rr1 = bn/n, datvec sequential sums (src/makermt.c:159-168)
    L = orc.read_lnf(gold / "mam15.lnf")
    rr = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
    ref = orc.read_rmt(gold / "g100.rmt")
    datvec, bn, rr1 = cg.scale_setup(L, rr)
    assert np.array_equal(datvec, ref["datvec"]) and np.array_equal(rr1, ref["rr"])
    assert bn.tolist() == [1707, 2048, 2389, 2731, 3072, 3414, 3755, 4096, 4438, 4779]


def test_smooth_count_helper_is_cntdist3(orc):
    lib = cg.load()
    rng = np.random.default_rng(0)
    for B in (2, 3, 10, 101):
        v = np.sort(rng.standard_normal(B))
        v[B // 2:] = np.maximum(v[B // 2:], v[B // 2])  # some ties
        for t in list(v) + [v[0] - 1, v[-1] + 1, 0.0, (v[0] + v[1]) / 2]:
            le = v[v <= t]
            gt = v[v > t]
            q = np.zeros((), dtype=cg.api.ROWSUM_DTYPE)
            q["cnt"] = len(le)
            q["lo"] = le.max() if len(le) else -1e30
            q["hi"] = gt.min() if len(gt) else 1e30
            q["min1"], q["min2"], q["max1"], q["max2"] = v[0], v[1], v[-1], v[-2]
            assert cg.smooth_count(q, B, float(t)) == orc.cntdist(v, float(t), 3)
