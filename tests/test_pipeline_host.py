"""Host-side flow of consel_b200.pipeline.run without a GPU: the row sums are formed on host threads WHILE the
(stand-in) device call runs and must reach consel() as the reference's sequential sums, with bn / rr1 as
src/makermt.c:165-168 computes them; an error on either side surfaces on the calling thread."""
import time

import numpy as np
import pytest

import consel_b200 as cg
from consel_b200 import api, pipeline


class _Reps:
    freed = False

    def free(self):
        self.freed = True


class _Mat:
    def __init__(self, L=None, fail=False):
        self.fail, self.args, self.freed = fail, None, False

    def scaleboot(self, bn, bb, seed, stream_offset, rng_mode, sum_mode):
        time.sleep(0.05)  # the device works; the host thread sums meanwhile
        if self.fail:
            raise api.CgpuError(-7, "device said no")
        self.args = (np.asarray(bn).copy(), np.asarray(bb).copy(), seed, stream_offset, rng_mode, sum_mode)
        return _Reps()

    def free(self):
        self.freed = True


def test_run_hands_consel_the_sequential_row_sums(monkeypatch):
    rng = np.random.default_rng(3)
    L = -rng.exponential(4.0, size=(23, 1000)) - 1.0
    rr, bb = [0.5, 0.77, 1.0, 1.4], [10, 20, 30, 40]
    want_datvec, want_bn, want_rr1 = cg.scale_setup(L, rr)
    seen = {}

    def consel(reps, datvec, rr1, bb_, **kw):
        seen.update(reps=reps, datvec=datvec.copy(), rr1=np.asarray(rr1).copy(), kw=kw)
        return dict(pv=np.zeros((23, 8)), timing={})
    monkeypatch.setattr(pipeline, "consel", consel)
    monkeypatch.setattr(api, "last_timing", lambda: {"total": 1.0})
    mat = _Mat()
    res = pipeline.run(L, rr, bb, 7, sum_mode=api.SUM_INT8, matrix=mat, nosort=True)
    assert np.array_equal(seen["datvec"], want_datvec) and np.array_equal(res["datvec"], want_datvec)
    assert np.array_equal(seen["rr1"], want_rr1) and np.array_equal(res["rr"], want_rr1)
    assert np.array_equal(mat.args[0], want_bn) and mat.args[2:] == (7, 0, api.RNG_MT19937, api.SUM_INT8)
    assert seen["kw"] == {"nosort": True} and seen["reps"].freed and not mat.freed  # a caller's matrix is not freed
    assert res["timing"]["makermt_s"] >= 0.05
    # without a resident matrix run() uploads L itself and frees the handle
    made = []
    monkeypatch.setattr(api, "Matrix", lambda L_: made.append(_Mat()) or made[-1])
    pipeline.run(L, rr, bb, 7)
    assert len(made) == 1 and made[0].freed


def test_run_surfaces_a_device_error_after_joining_the_host_thread(monkeypatch):
    L = -np.ones((8, 64))
    monkeypatch.setattr(pipeline, "consel", lambda *a, **k: pytest.fail("consel must not run"))
    with pytest.raises(api.CgpuError, match="device said no"):
        pipeline.run(L, [1.0], [4], 1, matrix=_Mat(fail=True))
