"""Host side of the MT19937 jump-ahead (consel_b200/csrc/gf2_mt.cpp): the
characteristic polynomial and x^J mod phi, checked against plain sequential
generation by the oracle.  Runs on CPU; the device kernels apply the same
polynomials (tests/test_gpu_makermt.py)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

import consel_b200 as cg


def _window(fn, seed, J):
    lib = cg.load()
    out = np.empty(624, np.uint32)
    getattr(lib, fn)(C.c_uint(seed), C.c_ulonglong(J), out.ctypes.data_as(C.POINTER(C.c_uint)))
    return out


def _raw_sequence(seed, count):
    x = [0] * count
    x[0] = seed & 0xFFFFFFFF
    for j in range(1, 624):
        x[j] = (1812433253 * (x[j - 1] ^ (x[j - 1] >> 30)) + j) & 0xFFFFFFFF
    for t in range(count - 624):
        y = (x[t] & 0x80000000) | (x[t + 1] & 0x7FFFFFFF)
        x[t + 624] = x[t + 397] ^ (y >> 1) ^ (0x9908B0DF if (x[t + 1] & 1) else 0)
    return np.array(x, dtype=np.uint32)


@pytest.mark.parametrize("seed,J", [(1234, 0), (1234, 1), (1234, 623), (1234, 624), (5, 19937), (5, 20000),
                                     (77, 123457)])
def test_polynomial_jump_equals_sequential_words(seed, J):
    seq = _raw_sequence(seed, J + 624)
    for fn in ("cgpu_test_mt_window_host", "cgpu_test_mt_window_table"):
        w = _window(fn, seed, J)
        # the 31 low bits of the first word are not part of the generator state
        assert np.array_equal(w[1:], seq[J + 1:J + 624]) and (w[0] >> 31) == (seq[J] >> 31), (fn, J)


@pytest.mark.parametrize("seed,blocks", [(1234, 1000), (99, 40_000)])
def test_block_aligned_jump_equals_oracle_state(orc, seed, blocks):
    ref = orc.mt_state_after_blocks(seed, blocks)
    w = _window("cgpu_test_mt_window_table", seed, 624 * blocks)
    assert np.array_equal(w[1:], ref[1:]) and (w[0] >> 31) == (ref[0] >> 31)


def test_characteristic_polynomial_rederived():
    sys.path.insert(0, str(Path(cg.__file__).resolve().parent / "csrc"))
    import gen_mt_phi
    exps = gen_mt_phi.charpoly_exponents(seed=20261019, plane=17)
    inc = (Path(cg.__file__).resolve().parent / "csrc" / "mt_phi.inc").read_text()
    body = " ".join(ln for ln in inc.splitlines() if not ln.startswith("//"))
    stored = [int(t) for t in body.replace(",", " ").split()]
    assert exps == stored and len(exps) == cg.load().cgpu_test_phi_terms() == 135 and exps[-1] == 19937
