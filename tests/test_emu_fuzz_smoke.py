"""A few cases of the two randomised sweeps (tools/emu_stress.py, tools/emu_kernel_stress.py) inside the suite, so
that the tools keep working: the programs on the emulated library against the unmodified reference binaries, and
the CUDA sources executed on the CPU against the oracle.  The long campaigns are run by hand."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, ROOT / "tools" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_programs_against_the_reference_binaries_on_random_inputs(emu_bin, orc, tmp_path):
    if orc.ref_bin("consel") is None or orc.ref_bin("makermt") is None:
        pytest.skip("oracle/_ref not built")
    es = _tool("emu_stress")
    rng = np.random.default_rng(2026)
    for c in range(4):
        d = tmp_path / f"case{c}"
        d.mkdir()
        assert es.one_case(rng, emu_bin, d) == []
    assert es.STATS["both_ok"] >= 4


def test_kernels_against_the_oracle_on_random_shapes(emu_cg):
    ks = _tool("emu_kernel_stress")
    rng = np.random.default_rng(2027)
    for c in range(8):
        ks.makermt_case(rng, c)
        ks.consel_case(rng, c)
