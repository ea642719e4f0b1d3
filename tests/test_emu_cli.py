"""The makermt / makerep / consel COMMAND LINES on the CPU: the host programs' own sources linked against the
emulated device library (conftest.emu_bin), run on the reference's inputs and compared with the files the
unmodified reference wrote -- `.rmt` / `.vt` / `.rep` / `.cnt` byte for byte, `.pv` / `.ci` to the printed
digits, `catpv` text equal -- and against a fresh run of the compiled reference.  These are the tests of
tests/test_gpu_cli.py, unchanged and all of them (the --sum int8 runs with the MMA instruction replaced by a plain
int32 product, see tests/emu/shim/i8_gemm_emu.h)."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))

import test_gpu_cli as _cli  # noqa: E402

for _name in dir(_cli):
    if _name.startswith("test_"):
        globals()[_name] = getattr(_cli, _name)


@pytest.fixture(autouse=True)
def _emulated_programs(emu_bin, emu_cg, monkeypatch):
    monkeypatch.setattr(_cli, "BIN", emu_bin)
