"""The makermt / makerep / consel COMMAND LINES on the CPU: the host programs' own sources linked against the
emulated device library (conftest.emu_bin), run on the reference's inputs and compared with the files the
unmodified reference wrote -- `.rmt` / `.vt` / `.rep` / `.cnt` byte for byte, `.pv` / `.ci` to the printed
digits, `catpv` text equal -- and against a fresh run of the compiled reference.  These are the tests of
tests/test_gpu_cli.py, unchanged; the four that use the int8 tensor-core mode stay GPU-only."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))

import test_gpu_cli as _cli  # noqa: E402

GPU_ONLY = {"test_makerep_matches_reference", "test_in_process_pipeline_matches_reference_pv", "test_makermt_extension_flags",
            "test_full_size_mam15_pipeline_and_philox_within_monte_carlo_error"}  # end with / consist of --sum int8 runs
for _name in dir(_cli):
    if _name.startswith("test_") and _name not in GPU_ONLY:
        globals()[_name] = getattr(_cli, _name)


@pytest.fixture(autouse=True)
def _emulated_programs(emu_bin, emu_cg, monkeypatch):
    monkeypatch.setattr(_cli, "BIN", emu_bin)
