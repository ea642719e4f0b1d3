"""The `consel` COMMAND LINE on the CPU: the host program's own sources linked against the emulated device
library (conftest.emu_bin), run on the reference's golden inputs and compared with the files the unmodified
reference wrote -- `.rep` / `.cnt` byte for byte, `.pv` / `.ci` to the printed digits, `catpv` text equal.
The consel-only tests of tests/test_gpu_cli.py run unchanged (edge associations, -R rep mode, -f rescaling,
--quick_mctest, --pdm, the error protocol); makermt needs the GPU (MT19937 stream + contraction) and is not
part of this tier."""
import shutil
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))

import test_gpu_cli as _cli  # noqa: E402

for _name in ("test_consel_with_edge_associations", "test_consel_rep_mode", "test_consel_rescaling_approximation",
              "test_quick_mctest", "test_pdm_aic_correction", "test_cli_error_protocol"):
    globals()[_name] = getattr(_cli, _name)


@pytest.fixture(autouse=True)
def _emulated_programs(emu_bin, monkeypatch):
    monkeypatch.setattr(_cli, "BIN", emu_bin)


@pytest.mark.parametrize("case", ["g100", "tiny"])
def test_consel_catpv_on_the_reference_rmt(case, gold, orc, tmp_path):
    """the consel -> catpv half of tests/test_gpu_cli.py::test_makermt_consel_catpv_pipeline, from the .rmt the
    reference's makermt wrote"""
    shutil.copy(gold / f"{case}.rmt", tmp_path / f"{case}.rmt")
    out = _cli.run("consel", ["-r", "-c", case], tmp_path)
    assert "# exit normally" in out
    assert (tmp_path / f"{case}.rep").read_bytes() == (gold / f"{case}.rep").read_bytes()
    assert (tmp_path / f"{case}.cnt").read_bytes() == (gold / f"{case}.cnt").read_bytes()
    for ext in ("pv", "ci"):
        _cli.same_numbers(tmp_path / f"{case}.{ext}", gold / f"{case}.{ext}", orc)
    assert _cli.run("catpv", [case], tmp_path) == (gold / f"{case}.txt").read_text()
