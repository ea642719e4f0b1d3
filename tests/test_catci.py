"""catci (host table printer for consel's .ci files): stdout equal, byte for
byte, to the unmodified reference's catci (src/catci.c:27-129) compiled under oracle/_ref, on
the .ci files the reference's consel wrote (tests/golden).  Host-only."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "consel_b200" / "bin" / "catci"

CASES = [["g100"], ["g100e", "tiny"], ["-v", "g100R"], ["--no_au", "g100f"], ["--no_np", "-v", "g100fe"],
         ["g100p.ci", "g100eR"], ["--bogus"], ["missing_file"]]


@pytest.mark.parametrize("args", CASES, ids=lambda a: " ".join(a))
def test_catci_equals_reference_binary(args, gold, orc):
    if orc.ref_bin("catci") is None:
        pytest.skip("oracle/_ref/catci not built")
    a = subprocess.run([str(EXE)] + args, cwd=gold, capture_output=True, text=True)
    b = subprocess.run([str(orc.ref_bin("catci"))] + args, cwd=gold, capture_output=True, text=True)
    assert a.returncode == b.returncode
    assert a.stdout == b.stdout


def test_catci_table_shape(gold):
    out = subprocess.run([str(EXE), "g100"], cwd=gold, capture_output=True, text=True).stdout
    lines = out.strip("\n").split("\n")
    assert lines[0] == "# reading g100.ci" and " au " in lines[1] and " np " in lines[1]
    assert lines[2].startswith("# rank item    obs") and len(lines) == 3 + 15
