"""catmt (host file tool on makermt's input side): site-wise concatenation of
.mt files and the --kht screening — output files byte-equal and stdout equal to the unmodified
reference's catmt (src/catmt.c:63-201) compiled under oracle/_ref.  Host-only."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "consel_b200" / "bin" / "catmt"

CASES = [
    (["ga.mt", "gb.mt"], []),
    (["ga.mt", "gb.mt", "gc.mt"], ["--kht", "0.3"]),   # screening: writes out.vt too
    (["tiny.mt"], ["--kht", "0"]),                     # table only
    (["ga.mt"], ["--kht", "0.05"]),
    (["ga.mt", "odd.mt"], []),                         # row mismatch -> "#! Terminated"
]


def body(out: str) -> str:
    return out.split("\n", 1)[1] if "\n" in out else ""


def run(exe, args, cwd):
    return subprocess.run([str(exe)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)


@pytest.mark.parametrize("files,flags", CASES)
def test_catmt_equals_reference_binary(files, flags, gold, orc, tmp_path):
    if orc.ref_bin("catmt") is None:
        pytest.skip("oracle/_ref/catmt not built")
    (tmp_path / "odd.mt").write_text("2 3\n-1 -2 -3\n-2 -1 -3\n")
    ins = [tmp_path / f if f == "odd.mt" else gold / f for f in files]
    for d in "ab":
        (tmp_path / d).mkdir()
    ours = run(EXE, ins + flags + ["out"], tmp_path / "a")
    ref = run(orc.ref_bin("catmt"), ins + flags + ["out"], tmp_path / "b")
    assert ours.returncode == ref.returncode
    assert body(ours.stdout) == body(ref.stdout)
    for ext in (".mt", ".vt"):
        fa, fb = tmp_path / "a" / ("out" + ext), tmp_path / "b" / ("out" + ext)
        assert fa.exists() == fb.exists()
        if fa.exists():
            assert fa.read_bytes() == fb.read_bytes()


def test_catmt_concatenates_sites(gold, orc, tmp_path):
    r = run(EXE, [gold / "ga.mt", gold / "gb.mt", "ab"], tmp_path)
    assert r.returncode == 0 and "# exit normally" in r.stdout
    a, b, ab = (orc.read_mt(p) for p in (gold / "ga.mt", gold / "gb.mt", tmp_path / "ab.mt"))
    assert ab.shape == (a.shape[0], a.shape[1] + b.shape[1])
    assert np.allclose(np.concatenate([a, b], axis=1), ab, rtol=1e-8, atol=0)  # ascii .mt: 10 printed digits, like the reference
    r = run(EXE, [gold / "ga.mt"], tmp_path)
    assert r.returncode == 1 and "must specify input-file and output-file" in r.stdout
