"""catrep (host file tool): merging .rep / .rmt files — same bytes and the same
stdout as the unmodified reference's catrep (src/catrep.c:61-285) compiled under oracle/_ref,
plus properties that need no reference binary.  Host-only: no GPU involved."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "consel_b200" / "bin" / "catrep"

CASES = [
    (["g100.rep", "g100f.rep"], []),                       # disjoint + equal scales, sorted
    (["g100.rep", "g100.rep"], ["--no_sort"]),             # every scale doubled, file order kept
    (["g100.rep", "g100.rep"], ["-p", "fscale.pa"]),       # selection: only r=1.0 survives, warnings
    (["g100.rmt", "g100.rmt"], ["-m"]),                    # .rmt mode (no id vector)
    (["tiny.rmt"], ["-m", "-a"]),                          # ascii output
    (["g100.rep", "g100e.rep"], []),                       # cm mismatch -> "#! Terminated"
]


def run_ours(args, cwd):
    return subprocess.run([str(EXE)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)


def body(out: str) -> str:  # the first line is the program's own identification
    return out.split("\n", 1)[1] if "\n" in out else ""


@pytest.mark.parametrize("files,flags", CASES)
def test_catrep_equals_reference_binary(files, flags, gold, orc, tmp_path):
    if orc.ref_bin("catrep") is None:
        pytest.skip("oracle/_ref/catrep not built")
    ext = ".rmt" if "-m" in flags else ".rep"
    fl = [str(gold / f) if f.endswith(".pa") else f for f in flags]
    ins = [gold / f for f in files]
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    ours = run_ours(fl + ins + ["out"], tmp_path / "a")
    ref = subprocess.run([str(orc.ref_bin("catrep"))] + fl + [str(p) for p in ins] + ["out"],
                         cwd=tmp_path / "b", capture_output=True, text=True)
    assert ours.returncode == ref.returncode
    assert body(ours.stdout) == body(ref.stdout)
    fa, fb = tmp_path / "a" / ("out" + ext), tmp_path / "b" / ("out" + ext)
    assert fa.exists() == fb.exists()
    if fa.exists():
        assert fa.read_bytes() == fb.read_bytes()


def test_catrep_merge_properties(gold, orc, tmp_path):
    """Without the reference binary: merging a file with itself doubles every block, keeps
    ids / observations / scales, and every row comes out ascending."""
    r = run_ours([gold / "g100.rep", gold / "g100.rep", "twice"], tmp_path)
    assert r.returncode == 0 and "# exit normally" in r.stdout
    a, b = orc.read_rep(gold / "g100.rep"), orc.read_rep(tmp_path / "twice.rep")
    assert np.array_equal(a["order"], b["order"]) and np.array_equal(a["obs"], b["obs"])
    assert np.array_equal(a["rr"], b["rr"]) and np.array_equal(2 * np.asarray(a["bb"]), b["bb"])
    for x, y in zip(a["stats"], b["stats"]):
        x, y = np.asarray(x), np.asarray(y)
        assert np.array_equal(np.sort(np.concatenate([x, x], axis=1), axis=1), y)


def test_catrep_command_line_errors(tmp_path, gold):
    # -n (headers only): the reference dereferences blocks it never read and crashes; ours
    # prints the headers and writes zero blocks
    r = run_ours(["-n", gold / "g100.rep", "hdr"], tmp_path)
    assert r.returncode == 0 and " --- skipped." in r.stdout and "# CM:15" in r.stdout
    r = run_ours([gold / "g100.rep"], tmp_path)
    assert r.returncode == 1 and "must specify input-file and output-file" in r.stdout
    r = run_ours(["--bogus", gold / "g100.rep", "x"], tmp_path)
    assert r.returncode == 1 and "error in command line" in r.stdout
