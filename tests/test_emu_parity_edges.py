"""Edge cases of the consel command line found by the randomised differential run (tools/emu_stress.py): the
product's consel (its own sources on the emulated device library, conftest.emu_bin) against the UNMODIFIED
reference binary on the same .rmt, file by file.

  * no AU test (a single scale, src/consel.c:868; or --no_au): the .cnt file carries the CNT header alone (cntmat
    is allocated by do_bootrep only, :1464 / :1770) and the .rep rows are in whatever state the tests left them --
    do_bptest's pvaldist sorts the rows of scale i0 only (:1298, :1870), nothing is sorted under --quick_mctest (:323)
  * a fit that ends in the reference's error() (`a too large, ITMAX too small in gcf`, raised inside the fit that
    here runs on a pool thread): `#! Terminated: ...` on stdout and exit status 1, not an abort"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def progs(emu_bin, orc):
    if orc.ref_bin("consel") is None or orc.ref_bin("makermt") is None:
        pytest.skip("oracle/_ref not built")
    return emu_bin


def make_rmt(orc, d: Path, seed: int, K: int, n: int, rr, bb, sd: float):
    rng = np.random.default_rng(seed)
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = base[None, :] + sd * rng.standard_normal((K, n))
    orc.write_mt(d / "in.mt", L)
    orc.write_pa(d / "p.pa", rr, bb)
    orc.run_ref("makermt", ["-s", seed, "-p", "p", "in", "x"], d)


def both(progs, orc, d: Path, args, tag):
    r = subprocess.run([str(orc.ref_bin("consel"))] + args + ["x", f"r{tag}"], cwd=d, capture_output=True, text=True)
    o = subprocess.run([str(progs / "consel")] + args + ["x", f"o{tag}"], cwd=d, capture_output=True, text=True)
    return r, o


def same_files(orc, d: Path, tag, binary=("rep", "cnt"), text=("pv", "ci")):
    for ext in binary:
        a, b = d / f"o{tag}.{ext}", d / f"r{tag}.{ext}"
        assert a.exists() == b.exists(), ext
        if b.exists():
            assert a.read_bytes() == b.read_bytes(), ext
    for ext in text:
        a, b = d / f"o{tag}.{ext}", d / f"r{tag}.{ext}"
        assert a.exists() == b.exists(), ext
        if b.exists():
            x, y = orc.read_text_blocks(a), orc.read_text_blocks(b)
            assert len(x) == len(y) and np.allclose(x, y, rtol=2e-9, atol=1e-12, equal_nan=True), ext


@pytest.mark.parametrize("extra", [[], ["--quick_mctest"], ["--smooth_cnt"]])
def test_single_scale_skips_the_au_test_like_the_reference(progs, orc, tmp_path, extra):
    make_rmt(orc, tmp_path, 5, 7, 120, [1.0], [90], 0.106)
    r, o = both(progs, orc, tmp_path, extra + ["-r", "-c"], "a")
    assert r.returncode == 0 and o.returncode == 0, o.stdout[-500:]
    assert "AU-TEST" not in o.stdout
    same_files(orc, tmp_path, "a")
    assert not (tmp_path / "oa.ci").exists()


@pytest.mark.parametrize("extra", [["--no_au"], ["--no_au", "--quick_mctest"], ["--no_au", "--no_bp"], ["--no_au", "--no_sh", "--no_pp"]])
def test_no_au_leaves_rep_and_cnt_as_the_reference_does(progs, orc, tmp_path, extra):
    make_rmt(orc, tmp_path, 9, 6, 200, [0.6, 1.0, 1.4], [70, 80, 60], 0.106)
    r, o = both(progs, orc, tmp_path, extra + ["-r", "-c"], "b")
    assert r.returncode == o.returncode == 0, o.stdout[-500:]
    same_files(orc, tmp_path, "b")


def test_fit_failure_terminates_like_the_reference(progs, orc, tmp_path):
    rng = np.random.default_rng(66)
    K, n = int(rng.integers(4, 10)), int(rng.integers(30, 200))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = base[None, :] + 1.0 * rng.standard_normal((K, n))
    rr = np.sort(np.concatenate([np.round(rng.uniform(0.3, 0.7, 3), 2), np.round(rng.uniform(1.4, 2.0, 5), 2)]))
    bb = rng.integers(20, 120, size=8)
    orc.write_mt(tmp_path / "in.mt", L)
    orc.write_pa(tmp_path / "p.pa", rr, bb)
    orc.run_ref("makermt", ["-s", 66, "-p", "p", "in", "x"], tmp_path)
    r, o = both(progs, orc, tmp_path, [], "c")
    assert r.returncode == 1 and "#! Terminated: a too large, ITMAX too small in gcf." in r.stdout, "the recipe no longer fails in the reference"
    assert o.returncode == 1, (o.returncode, o.stdout[-300:], o.stderr[-300:])
    assert "#! Terminated: a too large, ITMAX too small in gcf." in o.stdout


def test_small_variance_warnings_are_printed_where_the_reference_prints_them(progs, orc, tmp_path):
    """do_mctest warns about every tree pair whose centred replicate rows nearly coincide (src/consel.c:1403-1415),
    between the progress dots of the rows; the sums live on the device here, consel.cpp replays them on the host
    for blocks this small"""
    make_rmt(orc, tmp_path, 21, 6, 40, [0.7, 1.0, 1.3], [40, 50, 45], 0.02)
    r, o = both(progs, orc, tmp_path, [], "w")
    assert r.returncode == o.returncode == 0

    def section(text):
        a = text.index("# calculating the variances")
        return text[a:text.index("# calculating weighted kh-pvalue")]
    assert "small variance" in section(r.stdout), "the recipe no longer triggers the warning in the reference"
    assert section(o.stdout) == section(r.stdout)


def test_rep_mode_runs_the_au_test_on_a_single_scale_like_the_reference(progs, orc, tmp_path):
    """do_repmode has no `kk<2` guard (src/consel.c:1231-1240): `consel -R` on a one-scale .rep runs the degenerate
    fit and writes .pv / .ci / .cnt"""
    import shutil
    make_rmt(orc, tmp_path, 31, 5, 80, [1.0], [70], 0.106)
    orc.run_ref("consel", ["-r", "x", "s"], tmp_path)
    shutil.copy(tmp_path / "s.rep", tmp_path / "x.rep")
    r, o = both(progs, orc, tmp_path, ["-R", "-c"], "q")
    assert r.returncode == o.returncode == 0, o.stdout[-400:]
    assert (tmp_path / "rq.ci").exists() and (tmp_path / "rq.cnt").exists()
    same_files(orc, tmp_path, "q", binary=(), text=("pv", "ci", "cnt"))


def test_scale_without_draws_and_one_termination_line(progs, orc, tmp_path):
    """n' = (int)(r*n) = 0 (three sites under r = 0.31): the reference stores 0.0/0.0 for every replicate of that
    scale and goes on (src/makermt.c:52-70) -- the .rmt is byte-equal; consel on it ends in `luinverse: singular
    matrix` in several fits at once: one `#! Terminated` line, exit status 1"""
    rng = np.random.default_rng(4)
    L = (-2.3 - 4.07 * rng.exponential(size=3))[None, :] + 0.2 * rng.standard_normal((5, 3))
    orc.write_mt(tmp_path / "in.mt", L)
    orc.write_pa(tmp_path / "p.pa", [0.31, 1.0, 1.58], [9, 8, 6])
    orc.run_ref("makermt", ["-s", 3, "-p", "p", "in", "r"], tmp_path)
    o = subprocess.run([str(progs / "makermt"), "-s", "3", "-p", "p", "in", "o"], cwd=tmp_path, capture_output=True, text=True)
    assert o.returncode == 0, o.stdout[-300:]
    for ext in ("rmt", "vt"):
        assert (tmp_path / f"o.{ext}").read_bytes() == (tmp_path / f"r.{ext}").read_bytes(), ext
    (tmp_path / "x.rmt").write_bytes((tmp_path / "r.rmt").read_bytes())
    r, o = both(progs, orc, tmp_path, [], "z")
    assert r.returncode == 1 and "#! Terminated: luinverse: singular matrix." in r.stdout, "the recipe no longer fails in the reference"
    assert o.returncode == 1 and o.stdout.count("#! Terminated") == 1 and "#! Terminated: luinverse: singular matrix." in o.stdout


def test_makerep_scale_without_draws(progs, orc, tmp_path):
    """makerep on the same toy input: the statistics of a scale that draws nothing are the 1e30 calcassmins starts
    from (every comparison with 0.0/0.0 is false); .rep byte-equal to the reference's"""
    if orc.ref_bin("makerep") is None:
        pytest.skip("oracle/_ref/makerep not built")
    rng = np.random.default_rng(4)
    L = (-2.3 - 4.07 * rng.exponential(size=3))[None, :] + 0.2 * rng.standard_normal((5, 3))
    orc.write_mt(tmp_path / "in.mt", L)
    orc.write_pa(tmp_path / "p.pa", [0.31, 1.0, 1.58], [9, 8, 6])
    orc.run_ref("makerep", ["-s", 3, "-p", "p", "in", "r"], tmp_path)
    o = subprocess.run([str(progs / "makerep"), "-s", "3", "-p", "p", "in", "o"], cwd=tmp_path, capture_output=True, text=True)
    assert o.returncode == 0, o.stdout[-300:]
    assert (tmp_path / "o.rep").read_bytes() == (tmp_path / "r.rep").read_bytes()
