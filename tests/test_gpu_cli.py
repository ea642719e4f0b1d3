"""Drop-in command lines on the GPU: our makermt -> consel -> catpv against the files the
unmodified reference programs wrote for the same seeds (tests/golden)."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "consel_b200" / "bin"


def run(prog, args, cwd):
    r = subprocess.run([str(BIN / prog)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def same_numbers(a: Path, b: Path, orc):
    x, y = orc.read_text_blocks(a), orc.read_text_blocks(b)
    assert len(x) == len(y)
    assert np.allclose(x, y, rtol=2e-9, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("case", ["g100", "tiny"])
def test_makermt_consel_catpv_pipeline(case, gold, orc, tmp_path):
    if case == "g100":
        shutil.copy(gold / "mam15.lnf", tmp_path / "mam15.lnf")
        out = run("makermt", ["-s", "1234", "-b", "0.01", "--paml", "mam15", "g100"], tmp_path)
    else:
        shutil.copy(gold / "tiny.mt", tmp_path / "tiny.mt")
        shutil.copy(gold / "tiny.pa", tmp_path / "tiny.pa")
        out = run("makermt", ["-s", "7", "-p", "tiny", "tiny"], tmp_path)
    assert "# exit normally" in out and "# start generating total" in out
    assert (tmp_path / f"{case}.rmt").read_bytes() == (gold / f"{case}.rmt").read_bytes()
    assert (tmp_path / f"{case}.vt").read_bytes() == (gold / f"{case}.vt").read_bytes()
    out = run("consel", ["-r", "-c", case], tmp_path)
    assert "# exit normally" in out
    assert (tmp_path / f"{case}.rep").read_bytes() == (gold / f"{case}.rep").read_bytes()
    assert (tmp_path / f"{case}.cnt").read_bytes() == (gold / f"{case}.cnt").read_bytes()
    for ext in ("pv", "ci"):
        same_numbers(tmp_path / f"{case}.{ext}", gold / f"{case}.{ext}", orc)
    txt = run("catpv", [case], tmp_path)
    assert txt == (gold / f"{case}.txt").read_text()


def test_consel_with_edge_associations(gold, orc, tmp_path):
    shutil.copy(gold / "g100.rmt", tmp_path / "g100.rmt")
    shutil.copy(gold / "mam15.ass", tmp_path / "mam15.ass")
    run("consel", ["-a", "mam15", "-r", "-c", "g100", "g100e"], tmp_path)
    assert (tmp_path / "g100e.rep").read_bytes() == (gold / "g100e.rep").read_bytes()
    assert (tmp_path / "g100e.cnt").read_bytes() == (gold / "g100e.cnt").read_bytes()
    same_numbers(tmp_path / "g100e.pv", gold / "g100e.pv", orc)
    same_numbers(tmp_path / "g100e.ci", gold / "g100e.ci", orc)
    assert run("catpv", ["g100e"], tmp_path) == (gold / "g100e.txt").read_text()


@pytest.mark.parametrize("case", ["g100", "g100e"])
def test_consel_rep_mode(case, gold, orc, tmp_path):
    """consel -R (do_repmode, src/consel.c:1195-1242): BP and AU from the rows of a .rep file,
    against what the reference wrote for the same file."""
    shutil.copy(gold / f"{case}.rep", tmp_path / f"{case}.rep")
    out = run("consel", ["-R", "-c", case, case + "R"], tmp_path)
    assert "# exit normally" in out and "MC-TEST" not in out
    assert (tmp_path / f"{case}R.cnt").read_bytes() == (gold / f"{case}R.cnt").read_bytes()
    same_numbers(tmp_path / f"{case}R.pv", gold / f"{case}R.pv", orc)
    same_numbers(tmp_path / f"{case}R.ci", gold / f"{case}R.ci", orc)
    # -r in rep mode rewrites the INPUT file name (fname_rep), rows sorted: a .rep written by consel is a fixed point
    before = (tmp_path / f"{case}.rep").read_bytes()
    run("consel", ["-R", "-r", case, case + "R2"], tmp_path)
    assert (tmp_path / f"{case}.rep").read_bytes() == before
    assert not (tmp_path / f"{case}R2.rep").exists()
    # without the AU test nothing but the .pv is written (:1236-1240)
    run("consel", ["-R", "-c", "--no_au", case, case + "R3"], tmp_path)
    assert (tmp_path / f"{case}R3.pv").exists() and not (tmp_path / f"{case}R3.cnt").exists()


@pytest.mark.parametrize("case", ["g100f", "g100fe"])
def test_consel_rescaling_approximation(case, gold, orc, tmp_path):
    """consel -f [-p scales] (genrep's rescaling branch, src/consel.c:757-771): all target scales
    from the replicates of the scale nearest r=1, against the reference's own -f run."""
    shutil.copy(gold / "g100.rmt", tmp_path / "g100.rmt")
    args = ["-f", "-r", "-c", "g100", case]
    if case == "g100fe":
        shutil.copy(gold / "mam15.ass", tmp_path / "mam15.ass")
        shutil.copy(gold / "fscale.pa", tmp_path / "fscale.pa")
        args = ["-f", "-p", "fscale", "-a", "mam15", "-r", "-c", "g100", case]
    out = run("consel", args, tmp_path)
    assert "# RESCALING" in out and "# exit normally" in out
    assert (tmp_path / f"{case}.rep").read_bytes() == (gold / f"{case}.rep").read_bytes()
    assert (tmp_path / f"{case}.cnt").read_bytes() == (gold / f"{case}.cnt").read_bytes()
    same_numbers(tmp_path / f"{case}.pv", gold / f"{case}.pv", orc)
    same_numbers(tmp_path / f"{case}.ci", gold / f"{case}.ci", orc)


def test_makerep_matches_reference(gold, orc, tmp_path):
    """makerep (src/makerep.c:304-331): resampling + statistics in one process; the .rep files
    (rows unsorted) byte-equal to the ones the reference's makerep wrote for the same seeds, and
    consel -R on them gives the reference's BP/AU numbers."""
    for f in ("mam15.lnf", "mam15.ass", "mr.pa", "tiny.mt", "tiny.pa"):
        shutil.copy(gold / f, tmp_path / f)
    out = run("makerep", ["-s", "1234", "-p", "mr", "-a", "mam15", "--paml", "mam15", "mr100e"], tmp_path)
    assert "# exit normally" in out and "# CM:10" in out
    assert (tmp_path / "mr100e.rep").read_bytes() == (gold / "mr100e.rep").read_bytes()
    run("makerep", ["-s", "7", "-p", "tiny", "tiny", "tinymr"], tmp_path)
    assert (tmp_path / "tinymr.rep").read_bytes() == (gold / "tinymr.rep").read_bytes()
    # same replicates as makermt | consel -a: rep mode on makerep's file reproduces that run's BP/AU
    run("consel", ["-R", "-c", "mr100e", "mr100eR"], tmp_path)
    assert (tmp_path / "mr100eR.cnt").read_bytes() == (gold / "g100eR.cnt").read_bytes()
    same_numbers(tmp_path / "mr100eR.pv", gold / "g100eR.pv", orc)
    # int8 sums: exact sums correctly rounded, statistics within 1e-9 of the strict ones
    run("makerep", ["-s", "1234", "-p", "mr", "-a", "mam15", "--paml", "mam15", "--sum", "int8", "mr100i"], tmp_path)
    a, b = orc.read_rep(tmp_path / "mr100i.rep"), orc.read_rep(gold / "mr100e.rep")
    for x, y in zip(a["stats"], b["stats"]):
        assert np.allclose(x, y, rtol=0, atol=1e-9)


def test_multi_gene_mode(gold, orc, tmp_path):
    """makermt -g ; consel -g (do_rmtmultimode, src/consel.c:1056-1192): per-gene statistics and counts
    on the device, congcnt + count-matrix fit on the host, against the reference's own run."""
    for f in ("ga.mt", "gb.mt", "gc.mt", "genes.svt", "pairs.ass", "mg.pa", "fscale.pa"):
        shutil.copy(gold / f, tmp_path / f)
    run("makermt", ["-s", "11", "-p", "mg", "-g", "genes"], tmp_path)
    out = run("consel", ["-g", "-c", "genes", "mgout"], tmp_path)
    assert "# MF:3" in out and "# exit normally" in out
    assert (tmp_path / "mgout.cnt").read_bytes() == (gold / "mgout.cnt").read_bytes()
    same_numbers(tmp_path / "mgout.pv", gold / "mgout.pv", orc)
    out = run("consel", ["-g", "-f", "-p", "fscale", "-a", "pairs", "--smooth_cnt", "-c", "genes", "mgoutf"], tmp_path)
    assert "# RESCALING" in out
    same_numbers(tmp_path / "mgoutf.cnt", gold / "mgoutf.cnt", orc)
    same_numbers(tmp_path / "mgoutf.pv", gold / "mgoutf.pv", orc)


def test_quick_mctest(gold, orc, tmp_path):
    """--quick_mctest (sw_pvaldist=0, src/consel.c:1875-1884): KH/SH/wKH/wSH p-values from plain counts."""
    shutil.copy(gold / "g100.rmt", tmp_path / "g100.rmt")
    shutil.copy(gold / "mam15.ass", tmp_path / "mam15.ass")
    run("consel", ["--quick_mctest", "--no_au", "-a", "mam15", "g100", "g100q"], tmp_path)
    same_numbers(tmp_path / "g100q.pv", gold / "g100q.pv", orc)
    a, b = orc.read_text_blocks(gold / "g100q.pv"), orc.read_text_blocks(gold / "g100e.pv")
    assert len(a) != len(b) or not np.allclose(a, b)  # the switch does change the numbers


def test_pdm_aic_correction(gold, orc, tmp_path):
    """--pdm / --aictype / --ppn (read_pdm, do_pdm, BIC branch of do_batest: src/consel.c:675-744, 1325-1332)."""
    shutil.copy(gold / "g100.rmt", tmp_path / "g100.rmt")
    shutil.copy(gold / "pdm.vt", tmp_path / "pdm.vt")
    out = run("consel", ["--pdm", "pdm", "g100", "g100p"], tmp_path)
    assert "# aictype:2" in out and "# ppn:1" in out
    same_numbers(tmp_path / "g100p.pv", gold / "g100p.pv", orc)
    same_numbers(tmp_path / "g100p.ci", gold / "g100p.ci", orc)
    run("consel", ["--pdm", "pdm", "--aictype", "3", "--ppn", "3414", "-r", "-c", "g100", "g100p3"], tmp_path)
    assert (tmp_path / "g100p3.rep").read_bytes() == (gold / "g100p3.rep").read_bytes()
    assert (tmp_path / "g100p3.cnt").read_bytes() == (gold / "g100p3.cnt").read_bytes()
    same_numbers(tmp_path / "g100p3.pv", gold / "g100p3.pv", orc)


def test_cli_error_protocol(tmp_path):
    r = subprocess.run([str(BIN / "makermt"), "--bogus"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "#! Terminated: error in command line." in r.stdout
    r = subprocess.run([str(BIN / "consel"), "nofile"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "#! Terminated: cant open nofile.rmt." in r.stdout


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "makermt").exists(), reason="oracle/_ref not built")
def test_full_pipeline_against_fresh_reference_run(gold, orc, tmp_path):
    """k10s-sized run (10 x 1000 replicates): reference and ours side by side, same seed."""
    shutil.copy(gold / "mam15.lnf", tmp_path / "mam15.lnf")
    orc.write_pa(tmp_path / "k.pa", [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4], [1000] * 10)
    orc.run_ref("makermt", ["-s", "99", "-p", "k", "--paml", "mam15", "ref"], tmp_path)
    orc.run_ref("consel", ["-r", "-c", "ref"], tmp_path)
    run("makermt", ["-s", "99", "-p", "k", "--paml", "mam15", "our"], tmp_path)
    run("consel", ["-r", "-c", "our"], tmp_path)
    for ext in ("rmt", "rep", "cnt"):
        assert (tmp_path / f"our.{ext}").read_bytes() == (tmp_path / f"ref.{ext}").read_bytes(), ext
    same_numbers(tmp_path / "our.pv", tmp_path / "ref.pv", orc)
    same_numbers(tmp_path / "our.ci", tmp_path / "ref.ci", orc)
    assert run("catpv", ["our"], tmp_path).replace("our.pv", "X") == orc.run_ref("catpv", ["ref"], tmp_path).replace("ref.pv", "X")


def test_in_process_pipeline_matches_reference_pv(gold, orc):
    """consel_b200.pipeline.run: makermt+consel without files == the reference's .pv/.ci"""
    import consel_b200 as cg
    from consel_b200 import pipeline
    cg.init()
    cg.set_shard(0, 1)
    for name, seed in (("g100", 1234), ("tiny", 7)):
        ref_pv = orc.read_pv(gold / f"{name}.pv")
        rmt = orc.read_rmt(gold / f"{name}.rmt")
        if name == "g100":
            L, rr = orc.read_lnf(gold / "mam15.lnf"), [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
        else:
            L = orc.read_mt(gold / "tiny.mt")
            rr, _ = orc.read_pa(gold / "tiny.pa")
        res = pipeline.run(L, rr, rmt["bb"], seed)
        assert np.array_equal(res["item"], ref_pv["item"])
        assert np.allclose(res["stat"], ref_pv["stat"], rtol=6e-9)
        assert np.allclose(res["pv"], ref_pv["pv"], rtol=6e-9, atol=1e-12), (name, res["pv"], ref_pv["pv"])
        assert np.allclose(res["se"], ref_pv["se"], rtol=6e-9, atol=1e-12)
        assert np.allclose(res["ax"], ref_pv["ax"], rtol=6e-9, atol=1e-12)
        assert np.array_equal(res["cnt"], orc.read_cnt(gold / f"{name}.cnt")["cnt"])
    # the tensor-core sums give the same table to printed precision (catpv prints 3 decimals)
    res8 = pipeline.run(L, rr, rmt["bb"], 7, sum_mode=cg.SUM_INT8)
    assert np.allclose(res8["pv"], ref_pv["pv"], rtol=1e-6, atol=1e-9)


def test_makermt_extension_flags(gold, orc, tmp_path):
    """--sum int8 (tensor-core fixed point): every replicate sum within 1e-9 of the reference's
    .rmt (north star tolerance; the value is the correctly rounded exact sum), same header;
    --rng philox: a valid .rmt of the same shape that consel accepts."""
    shutil.copy(gold / "mam15.lnf", tmp_path / "mam15.lnf")
    run("makermt", ["-s", "1234", "-b", "0.01", "--paml", "--sum", "int8", "mam15", "i8"], tmp_path)
    ref, got = orc.read_rmt(gold / "g100.rmt"), orc.read_rmt(tmp_path / "i8.rmt")
    assert np.array_equal(ref["datvec"], got["datvec"]) and np.array_equal(ref["rr"], got["rr"]) and np.array_equal(ref["bb"], got["bb"])
    for a, b in zip(ref["reps"], got["reps"]):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * np.abs(a).max()
    run("makermt", ["-s", "1234", "-b", "0.01", "--paml", "--rng", "philox", "--sum", "int8", "mam15", "ph"], tmp_path)
    ph = orc.read_rmt(tmp_path / "ph.rmt")
    assert [r.shape for r in ph["reps"]] == [r.shape for r in ref["reps"]]
    assert all(np.isfinite(r).all() for r in ph["reps"])
    out = run("consel", ["ph"], tmp_path)
    assert "# exit normally" in out and (tmp_path / "ph.pv").exists()


def test_full_size_mam15_pipeline_and_philox_within_monte_carlo_error(gold, orc):
    """C1 at full size (10 x 10^4 replicates).  (1) Driven by the reference's seeded stream the in-process
    pipeline reproduces the reference's own full-size .pv (tests/golden/full.pv, md5 pinned in
    MANIFEST.json) to printed precision.  (2) The north star's third correctness clause: under an
    INDEPENDENT generator (Philox, two different keys) every BP/KH/SH/wKH/wSH/AU/NP p-value agrees
    with the reference's within Monte-Carlo error: 5 standard errors of the difference, the SEs being
    the .pv SE columns (getseval, src/consel.c:2151-2158, and the AU fit's own SE)."""
    import consel_b200 as cg
    from consel_b200 import pipeline
    cg.init()
    cg.set_shard(0, 1)
    L = orc.read_lnf(gold / "mam15.lnf")
    rr, bb = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4], [10000] * 10
    ref = orc.read_pv(gold / "full.pv")
    same = pipeline.run(L, rr, bb, 1234)
    assert np.array_equal(same["item"], ref["item"])
    assert np.allclose(same["pv"], ref["pv"], rtol=6e-9, atol=1e-12)
    assert np.allclose(same["se"], ref["se"], rtol=6e-9, atol=1e-12)
    runs = [pipeline.run(L, rr, bb, seed, rng_mode=cg.RNG_PHILOX, sum_mode=cg.SUM_INT8) for seed in (1234, 99)]
    cols = [0, 2, 3, 4, 5, 6, 7]  # bp kh sh wkh wsh au np (column 1, pp, does not depend on the replicates)
    pairs = [(runs[0], ref), (runs[1], ref), (runs[0], runs[1])]
    worst = 0.0
    for a, b in pairs:
        assert np.array_equal(a["item"], b["item"])
        assert np.array_equal(a["pv"][:, 1], b["pv"][:, 1]) or np.allclose(a["pv"][:, 1], b["pv"][:, 1], rtol=1e-12)
        for c in cols:
            tol = 5.0 * np.sqrt(a["se"][:, c] ** 2 + b["se"][:, c] ** 2) + 2.0e-4  # + two counts of 10^4
            diff = np.abs(a["pv"][:, c] - b["pv"][:, c])
            worst = max(worst, float(np.max(diff / tol)))
            assert np.all(diff <= tol), (c, a["pv"][:, c], b["pv"][:, c], tol)
    # and the generator is not accidentally the reference stream
    assert not np.array_equal(runs[0]["cnt"], same["cnt"])
