#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference binaries in oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference
exists and `make -C oracle ref` has compiled it).  The fixtures are small on
purpose; full-size runs are pinned by md5 in MANIFEST.json.

Fixtures
  mam15.lnf                 reference example input (example/mam15.lnf), config C1
  g100.{rmt,vt,pv,ci,cnt,rep,txt}   makermt -s 1234 -b 0.01 --paml mam15 ; consel -r -c ; catpv
  g100e.{pv,ci,cnt,rep,txt} + mam15.ass   treeass mam15 ; consel -a (edge associations)
  g100R / g100eR .{pv,ci,cnt}             consel -R on the two .rep files above (rep mode)
  g100f / g100fe .{pv,ci,cnt,rep} + fscale.pa   consel -f [-p fscale -a mam15] (rescaling approximation)
  tiny.{mt,pa,rmt,vt,...}   synthetic K=6 n=40, ragged scales/replicate counts
  mr100e.rep, tinymr.rep + mr.pa   makerep -s 1234 -p mr -a mam15 --paml mam15 ; makerep -s 7 -p tiny tiny
  {g100,g100e,tiny,g100fe}C.pv   consel -C on the .cnt files above (cnt mode)
  ga/gb/gc.mt genes.svt pairs.ass mg.pa mgout* mgoutf*   makermt -s 11 -p mg -g genes ; consel -g (multi-gene mode)
  MANIFEST.json             md5 of every fixture + md5s of the full-size seeded runs
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

GOLD = ROOT / "tests" / "golden"
REF_EXAMPLE = Path("/root/reference/example")


def md5(p: Path) -> str:
    return hashlib.md5(p.read_bytes()).hexdigest()


def main() -> None:
    orc.build(ref=True)
    GOLD.mkdir(parents=True, exist_ok=True)
    work = Path(tempfile.mkdtemp(prefix="consel_golden_"))
    manifest = {"generated_by": "oracle/gen_golden.py", "reference_flags": "-O2 -ffp-contract=off", "files": {},
                "full_size": {}}

    shutil.copy(REF_EXAMPLE / "mam15.lnf", work / "mam15.lnf")
    shutil.copy(REF_EXAMPLE / "mam15.tpl", work / "mam15.tpl")
    keep = ["mam15.lnf"]

    # --- g100: mam15, 10 scales x 100 replicates, seed 1234
    orc.run_ref("makermt", ["-s", "1234", "-b", "0.01", "--paml", "mam15", "g100"], work)
    orc.run_ref("consel", ["-r", "-c", "g100"], work)
    (work / "g100.txt").write_text(orc.run_ref("catpv", ["g100"], work))
    keep += [f"g100.{e}" for e in ("rmt", "vt", "pv", "ci", "cnt", "rep", "txt")]

    # --- g100e: edge associations on the same replicates
    (work / "mam15ass.log").write_text(orc.run_ref("treeass", ["mam15"], work))
    orc.run_ref("consel", ["-a", "mam15", "-r", "-c", "g100", "g100e"], work)
    (work / "g100e.txt").write_text(orc.run_ref("catpv", ["g100e"], work))
    keep += ["mam15.ass"] + [f"g100e.{e}" for e in ("pv", "ci", "cnt", "rep", "txt")]

    # --- rep mode (consel -R): BP/AU from the statistic rows of a .rep file
    for case in ("g100", "g100e"):
        orc.run_ref("consel", ["-R", "-c", case, case + "R"], work)
        keep += [f"{case}R.{e}" for e in ("pv", "ci", "cnt")]

    # --- rescaling approximation (consel -f): default target scales, and scales from a .pa file
    orc.run_ref("consel", ["-f", "-r", "-c", "g100", "g100f"], work)
    orc.write_pa(work / "fscale.pa", [0.4, 0.75, 1.0, 1.6], [1, 1, 1, 1])
    orc.run_ref("consel", ["-f", "-p", "fscale", "-a", "mam15", "-r", "-c", "g100", "g100fe"], work)
    keep += ["fscale.pa"] + [f"{c}.{e}" for c in ("g100f", "g100fe") for e in ("pv", "ci", "cnt", "rep")]

    # --- tiny: synthetic ragged case (different r and B per scale, bn < n and > n)
    rng = np.random.Generator(np.random.PCG64(20261019))
    K, n = 6, 40
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
    orc.write_mt(work / "tiny.mt", L)
    orc.write_pa(work / "tiny.pa", [0.5, 0.77, 1.0, 1.37, 2.05], [37, 64, 50, 33, 41])
    orc.run_ref("makermt", ["-s", "7", "-p", "tiny", "tiny"], work)
    orc.run_ref("consel", ["-r", "-c", "tiny"], work)
    (work / "tiny.txt").write_text(orc.run_ref("catpv", ["tiny"], work))
    keep += [f"tiny.{e}" for e in ("mt", "pa", "rmt", "vt", "pv", "ci", "cnt", "rep", "txt")]

    # --- makerep: resampling + statistics in one process (rows unsorted), same seeds as above
    orc.write_pa(work / "mr.pa", [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4], [100] * 10)
    orc.run_ref("makerep", ["-s", "1234", "-p", "mr", "-a", "mam15", "--paml", "mam15", "mr100e"], work)
    orc.run_ref("makerep", ["-s", "7", "-p", "tiny", "tiny", "tinymr"], work)
    keep += ["mr.pa", "mr100e.rep", "tinymr.rep"]

    # --- multi-gene mode (makermt -g ; consel -g): three synthetic genes of different lengths
    rng = np.random.Generator(np.random.PCG64(77))
    for gname, gn in (("ga", 60), ("gb", 85), ("gc", 47)):
        gbase = -2.3 - 4.07 * rng.exponential(size=gn)
        gL = gbase[None, :] + 0.25 * rng.standard_normal((6, gn)) - 0.9 * np.arange(6)[:, None] / gn
        orc.write_mt(work / f"{gname}.mt", gL)
    (work / "genes.svt").write_text("3\nga gb gc\n")
    (work / "pairs.ass").write_text("3\n2 0 1\n2 2 3\n2 4 5\n")
    orc.write_pa(work / "mg.pa", [0.5, 0.7, 0.85, 1.0, 1.2, 1.4], [300] * 6)
    orc.run_ref("makermt", ["-s", "11", "-p", "mg", "-g", "genes"], work)
    orc.run_ref("consel", ["-g", "-c", "genes", "mgout"], work)
    orc.run_ref("consel", ["-g", "-f", "-p", "fscale", "-a", "pairs", "--smooth_cnt", "-c", "genes", "mgoutf"], work)
    keep += ["ga.mt", "gb.mt", "gc.mt", "genes.svt", "pairs.ass", "mg.pa", "mgout.pv", "mgout.cnt", "mgoutf.pv", "mgoutf.cnt"]

    # --- --quick_mctest: KH/SH p-values from plain counts
    orc.run_ref("consel", ["--quick_mctest", "--no_au", "-a", "mam15", "g100", "g100q"], work)
    keep.append("g100q.pv")

    # --- AIC correction (--pdm), default aictype 2 and aictype 3 with --ppn; --pmin in cnt mode
    (work / "pdm.vt").write_text("15\n" + " ".join(str(v) for v in [3, 5, 2, 8, 4, 6, 1, 7, 3, 2, 9, 4, 5, 6, 2]) + "\n")
    orc.run_ref("consel", ["--pdm", "pdm", "g100", "g100p"], work)
    orc.run_ref("consel", ["--pdm", "pdm", "--aictype", "3", "--ppn", "3414", "-r", "-c", "g100", "g100p3"], work)
    orc.run_ref("consel", ["-C", "--pmin", "0.5", "g100", "g100Cp"], work)
    keep += ["pdm.vt", "g100p.pv", "g100p.ci", "g100p3.pv", "g100p3.rep", "g100p3.cnt", "g100Cp.pv"]

    # --- cnt mode (consel -C): the AU fit from the counts alone
    for case in ("g100", "g100e", "tiny", "g100fe"):
        orc.run_ref("consel", ["-C", case, case + "C"], work)
        keep.append(case + "C.pv")

    for name in keep:
        shutil.copy(work / name, GOLD / name)
        manifest["files"][name] = md5(GOLD / name)

    # --- full-size pins (md5 only): SURVEY Appendix C
    orc.run_ref("makermt", ["-s", "1234", "--paml", "mam15", "full"], work)
    orc.run_ref("consel", ["-r", "-c", "full"], work)
    for e in ("rmt", "vt", "pv", "ci", "cnt", "rep"):
        manifest["full_size"][f"mam15_s1234.{e}"] = md5(work / f"full.{e}")
    # the full-size p-value table itself is small: kept as a fixture (C1's reference answer; the
    # Monte-Carlo-error test of the counter-based resampler compares against it)
    shutil.copy(work / "full.pv", GOLD / "full.pv")
    manifest["files"]["full.pv"] = md5(GOLD / "full.pv")
    orc.run_ref("makermt", ["-s", "1234", "-p", str(REF_EXAMPLE / "k10s"), "--paml", "mam15", "g10s"], work)
    orc.run_ref("consel", ["-r", "-c", "g10s"], work)
    for e in ("rmt", "pv", "ci", "cnt", "rep"):
        manifest["full_size"][f"mam15_s1234_k10s.{e}"] = md5(work / f"g10s.{e}")
    # a few spot values of the full run
    full = orc.read_rmt(work / "full.rmt")
    manifest["full_size"]["spot"] = {
        "datvec0_2": [float(x) for x in full["datvec"][:3]],
        "rr1": [float(x) for x in full["rr"]],
        "rep_s0_k0_b0_3": [float(x) for x in full["reps"][0][0, :4]],
        "cnt_row0": [float(x) for x in orc.read_cnt(work / "full.cnt")["cnt"][0]],
    }
    (GOLD / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")
    shutil.rmtree(work)
    print("wrote", len(keep) + 1, "files to", GOLD)


if __name__ == "__main__":
    main()
