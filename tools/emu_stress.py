#!/usr/bin/env python
"""Randomised differential test on the CPU: the product's makermt / consel programs (their own sources linked
against the EMULATED device library, tests/emu/) against the UNMODIFIED reference binaries (oracle/_ref) on random
small inputs -- random tree / site counts, scale sets, replicate counts, tied rows and columns, random edge
associations, the -f / -R / --quick_mctest variants.  `.rmt` `.vt` `.rep` `.cnt` must be byte-equal, `.pv` `.ci`
equal to the printed digits.

    python tools/emu_stress.py [SEED=1] [CASES=50] [workdir=/tmp/emu_stress]      (EMU_STRESS_KEEP=1 keeps every case directory;
                                                                                   EMU_STRESS_SIZE=medium: tens of trees, ten scales x thousands of replicates)
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "emu"))
from oracle import oracle as orc  # noqa: E402


def build_programs(work: Path) -> Path:
    import build_emu_lib
    lib = build_emu_lib.build(work / "lib")
    host = ROOT / "consel_b200" / "host"
    out = work / "bin"
    out.mkdir(exist_ok=True)
    programs = {"consel": ["consel.cpp", "pipeline.cpp", "audriver.cpp", "aufit.cpp", "specfun.cpp", "refio.cpp"],
                "makermt": ["makermt.cpp", "refio.cpp"], "makerep": ["makerep.cpp", "refio.cpp"]}
    for prog, srcs in programs.items():
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-w", "-o", str(out / prog),
                        *[str(host / f) for f in srcs], f"-L{lib.parent}", "-lconselgpu_emu", f"-Wl,-rpath,{lib.parent}", "-lm"],
                       check=True)
    return out


def write_ass(path: Path, sets):
    with open(path, "w") as f:
        f.write(f"\n# EDGE->TREE\n{len(sets)}\n")
        for i, s in enumerate(sets):
            f.write(f"\n# item: {i}\n#!VEC:\n{len(s)}\n" + " ".join(f"{int(x):10d}" for x in s) + " \n")


def protocol(text: str, stem: str) -> list:
    """the stdout protocol without what differs by design: the program id line, wall times, the wording of the
    sort / count step, and the output file stem"""
    import re
    out = []
    for ln in text.replace(stem, "OUT").splitlines():
        if ln.startswith("# $Id") or ln.startswith("# consel_b200") or "time elapsed" in ln:
            continue
        ln = re.sub(r"# (sorting the replicates|counting the replicates \(sort-free row summaries on the device\))", "# SORT", ln)
        out.append(ln.rstrip())
    return out


def numbers_equal(a: Path, b: Path) -> bool:
    x, y = orc.read_text_blocks(a), orc.read_text_blocks(b)
    return len(x) == len(y) and bool(np.allclose(x, y, rtol=2e-9, atol=1e-12, equal_nan=True))


STATS = {"variants": 0, "both_ok": 0, "both_refused": 0}
RR10 = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]


def one_case(rng, bindir: Path, d: Path) -> list:
    bad = []
    K = int(rng.integers(2, 15))
    n = int(rng.choice([int(rng.integers(3, 40)), int(rng.integers(40, 700))]))
    if os.environ.get("EMU_STRESS_SIZE") == "toy":   # a handful of sites: scales that draw nothing at all (n' = 0), NaN blocks
        n = int(rng.integers(2, 7))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = base[None, :] + rng.choice([0.01, 0.106, 1.0]) * rng.standard_normal((K, n))
    if rng.random() < 0.3 and K > 2:
        L[int(rng.integers(1, K))] = L[0]                       # tied trees
    if rng.random() < 0.3 and n > 4:
        L[:, int(rng.integers(1, n))] = L[:, 0]                 # repeated site pattern
    if rng.random() < 0.2:
        L = np.round(L, 3)                                      # few significant digits
    kk = int(rng.integers(1, 12))
    rr = np.sort(np.round(rng.uniform(0.3, 2.0, size=kk), 2))
    if rng.random() < 0.7:
        rr[int(rng.integers(0, kk))] = 1.0
        rr = np.sort(rr)
    bb = rng.integers(2, 250, size=kk)
    if os.environ.get("EMU_STRESS_SIZE") == "medium":  # the regime of real use: tens of trees, thousands of replicates, well-behaved fits
        K = int(rng.integers(15, 70))
        n = int(rng.integers(200, 1500))
        base = -2.3 - 4.07 * rng.exponential(size=n)
        L = base[None, :] + rng.choice([0.106, 0.3]) * rng.standard_normal((K, n)) - rng.uniform(0, 3) * np.arange(K)[:, None] / n
        kk = 10
        rr = np.array(RR10)
        bb = np.full(kk, int(rng.integers(500, 3000)))
    seed = int(rng.integers(1, 10**6))
    orc.write_mt(d / "in.mt", L)
    orc.write_pa(d / "p.pa", rr, bb)
    desc = f"K={K} n={n} kk={kk} rr={rr.tolist()} bb={bb.tolist()} seed={seed}"

    def launch(exe, args):  # a program that does not return within two minutes is reported as status 124, not waited for
        try:
            r = subprocess.run([str(exe)] + [str(a) for a in args], cwd=d, capture_output=True, text=True, timeout=120)
            return r.returncode, r.stdout
        except subprocess.TimeoutExpired:
            print(f"  TIMEOUT {exe} {args} in {d}", flush=True)
            return 124, ""

    def ours(prog, args):
        return launch(bindir / prog, args)

    def ref(prog, args):
        return launch(orc.ref_bin(prog), args)

    ra, rout = ref("makermt", ["-s", seed, "-p", "p", "in", "ref"])
    oa, oout = ours("makermt", ["-s", seed, "-p", "p", "in", "our"])
    if ra != oa:
        return [f"makermt exit status {oa} vs reference {ra}: {desc}\n{oout[-400:]}"]
    if protocol(rout, "ref") != protocol(oout, "our"):
        bad.append(f"makermt stdout differs: {desc}")
    if ra != 0:
        return []
    for ext in ("rmt", "vt"):
        if (d / f"our.{ext}").read_bytes() != (d / f"ref.{ext}").read_bytes():
            bad.append(f".{ext} differs: {desc}")
    # makermt variants: -b (replicate factor, truncated to int), -f (the rescaling defaults: 1 scale), the built-in ten
    # scales, -g (multi-gene: file names from a .svt, the MT19937 stream continues from gene to gene), and makerep
    mk = []
    if rng.random() < 0.5:
        mk.append((["-s", seed, "-b", f"{rng.uniform(0.2, 2.5):.3f}", "-p", "p", "in"], "mb"))
    if rng.random() < 0.3:
        mk.append((["-s", seed, "-b", "0.004", "in"], "md"))        # built-in scales, 40 replicates each
    if rng.random() < 0.3:
        mk.append((["-s", seed, "-f", "-b", "0.01", "in"], "mf"))
    for args, stem in mk:
        r1, t1 = ref("makermt", args + [f"r{stem}"])
        o1, t2 = ours("makermt", args + [f"o{stem}"])
        if r1 < 0:  # the reference itself crashed (e.g. SIGSEGV on a scale whose -b factor leaves 0 replicates): nothing to match
            STATS["ref_crashed"] = STATS.get("ref_crashed", 0) + 1
        elif r1 != o1:
            bad.append(f"makermt {args} exit status {o1} vs reference {r1}: {desc}")
        elif r1 == 0:
            for ext in ("rmt", "vt"):
                if (d / f"o{stem}.{ext}").read_bytes() != (d / f"r{stem}.{ext}").read_bytes():
                    bad.append(f"makermt {args} .{ext} differs: {desc}")
            if protocol(t1, f"r{stem}") != protocol(t2, f"o{stem}"):
                bad.append(f"makermt {args} stdout differs: {desc}")
    if rng.random() < 0.35:
        n2 = int(max(2, n // 2))
        L2 = (-2.3 - 4.07 * rng.exponential(size=n2))[None, :] + 0.2 * rng.standard_normal((K, n2))
        orc.write_mt(d / "g2.mt", L2)
        shutil.copy(d / "in.mt", d / "g1.mt")
        for who in ("r", "o"):
            (d / f"{who}g.svt").write_text("2\ng1 g2\n")
        # input names and output names come from two .svt files (src/makermt.c:303-312)
        (d / "gin.svt").write_text("2\ng1 g2\n")
        (d / "rgo.svt").write_text("2\nrg1 rg2\n")
        (d / "ogo.svt").write_text("2\nog1 og2\n")
        r1, t1 = ref("makermt", ["-s", seed, "-g", "-p", "p", "gin", "rgo"])
        o1, t2 = ours("makermt", ["-s", seed, "-g", "-p", "p", "gin", "ogo"])
        if r1 != o1:
            bad.append(f"makermt -g exit status {o1} vs reference {r1}: {desc}\n{t2[-300:]}")
        elif r1 == 0:
            for g in ("g1", "g2"):
                for ext in ("rmt", "vt"):
                    pa, pb = d / f"o{g}.{ext}", d / f"r{g}.{ext}"
                    if pa.exists() != pb.exists() or (pb.exists() and pa.read_bytes() != pb.read_bytes()):
                        bad.append(f"makermt -g {g}.{ext} differs: {desc}")
            # consel -g on the two genes the REFERENCE wrote (do_rmtmultimode, src/consel.c:1056-1192)
            (d / "gg.svt").write_text("2\nrg1 rg2\n")
            gx = ["-g"] + (["--smooth_cnt"] if rng.random() < 0.4 else []) + (["-f"] if rng.random() < 0.3 else []) + ["-c"]
            r2, t3 = ref("consel", gx + ["gg", "rgg"])
            o2, t4 = ours("consel", gx + ["gg", "ogg"])
            STATS["variants"] += 1
            if r2 != o2:
                bad.append(f"consel {gx} exit status {o2} vs reference {r2}: {desc}\n{t4[-300:]}")
            elif r2 == 0:
                STATS["both_ok"] += 1
                for ext in ("cnt", "pv", "ci"):
                    pa, pb = d / f"ogg.{ext}", d / f"rgg.{ext}"
                    if pa.exists() != pb.exists() or (pb.exists() and not numbers_equal(pa, pb)):
                        bad.append(f"consel {gx} .{ext} differs: {desc}")
            else:
                STATS["both_refused"] += 1
    if rng.random() < 0.35:   # the PAML .lnf reader (src/freadmat.c:110-154): patterns with multiplicities
        npat = int(rng.integers(1, 60))
        mult = rng.integers(1, 6, size=npat)
        with open(d / "pl.lnf", "w") as f:
            f.write(f"  {K}  {int(mult.sum())}  {npat}\n\n")
            for i in range(K):
                f.write(f"\n {i + 1}\n\n")
                v = -rng.exponential(4.0, size=npat) - 1.0
                for j in range(npat):
                    f.write(f"{j + 1:6d}{int(mult[j]):6d}{v[j]:16.10f}{rng.uniform(0, 99):9.2f}  ACGT\n")
        r1, t1 = ref("makermt", ["-s", seed, "-p", "p", "--paml", "pl", "rpl"])
        o1, t2 = ours("makermt", ["-s", seed, "-p", "p", "--paml", "pl", "opl"])
        if r1 != o1:
            bad.append(f"makermt --paml exit status {o1} vs reference {r1}: {desc}\n{t2[-300:]}")
        elif r1 == 0:
            for ext in ("rmt", "vt"):
                if (d / f"opl.{ext}").read_bytes() != (d / f"rpl.{ext}").read_bytes():
                    bad.append(f"makermt --paml .{ext} differs: {desc}")
            # the tensor-core mode on the same file: every sum within 1e-9 of the reference's (contract of --sum int8)
            o3, t4 = ours("makermt", ["-s", seed, "-p", "p", "--paml", "--sum", "int8", "pl", "opl8"])
            if o3 != 0:
                bad.append(f"makermt --sum int8 failed: {desc}\n{t4[-300:]}")
            else:
                a8, b8 = orc.read_rmt(d / "opl8.rmt"), orc.read_rmt(d / "rpl.rmt")
                for x8, y8 in zip(a8["reps"], b8["reps"]):
                    if x8.shape != y8.shape or not np.allclose(x8, y8, rtol=1e-9, atol=0, equal_nan=True):
                        bad.append(f"makermt --sum int8 differs from the reference by more than 1e-9: {desc}")
                        break
    if rng.random() < 0.35 and orc.ref_bin("makerep") is not None:
        r1, t1 = ref("makerep", ["-s", seed, "-p", "p", "in", "rmr"])
        o1, t2 = ours("makerep", ["-s", seed, "-p", "p", "in", "omr"])
        if r1 != o1:
            bad.append(f"makerep exit status {o1} vs reference {r1}: {desc}\n{t2[-300:]}")
        elif r1 == 0:
            for ext in ("rep", "vt"):
                pa, pb = d / f"omr.{ext}", d / f"rmr.{ext}"
                if pa.exists() != pb.exists() or (pb.exists() and pa.read_bytes() != pb.read_bytes()):
                    bad.append(f"makerep .{ext} differs: {desc}")

    # consel variants on the REFERENCE's .rmt (so that a makermt difference does not cascade)
    shutil.copy(d / "ref.rmt", d / "x.rmt")
    variants = [([], "plain")]
    if rng.random() < 0.5:
        nass = int(rng.integers(1, 8))
        sets = [sorted(rng.choice(K, size=int(rng.integers(1, K)), replace=False).tolist()) for _ in range(nass)]
        write_ass(d / "a.ass", sets)
        variants.append((["-a", "a"], f"assoc {sets}"))
    if rng.random() < 0.4:
        variants.append((["-f"], "rescaling approximation"))
    if rng.random() < 0.3:
        variants.append((["--quick_mctest"], "quick mctest"))
    # a random handful of the switches that select tests, smoothing, sorting and fit parameters
    pool = [["--no_au"], ["--no_bp"], ["--no_pp"], ["--no_sh"], ["--smooth_cnt"], ["--no_sort"], ["--wls"],
            ["--th", f"{rng.uniform(-2, 2):.3f}"], ["--kappa", f"{rng.uniform(0.5, 1.5):.2f}"], ["--rmin", "0.5"], ["--rmax", "1.5"],
            ["-t", f"{rng.uniform(0.2, 2):.2f}"], ["--cieps", "0.05"], ["--dmin", "1"], ["--mleeps", "0.001"], ["--ppcoef", "0.5"]]
    if rng.random() < 0.3:   # AIC / BIC correction of the log-likelihoods (read_pdm, do_pdm: src/consel.c:675-744)
        with open(d / "pdm.vt", "w") as f:
            f.write(f"#!VEC:\n{K}\n" + " ".join(str(int(x)) for x in rng.integers(1, 30, size=K)) + "\n")
        pool.append(["--pdm", "pdm", "--aictype", str(int(rng.integers(1, 4))), "--ppn", str(n)])
    if rng.random() < 0.8:
        pick = rng.choice(len(pool), size=int(rng.integers(1, 4)), replace=False)
        extra = [x for i in sorted(pick) for x in pool[i]]
        if rng.random() < 0.3 and len(variants) > 1:
            extra = variants[1][0] + extra
        variants.append((extra, "switches " + " ".join(extra)))
    for vi, (extra, what) in enumerate(variants):
        for tag, fn in (("r", ref), ("o", ours)):
            rc, out = fn("consel", extra + ["-r", "-c", "x", f"{tag}{vi}"])
            if tag == "r":
                rrc, rtext = rc, out
            elif rc != rrc:
                bad.append(f"consel ({what}) exit status {rc} vs reference {rrc}: {desc}\n{out[-400:]}")
            elif protocol(rtext, f"r{vi}") != protocol(out, f"o{vi}"):
                STATS["stdout_diffs"] = STATS.get("stdout_diffs", 0) + 1
                if os.environ.get("EMU_STRESS_STDOUT"):
                    import difflib
                    bad.append(f"consel ({what}) stdout differs: {desc}\n" + "\n".join(list(difflib.unified_diff(
                        protocol(rtext, f"r{vi}"), protocol(out, f"o{vi}"), lineterm="", n=0))[:12]))
        STATS["variants"] += 1
        STATS["both_refused" if rrc != 0 else "both_ok"] += 0 if bad else 1
        if rrc != 0 or bad:
            continue
        stem = str(vi)
        for ext in ("rep", "cnt"):
            pa, pb = d / f"o{stem}.{ext}", d / f"r{stem}.{ext}"
            if pb.exists() and (not pa.exists() or pa.read_bytes() != pb.read_bytes()):
                bad.append(f"consel ({what}) .{ext} differs: {desc}")
        for ext in ("pv", "ci"):
            pa, pb = d / f"o{stem}.{ext}", d / f"r{stem}.{ext}"
            if pb.exists() and (not pa.exists() or not numbers_equal(pa, pb)):
                bad.append(f"consel ({what}) .{ext} differs: {desc}")
    # catrep (merge of .rep / .rmt files) on the reference's own files
    if orc.ref_bin("catrep") is not None:
        jobs = []
        if (d / "r0.rep").exists():
            jobs.append((["r0", "r0"], "rep"))
        if (d / "rmb.rmt").exists():
            jobs.append((["-m", "ref", "rmb"], "rmt"))
        for args, ext in jobs:
            a = subprocess.run([str(orc.ref_bin("catrep"))] + args + ["rcat"], cwd=d, capture_output=True, text=True)
            b = subprocess.run([str(ROOT / "consel_b200" / "bin" / "catrep")] + args + ["ocat"], cwd=d, capture_output=True, text=True)
            STATS["catrep"] = STATS.get("catrep", 0) + 1
            pa, pb = d / f"ocat.{ext}", d / f"rcat.{ext}"
            if a.returncode != b.returncode or pa.exists() != pb.exists() or (pb.exists() and pa.read_bytes() != pb.read_bytes()):
                bad.append(f"catrep {args} differs (exit {b.returncode} vs {a.returncode}): {desc}")
            for q in (pa, pb):
                if q.exists():
                    q.unlink()

    # catpv on the reference's .pv files: the table text must be the reference's, for random display switches
    real_catpv = ROOT / "consel_b200" / "bin" / "catpv"
    for vi in range(len(variants)):
        if not (d / f"r{vi}.pv").exists() or orc.ref_bin("catpv") is None:
            continue
        pool = [["-v"], ["-e"], ["-r"], ["-s", str(int(rng.integers(1, 10)))], ["-l"], ["--no_au"], ["--no_bp"], ["--no_pp"], ["--no_sh"]]
        pick = rng.choice(len(pool), size=int(rng.integers(0, 4)), replace=False)
        args = [x for i in sorted(pick) for x in pool[i]] + [f"r{vi}"]
        a = subprocess.run([str(orc.ref_bin("catpv"))] + args, cwd=d, capture_output=True, text=True)
        b = subprocess.run([str(real_catpv)] + args, cwd=d, capture_output=True, text=True)
        STATS["catpv"] = STATS.get("catpv", 0) + 1
        if a.returncode != b.returncode or protocol(a.stdout, "\0") != protocol(b.stdout, "\0"):
            bad.append(f"catpv {args} differs (exit {b.returncode} vs {a.returncode}): {desc}")

    # the reference's own .rep / .cnt of the plain run as INPUT: consel -R (do_repmode) and consel -C (do_cntmode)
    if (d / "r0.rep").exists() and (d / "r0.cnt").exists() and not bad:
        for flag, src in (("-R", "r0.rep"), ("-C", "r0.cnt")):
            shutil.copy(d / src, d / ("y" + Path(src).suffix))
            extra = [flag] + (["-c"] if flag == "-R" and rng.random() < 0.5 else [])
            rrc, rtext = ref("consel", extra + ["y", "rY" + flag[1]])
            orc_, otext = ours("consel", extra + ["y", "oY" + flag[1]])
            STATS["variants"] += 1
            if rrc != orc_:
                bad.append(f"consel {flag} exit status {orc_} vs reference {rrc}: {desc}\n{otext[-400:]}")
                continue
            STATS["both_refused" if rrc != 0 else "both_ok"] += 1
            if rrc != 0:
                continue
            for ext in ("cnt",):
                pa, pb = d / f"oY{flag[1]}.{ext}", d / f"rY{flag[1]}.{ext}"
                if pa.exists() != pb.exists() or (pb.exists() and pa.read_bytes() != pb.read_bytes()):
                    bad.append(f"consel {flag} .{ext} differs: {desc}")
            for ext in ("pv", "ci"):
                pa, pb = d / f"oY{flag[1]}.{ext}", d / f"rY{flag[1]}.{ext}"
                if pa.exists() != pb.exists() or (pb.exists() and not numbers_equal(pa, pb)):
                    bad.append(f"consel {flag} .{ext} differs: {desc}")
    return bad


def main() -> int:
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    work = Path(sys.argv[3] if len(sys.argv) > 3 else "/tmp/emu_stress")
    work.mkdir(parents=True, exist_ok=True)
    bindir = build_programs(work)
    rng = np.random.default_rng(seed)
    nbad = 0
    for c in range(cases):
        d = work / f"case{c}"
        if d.exists():
            shutil.rmtree(d)
        d.mkdir()
        bad = one_case(rng, bindir, d)
        for line in bad:
            print(f"[case {c}] MISMATCH {line}", flush=True)
        nbad += bool(bad)
        if not bad and not os.environ.get("EMU_STRESS_KEEP"):
            shutil.rmtree(d)
    print(f"emu_stress seed {seed}: {cases - nbad} of {cases} cases identical to the reference "
          f"({STATS['variants']} consel runs: {STATS['both_ok']} compared file by file, {STATS['both_refused']} refused by both programs alike, "
          f"{STATS.get('stdout_diffs', 0)} with a different stdout protocol; {STATS.get('catpv', 0)} catpv tables, {STATS.get('catrep', 0)} catrep merges compared)")
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
