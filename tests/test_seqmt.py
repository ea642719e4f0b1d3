"""seqmt (host converter on makermt's input side) and, through it, the five
foreign site-likelihood readers our makermt shares (refio.cpp): output .mt byte-equal and
stdout equal to the unmodified reference's seqmt (src/seqmt.c, src/freadmat.c) compiled under
oracle/_ref, on small files written here in each program's layout.  Host-only."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "consel_b200" / "bin" / "seqmt"


def sample(fmt: str, path: Path):
    rng = np.random.Generator(np.random.PCG64(11))
    m, n = 3, 7
    L = -rng.exponential(3.0, size=(m, n)) - 0.5
    if fmt == "molphy":      # "m n" + rest of line ignored, then the values
        path.write_text(f"{m} {n} protml\n" + "\n".join(" ".join(f"{v:.6f}" for v in r) for r in L) + "\n")
    elif fmt == "puzzle":    # "m n" line, then a name word and n values per tree
        path.write_text(f"{m} {n}\n" + "\n".join(f"tr{i+1} " + " ".join(f"{v:.5f}" for v in r) for i, r in enumerate(L)) + "\n")
    elif fmt == "paup":      # PAUP 4b10 "Tree -lnL Site -lnL" layout, one gap line (dropped)
        s = "Tree\t-lnL\tSite\t-lnL\n"
        for i, r in enumerate(L):
            s += f"{i+1}\t{-r.sum():.5f}\n"
            for j, v in enumerate(r):
                s += f"\t\t{j+1}\t{-v:.8f}\n"
            if i == 1:
                s += f"\t\t{n+1}\t-\n"
        path.write_text(s)
    elif fmt == "paup1":     # old PAUP layout: "Single-site ..." line per tree, then "site value" pairs
        s = "PAUP old style output\n"
        for i, r in enumerate(L):
            s += f"Single-site -ln likelihoods for tree {i+1}:\n" + "".join(f"{j+1}\t{v:.6f}\n" for j, v in enumerate(r)) + "\n"
        path.write_text(s)
    elif fmt == "phyml":     # blocks headed by "Site   P(D|M)", probabilities (log taken)
        s = "Note : P(D|M) is the probability of site D given the model M\n\n"
        for i, r in enumerate(L):
            s += "Site   P(D|M)   P(D|M,r=1)\n" + "".join(f"{j+1:<6d} {np.exp(v):.6e} {np.exp(v) / 2:.6e}\n" for j, v in enumerate(r)) + "\n"
        path.write_text(s)
    return L


CASES = [("molphy", ".lls"), ("puzzle", ".sitelh"), ("paup", ".txt"), ("paup1", ".txt"), ("phyml", ".txt")]


def body(out: str) -> str:
    return out.split("\n", 1)[1] if "\n" in out else ""


def run(exe, args, cwd, stdin=None):
    return subprocess.run([str(exe)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, input=stdin)


@pytest.mark.parametrize("fmt,ext", CASES)
def test_seqmt_foreign_formats_equal_reference_binary(fmt, ext, orc, tmp_path):
    if orc.ref_bin("seqmt") is None:
        pytest.skip("oracle/_ref/seqmt not built")
    flag = "--paup" if fmt == "paup1" else "--" + fmt
    for d in "ab":
        (tmp_path / d).mkdir()
        sample(fmt, tmp_path / d / ("in" + ext))
    ours = run(EXE, [flag, "in"], tmp_path / "a")
    ref = run(orc.ref_bin("seqmt"), [flag, "in"], tmp_path / "b")
    assert ours.returncode == ref.returncode == 0
    assert body(ours.stdout) == body(ref.stdout)
    assert (tmp_path / "a" / "in.mt").read_bytes() == (tmp_path / "b" / "in.mt").read_bytes()
    L = sample(fmt, tmp_path / "again")
    got = orc.read_mt(tmp_path / "a" / "in.mt")
    assert got.shape == L.shape and np.allclose(got, L, rtol=2e-5)  # fixtures carry 5-8 digits


def test_seqmt_paml_mt_and_stdio(gold, orc, tmp_path):
    if orc.ref_bin("seqmt") is None:
        pytest.skip("oracle/_ref/seqmt not built")
    for d in "ab":
        (tmp_path / d).mkdir()
    a = run(EXE, ["--paml", gold / "mam15.lnf", "mam15"], tmp_path / "a")
    b = run(orc.ref_bin("seqmt"), ["--paml", gold / "mam15.lnf", "mam15"], tmp_path / "b")
    assert a.returncode == b.returncode == 0 and body(a.stdout) == body(b.stdout)
    assert (tmp_path / "a" / "mam15.mt").read_bytes() == (tmp_path / "b" / "mam15.mt").read_bytes()
    txt = (gold / "tiny.mt").read_text()  # stdin -> stdout
    a, b = run(EXE, [], tmp_path, stdin=txt), run(orc.ref_bin("seqmt"), [], tmp_path, stdin=txt)
    assert a.returncode == b.returncode == 0 and body(a.stdout) == body(b.stdout)
    assert run(EXE, ["--nope"], tmp_path).returncode == 1


def _write_lnf(path: Path, m: int, p: int, seed: int, comment_at=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    reps = rng.integers(1, 5, size=p)
    n = int(reps.sum())
    s = f"  {m}  {n}  {p}\n\n"
    for i in range(m):
        s += f"\n {i + 1}\n\n"
        v = -rng.exponential(4.0, size=p) - 1.0
        for j in range(p):
            if comment_at == (i, j):
                s += "# a comment line in the middle of a tree block\n"
            s += f"{j + 1:6d}{reps[j]:6d}{v[j]:16.10f}{rng.uniform(0, 99):9.2f}  ACGT\n"
    path.write_text(s)
    return n


@pytest.mark.parametrize("comment_at", [None, (17, 400)])
def test_parallel_paml_reader_equals_reference_binary(orc, tmp_path, comment_at):
    """a .lnf large enough for the threaded tree-block parser (m*p >= 2^15); with a comment line inside a
    block the parser must step aside and the serial loop reproduces the reference's reading"""
    if orc.ref_bin("seqmt") is None:
        pytest.skip("oracle/_ref/seqmt not built")
    for d in ("ours", "ref"):
        (tmp_path / d).mkdir()
        _write_lnf(tmp_path / d / "big.lnf", 40, 900, 3, comment_at)
    a = run(EXE, ["--paml", "big"], tmp_path / "ours")
    b = run(orc.ref_bin("seqmt"), ["--paml", "big"], tmp_path / "ref")
    assert a.returncode == b.returncode
    assert body(a.stdout) == body(b.stdout)
    if b.returncode == 0:
        assert (tmp_path / "ours" / "big.mt").read_bytes() == (tmp_path / "ref" / "big.mt").read_bytes()
