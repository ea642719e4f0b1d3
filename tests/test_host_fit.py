"""The host-side AU fit (consel_b200/host/aufit.cpp) driven exactly like our consel
drives it — lockstep threshold iterations fed by row counts — but with the counts taken
on the CPU from the reference's golden sorted .rep rows, so the fit arithmetic is pinned
against the reference's .pv / .ci digits without a GPU."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "consel_b200" / "bin" / "libaufit_test.so"

pytestmark = pytest.mark.skipif(not SO.exists(), reason="host programs not built (run __graft_entry__.build())")


class Opts(C.Structure):
    _fields_ = [("fitmode", C.c_int), ("kappa", C.c_double), ("rrmin", C.c_double), ("rrmax", C.c_double),
                ("dfmin", C.c_int)]


@pytest.fixture(scope="module")
def fit():
    lib = C.CDLL(str(SO))
    for f in ("aufit_vc_new", "aufit_ci_new"):
        getattr(lib, f).restype = C.c_void_p
    for f in ("aufit_vc_x", "aufit_ci_x", "aufit_pochisq", "aufit_pnorm", "aufit_qnorm"):
        getattr(lib, f).restype = C.c_double
    lib.aufit_vc_new.argtypes = [C.c_double, C.c_double]
    lib.aufit_ci_new.argtypes = [C.c_double] * 4
    lib.aufit_pochisq.argtypes = [C.c_double, C.c_int]
    lib.aufit_pnorm.argtypes = [C.c_double]
    lib.aufit_qnorm.argtypes = [C.c_double]
    for f in ("aufit_vc_free", "aufit_vc_x", "aufit_vc_done", "aufit_vc_get", "aufit_ci_free", "aufit_ci_x",
              "aufit_ci_done", "aufit_ci_get"):
        getattr(lib, f).argtypes = [C.c_void_p] + ([C.POINTER(C.c_double)] if f.endswith("get") else [])
    for f in ("aufit_vc_step", "aufit_ci_step"):
        getattr(lib, f).argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.c_int, C.POINTER(Opts)]
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("name", ["g100", "tiny", "g100e"])
def test_au_pvalues_and_confidence_limits_match_reference(fit, orc, gold, name):
    rep = orc.read_rep(gold / f"{name}.rep")
    pv = orc.read_pv(gold / f"{name}.pv")
    ci_txt = orc.read_text_blocks(gold / f"{name}.ci")
    rr = np.ascontiguousarray(rep["rr"], dtype=np.float64)
    bb = np.ascontiguousarray(rep["bb"], dtype=np.float64)
    kk = len(rr)
    cm = len(rep["order"])
    i0 = orc.find_i0(rr)
    opts = Opts(1, 1.0, 0.0, 100.0, 0)
    got_pv = np.zeros((cm, 2)); got_se = np.zeros((cm, 2)); got_ax = np.zeros((cm, 7))
    for i in range(cm):
        rows = [rep["stats"][k][i] for k in range(kk)]
        h = fit.aufit_vc_new(float(rows[i0][int(bb[i0]) // 2]), 0.0)
        while not fit.aufit_vc_done(h):
            x = fit.aufit_vc_x(h)
            cnts = np.array([orc.cntdist(rows[k], x, 3) for k in range(kk)])
            fit.aufit_vc_step(h, _dp(cnts), _dp(rr), _dp(bb), kk, C.byref(opts))
        out = np.zeros(11)
        fit.aufit_vc_get(h, _dp(out))
        fit.aufit_vc_free(h)
        ret = int(out[10])
        got_pv[i] = out[0], out[2]
        got_se[i] = out[1], out[3]
        pf = fit.aufit_pochisq(out[4], int(out[5])) if ret == 0 else 0.0
        beta = out[7:10] if ret == 0 else np.zeros(3)
        got_ax[i] = pf, out[4], out[5], beta[0], beta[1], out[6], beta[2]
    tol = dict(rtol=6e-9, atol=1e-12)
    assert np.allclose(got_pv, pv["pv"][:, 6:8], **tol), (got_pv, pv["pv"][:, 6:8])
    assert np.allclose(got_se, pv["se"][:, 6:8], **tol)
    assert np.allclose(got_ax, pv["ax"], **tol), (got_ax, pv["ax"])

    # confidence limits: .ci holds ITEM, STAT, ALPHA then CI SE EI CI0 SE0 EI0 (cm x nalpha each)
    p = 0
    def vec():
        nonlocal p
        m = int(ci_txt[p]); p += 1
        a = np.array(ci_txt[p:p + m]); p += m
        return a
    def mat():
        nonlocal p
        m, n = int(ci_txt[p]), int(ci_txt[p + 1]); p += 2
        a = np.array(ci_txt[p:p + m * n]).reshape(m, n); p += m * n
        return a
    vec(); vec()
    alpha = vec()
    want = [mat() for _ in range(6)]
    got = [np.zeros((cm, len(alpha))) for _ in range(6)]
    for i in range(cm):
        rows = [rep["stats"][k][i] for k in range(kk)]
        for a, al in enumerate(alpha):
            for which, kappa in enumerate((1.0, -1.0)):
                h = fit.aufit_ci_new(float(rows[i0][0]), float(rows[i0][-1]), float(al), kappa)
                while not fit.aufit_ci_done(h):
                    x = fit.aufit_ci_x(h)
                    cnts = np.array([orc.cntdist(rows[k], x, 3) for k in range(kk)])
                    fit.aufit_ci_step(h, _dp(cnts), _dp(rr), _dp(bb), kk, C.byref(opts))
                out = np.zeros(4)
                fit.aufit_ci_get(h, _dp(out))
                fit.aufit_ci_free(h)
                for c in range(3):
                    got[3 * which + c][i, a] = out[c]
    for g, w in zip(got, want):
        assert np.allclose(g, w, rtol=6e-9, atol=1e-12, equal_nan=True), (g, w)


def test_distribution_functions(fit):
    assert fit.aufit_pnorm(0.0) == 0.5
    assert abs(fit.aufit_pnorm(1.959963984540054) - 0.975) < 1e-7
    assert abs(fit.aufit_qnorm(0.975) - 1.959963984540054) < 1e-6
    assert abs(fit.aufit_pochisq(3.841458820694124, 1) - 0.05) < 1e-5
    assert fit.aufit_pochisq(0.0, 3) == 1.0


def test_parallel_matrix_reader_equals_token_reader(tmp_path):
    """refio::read_mat switches to the parallel from_chars parser for large regular files; it must
    return exactly what the fscanf loop returns, leave the stream behind the matrix, and fall back
    (same values) when a token is outside its grammar."""
    import ctypes as C
    import numpy as np
    lib = C.CDLL(str(ROOT / "consel_b200" / "bin" / "libaufit_test.so"))
    lib.refio_test_read_mat.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_longlong,
                                        C.c_char_p, C.c_int]
    rng = np.random.default_rng(5)
    m, n = 40, 3000  # 120000 values: above the parallel threshold
    a = -np.exp(rng.uniform(-3, 4, size=(m, n)))
    a[0, 0], a[1, 7], a[2, 2] = 0.0, -np.inf, 1e-310

    def write(path, hexfirst=False):
        with open(path, "w") as f:
            f.write("# a comment before the header\n%d %d\n" % (m, n))
            for i in range(m):
                f.write("# row: %d\n" % i)
                for j in range(n):
                    tok = float(a[i, j]).hex() if (hexfirst and i == 3 and j == 5) else ("+%.17g" % a[i, j] if a[i, j] > 0 else "%.17g" % a[i, j])
                    f.write(tok + ("\n" if j % 5 == 4 else "  # c\n" if j % 997 == 0 else " "))
            f.write("\nTRAILER 1 2 3\n")

    for hexfirst in (False, True):
        p = tmp_path / ("m%d.mt" % hexfirst)
        write(p, hexfirst)
        out = np.zeros(m * n)
        mm, nn = C.c_int(0), C.c_int(0)
        word = C.create_string_buffer(64)
        assert lib.refio_test_read_mat(str(p).encode(), C.byref(mm), C.byref(nn), out.ctypes.data_as(C.POINTER(C.c_double)),
                                       m * n, word, 64) == 0
        assert (mm.value, nn.value) == (m, n)
        assert np.array_equal(out.view(np.uint64), a.ravel().view(np.uint64))
        assert word.value == b"TRAILER"

    # a file above 8 MB is also LOADED by all host threads (pread slices into one uninitialised buffer)
    m, n = 230, 3000
    a = -np.exp(rng.uniform(-3, 4, size=(m, n)))
    p = tmp_path / "big.mt"
    with open(p, "w") as f:
        f.write("%d %d\n" % (m, n))
        np.savetxt(f, a.reshape(-1, 6), fmt="%.17g")
        f.write("TRAILER\n")
    assert p.stat().st_size > (8 << 20)
    out = np.zeros(m * n)
    assert lib.refio_test_read_mat(str(p).encode(), C.byref(mm), C.byref(nn), out.ctypes.data_as(C.POINTER(C.c_double)),
                                   m * n, word, 64) == 0
    assert (mm.value, nn.value) == (m, n) and word.value == b"TRAILER"
    assert np.array_equal(out.view(np.uint64), a.ravel().view(np.uint64))


@pytest.mark.parametrize("case", ["g100", "g100e", "tiny", "g100fe"])
def test_consel_cnt_mode_cli(case, gold, tmp_path):
    """consel -C (do_cntmode / do_bootcnt, src/consel.c:1244-1283, 1540-1592): host arithmetic only in
    the reference too - the .pv is byte-equal to the reference's, degenerate items included."""
    import shutil
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "consel_b200" / "bin" / "consel"
    shutil.copy(gold / f"{case}.cnt", tmp_path / f"{case}.cnt")
    r = subprocess.run([str(exe), "-C", case, case + "C"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "# exit normally" in r.stdout, r.stdout + r.stderr
    assert (tmp_path / f"{case}C.pv").read_bytes() == (gold / f"{case}C.pv").read_bytes()
    r = subprocess.run([str(exe), "-C", "-R", case], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "only one of -C and -R can be specified" in r.stdout


def test_consel_cnt_mode_pmin(gold, tmp_path):
    """--pmin (rcalpval, src/consel.c:2727-2734): refit on fewer scales while the fit's p-value is too small."""
    import shutil
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "consel_b200" / "bin" / "consel"
    shutil.copy(gold / "g100.cnt", tmp_path / "g100.cnt")
    r = subprocess.run([str(exe), "-C", "--pmin", "0.5", "g100", "g100Cp"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "g100Cp.pv").read_bytes() == (gold / "g100Cp.pv").read_bytes()
    assert (gold / "g100Cp.pv").read_bytes() != (gold / "g100C.pv").read_bytes()


def test_fast_qnorm_is_the_reference_bisection_bit_for_bit(fit):
    """qnorm (src/rand.c:465-493) is a 40-step bisection on a 3e-7-accurate pnorm; our qnorm finds the same
    final cell with ~4 evaluations.  The property its coarse step rests on (the computed pnorm is
    non-decreasing on the 2^20-cell grid of the first 20 halvings) is checked exhaustively, and the two
    functions are compared bit for bit on 1.2 million arguments: uniform, both tails, exact pnorm values,
    the neighbourhoods of z = 0 and of the series / continued-fraction switch."""
    fit.aufit_qnorm_grid_monotone.restype = C.c_int
    fit.aufit_qnorm_mismatches.restype = C.c_longlong
    fit.aufit_qnorm_mismatches.argtypes = [C.c_longlong, C.c_ulonglong]
    assert fit.aufit_qnorm_grid_monotone() == 1
    assert fit.aufit_qnorm_mismatches(1200000, 20261020) == 0
