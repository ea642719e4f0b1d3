"""ctypes front-end of oracle.c plus byte-level parsers of the reference's files.

TEST INFRASTRUCTURE ONLY — see the header of oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

File formats follow /root/reference/src/misc.c:525-645 (binary) and
src/makermt.c:178-205, src/consel.c:1713-1774 (layout of .rmt/.rep/.cnt).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
_LIB = None


def build(ref: bool = True) -> None:
    """Compile liboracle.so (and the reference binaries when /root/reference exists)."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", str(HERE)] + targets, check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = REF_DIR / "liboracle.so"
        if not so.exists():
            build(ref=False)
        _LIB = C.CDLL(str(so))
        _LIB.orc_cntdist.restype = C.c_double
        _LIB.orc_bp.restype = C.c_double
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def ref_bin(name: str):
    p = REF_DIR / name
    return p if p.exists() else None


# ----------------------------------------------------------------- MT19937
def mt_stream(seed: int, skip: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint32)
    lib().orc_mt_stream(C.c_uint64(seed), C.c_uint64(skip), C.c_uint64(count), _p(out, C.c_uint32))
    return out


def mt_state_after_blocks(seed: int, blocks: int) -> np.ndarray:
    out = np.empty(624, dtype=np.uint32)
    lib().orc_mt_state_after_blocks(C.c_uint64(seed), C.c_uint64(blocks), _p(out, C.c_uint32))
    return out


def site_indices(y: np.ndarray, n: int) -> np.ndarray:
    """src/mt19937.c:195 + src/makermt.c:55: (int)((double)y * 2^-32 * n) with the clamp."""
    k = (y.astype(np.float64) * (1.0 / 4294967296.0) * np.float64(n)).astype(np.int64)
    return np.clip(k, 0, n - 1)


# --------------------------------------------------------------- makermt side
def scale_setup(L: np.ndarray, rr):
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    rr = np.ascontiguousarray(rr, dtype=np.float64)
    S = len(rr)
    datvec = np.empty(K)
    bn = np.empty(S, dtype=np.int32)
    rr1 = np.empty(S)
    lib().orc_scale_setup(_p(L, C.c_double), K, n, S, _p(rr, C.c_double), _p(datvec, C.c_double),
                          _p(bn, C.c_int), _p(rr1, C.c_double))
    return datvec, bn, rr1


def makermt(L: np.ndarray, bn, bb, seed: int):
    """All scales with one global stream; returns list of K x bb[s] arrays."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    bn = np.ascontiguousarray(bn, dtype=np.int32)
    bb = np.ascontiguousarray(bb, dtype=np.int32)
    out = np.empty(int(K * bb.sum()), dtype=np.float64)
    lib().orc_makermt(_p(L, C.c_double), K, n, len(bn), _p(bn, C.c_int), _p(bb, C.c_int),
                      C.c_uint64(seed), _p(out, C.c_double))
    reps, off = [], 0
    for b in bb:
        reps.append(out[off:off + K * b].reshape(K, b))
        off += K * b
    return reps


def scaleboot_counts(n: int, bn, bb, seed: int):
    """Site-count matrices c[s][b, j] of the seeded reference stream (int32)."""
    res = []
    skip = 0
    for s in range(len(bn)):
        y = mt_stream(seed, skip, int(bn[s]) * int(bb[s]))
        skip += int(bn[s]) * int(bb[s])
        idx = site_indices(y, n).reshape(int(bb[s]), int(bn[s]))
        c = np.zeros((int(bb[s]), n), dtype=np.int32)
        rows = np.repeat(np.arange(int(bb[s])), int(bn[s]))
        np.add.at(c, (rows, idx.ravel()), 1)
        res.append(c)
    return res


def contract_counts(L: np.ndarray, counts: np.ndarray, r: float) -> np.ndarray:
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    bb = counts.shape[0]
    rep = np.empty((K, bb))
    lib().orc_contract_counts(_p(L, C.c_double), K, n, _p(counts, C.c_int32), bb, C.c_double(r),
                              _p(rep, C.c_double))
    return rep


# ---------------------------------------------------------------- consel side
def identity_assoc(K: int):
    return np.arange(K + 1, dtype=np.int32), np.arange(K, dtype=np.int32)


def repminmaxs(rep: np.ndarray, ass_ptr, ass_idx) -> np.ndarray:
    rep = np.ascontiguousarray(rep, dtype=np.float64)
    K, B = rep.shape
    ass_ptr = np.ascontiguousarray(ass_ptr, dtype=np.int32)
    ass_idx = np.ascontiguousarray(ass_idx, dtype=np.int32)
    cm = len(ass_ptr) - 1
    stat = np.empty((cm, B))
    lib().orc_repminmaxs(_p(rep, C.c_double), K, B, cm, _p(ass_ptr, C.c_int), _p(ass_idx, C.c_int),
                         _p(stat, C.c_double))
    return stat


def repminmaxs_rescaled(rep: np.ndarray, ass_ptr, ass_idx, rs: float) -> np.ndarray:
    """consel -f: genrep's rescaling branch (src/consel.c:762-769) + repminmaxsrs (:1961-1984).
    Row mean = left-to-right sum / B (cumsum is that chain), x' = rs*(x - mean) + mean with
    every operation rounded on its own, then the ordinary statistics."""
    rep = np.ascontiguousarray(rep, dtype=np.float64)
    mean = (np.cumsum(rep, axis=1)[:, -1] / rep.shape[1])[:, None]
    d = rep - mean
    d = rs * d
    return repminmaxs(d + mean, ass_ptr, ass_idx)


def obs_stat(datvec, ass_ptr, ass_idx) -> np.ndarray:
    datvec = np.ascontiguousarray(datvec, dtype=np.float64)
    ass_ptr = np.ascontiguousarray(ass_ptr, dtype=np.int32)
    ass_idx = np.ascontiguousarray(ass_idx, dtype=np.int32)
    cm = len(ass_ptr) - 1
    obs = np.empty(cm)
    lib().orc_obs_stat(_p(datvec, C.c_double), len(datvec), cm, _p(ass_ptr, C.c_int),
                       _p(ass_idx, C.c_int), _p(obs, C.c_double))
    return obs


def cntdist(sorted_row: np.ndarray, t: float, mode: int) -> float:
    v = np.ascontiguousarray(sorted_row, dtype=np.float64)
    return float(lib().orc_cntdist(_p(v, C.c_double), len(v), C.c_double(t), mode))


def mctest(rep0: np.ndarray, datvec, ass_ptr, ass_idx, varadd: float = 1.0):
    rep0 = np.ascontiguousarray(rep0, dtype=np.float64)
    K, B = rep0.shape
    datvec = np.ascontiguousarray(datvec, dtype=np.float64)
    ass_ptr = np.ascontiguousarray(ass_ptr, dtype=np.int32)
    ass_idx = np.ascontiguousarray(ass_idx, dtype=np.int32)
    cm = len(ass_ptr) - 1
    outs = [np.empty(cm) for _ in range(4)]
    lib().orc_mctest(_p(rep0, C.c_double), _p(datvec, C.c_double), K, B, cm, _p(ass_ptr, C.c_int),
                     _p(ass_idx, C.c_int), C.c_double(varadd), *[_p(o, C.c_double) for o in outs])
    return tuple(outs)  # kh, sh, wkh, wsh


def find_i0(rr) -> int:
    """src/consel.c:1778-1789: first index minimising |r-1|."""
    return int(np.argmin(np.abs(np.asarray(rr, dtype=np.float64) - 1.0)))


# ------------------------------------------------------------------- file I/O
def _rd_i(buf, off):
    return int(np.frombuffer(buf, dtype=np.int32, count=1, offset=off)[0]), off + 4


def _rd_vec(buf, off, dtype):
    m, off = _rd_i(buf, off)
    a = np.frombuffer(buf, dtype=dtype, count=m, offset=off).copy()
    return a, off + a.nbytes


def read_rmt(path):
    """-> dict(datvec, rr, bb, reps=[K x bb[s]])   (src/makermt.c:178-183,205)"""
    buf = Path(path).read_bytes()
    off = 0
    datvec, off = _rd_vec(buf, off, np.float64)
    rr, off = _rd_vec(buf, off, np.float64)
    bb, off = _rd_vec(buf, off, np.int32)
    S, off = _rd_i(buf, off)
    assert S == len(rr) == len(bb)
    reps = []
    for s in range(S):
        K, off = _rd_i(buf, off)
        B, off = _rd_i(buf, off)
        a = np.frombuffer(buf, dtype=np.float64, count=K * B, offset=off).reshape(K, B).copy()
        off += a.nbytes
        reps.append(a)
    assert off == len(buf)
    return dict(datvec=datvec, rr=rr, bb=bb, reps=reps)


def write_rmt(path, datvec, rr1, bb, reps):
    with open(path, "wb") as f:
        def wi(x): f.write(np.int32(x).tobytes())
        wi(len(datvec)); f.write(np.asarray(datvec, np.float64).tobytes())
        wi(len(rr1)); f.write(np.asarray(rr1, np.float64).tobytes())
        wi(len(bb)); f.write(np.asarray(bb, np.int32).tobytes())
        wi(len(bb))
        for r in reps:
            wi(r.shape[0]); wi(r.shape[1]); f.write(np.ascontiguousarray(r, np.float64).tobytes())


def read_rep(path):
    """-> dict(order, obs, rr, bb, stats=[cm x bb[s]] sorted rows)  (src/consel.c:1713-1740)"""
    buf = Path(path).read_bytes()
    off = 0
    order, off = _rd_vec(buf, off, np.int32)
    obs, off = _rd_vec(buf, off, np.float64)
    rr, off = _rd_vec(buf, off, np.float64)
    bb, off = _rd_vec(buf, off, np.int32)
    S, off = _rd_i(buf, off)
    stats = []
    for s in range(S):
        cm, off = _rd_i(buf, off)
        B, off = _rd_i(buf, off)
        a = np.frombuffer(buf, dtype=np.float64, count=cm * B, offset=off).reshape(cm, B).copy()
        off += a.nbytes
        stats.append(a)
    return dict(order=order, obs=obs, rr=rr, bb=bb, stats=stats)


def read_text_blocks(path):
    """Parse the reference's ascii vec/mat files (.pv/.ci/.cnt/.vt): returns the
    list of numeric blocks in file order, each a 1-D or 2-D float array
    (reader rules: src/misc.c:174-193, 334-404)."""
    toks = []
    for line in Path(path).read_text().splitlines():
        line = line.split("#", 1)[0]
        toks += line.split()
    return [float(t) for t in toks]


def read_pv(path):
    """.pv layout: src/consel.c:1611-1683."""
    t = read_text_blocks(path)
    p = 0

    def geti():
        nonlocal p
        v = int(t[p]); p += 1
        return v

    def vec():
        nonlocal p
        m = geti()
        a = np.array(t[p:p + m]); p += m
        return a

    def mat():
        nonlocal p
        m, n = geti(), geti()
        a = np.array(t[p:p + m * n]).reshape(m, n); p += m * n
        return a

    fid = geti()
    item = vec().astype(int)
    stat = vec()
    bit = geti()
    pv, se, ax = mat(), mat(), mat()
    return dict(id=fid, item=item, stat=stat, bit=bit, pv=pv, se=se, ax=ax)


def read_cnt(path):
    """.cnt layout: src/consel.c:1754-1774."""
    t = read_text_blocks(path)
    p = 0

    def geti():
        nonlocal p
        v = int(t[p]); p += 1
        return v

    def vec():
        nonlocal p
        m = geti()
        a = np.array(t[p:p + m]); p += m
        return a

    def mat():
        nonlocal p
        m, n = geti(), geti()
        a = np.array(t[p:p + m * n]).reshape(m, n); p += m * n
        return a

    item = vec().astype(int)
    stat = vec()
    return dict(item=item, stat=stat, r=mat(), b=mat(), cnt=mat())


def read_lnf(path) -> np.ndarray:
    """PAML .lnf reader (src/freadmat.c:110-154): header `m n p`, then per tree an
    id line and p pattern lines of which only the first three columns
    (`idx count lnL`) are used; each pattern's lnL is repeated `count` times."""
    lines = [ln.split() for ln in Path(path).read_text().splitlines()]
    lines = [t for t in lines if t]
    m, n, npat = (int(x) for x in lines[0][:3])
    L = np.empty((m, n))
    p = 1
    for i in range(m):
        assert int(lines[p][0]) == i + 1, "wrong row index"
        p += 1
        col = 0
        for j in range(npat):
            t = lines[p]; p += 1
            assert int(t[0]) == j + 1, "wrong column index"
            cnt = int(t[1])
            L[i, col:col + cnt] = float(t[2])
            col += cnt
        assert col == n, (col, n)
    return L


def write_mt(path, L: np.ndarray):
    """.mt text matrix (src/misc.c:385-404 reader) with 17 significant digits."""
    K, n = L.shape
    with open(path, "w") as f:
        f.write(f"{K} {n}\n")
        for k in range(K):
            f.write(" ".join(repr(float(x)) for x in L[k]))
            f.write("\n")


def write_pa(path, rr, bb):
    with open(path, "w") as f:
        f.write(f"{len(rr)}\n" + " ".join(repr(float(r)) for r in rr) + "\n")
        f.write(f"{len(bb)}\n" + " ".join(str(int(b)) for b in bb) + "\n")


def run_ref(prog: str, args, cwd, check=True) -> str:
    exe = ref_bin(prog)
    if exe is None:
        raise FileNotFoundError(f"oracle/_ref/{prog} not built")
    r = subprocess.run([str(exe)] + [str(a) for a in args], cwd=str(cwd), capture_output=True, text=True)
    if check and r.returncode != 0:
        raise RuntimeError(f"{prog} failed: {r.stdout}\n{r.stderr}")
    return r.stdout


def read_mt(path) -> np.ndarray:
    t = read_text_blocks(path)
    m, n = int(t[0]), int(t[1])
    return np.array(t[2:2 + m * n]).reshape(m, n)


def read_pa(path):
    t = read_text_blocks(path)
    k = int(t[0])
    rr = np.array(t[1:1 + k])
    k2 = int(t[1 + k])
    bb = np.array(t[2 + k:2 + k + k2]).astype(np.int32)
    return rr, bb


def read_ass(path):
    """.ass: count, then per association an ivec of 0-based tree ids (src/consel.c:467-480).
    Returns CSR (ptr, idx)."""
    t = [int(x) for x in read_text_blocks(path)]
    cm = t[0]
    p = 1
    ptr, idx = [0], []
    for _ in range(cm):
        m = t[p]; p += 1
        idx += t[p:p + m]; p += m
        ptr.append(len(idx))
    return np.array(ptr, dtype=np.int32), np.array(idx, dtype=np.int32)


def order_items(datvec, ass_ptr, ass_idx):
    """Item order of read_assoc (src/consel.c:499-517) for the tie-free case:
    ascending by the sorted vector of per-member observed statistics, compared
    lexicographically after a length comparison (compdvec :1986-2000).
    Returns (orderv, obsvec, ptr_sorted, idx_sorted)."""
    datvec = np.asarray(datvec, dtype=np.float64)
    cm = len(ass_ptr) - 1
    keys = []
    for e in range(cm):
        a = ass_idx[ass_ptr[e]:ass_ptr[e + 1]]
        mask = np.ones(len(datvec), bool); mask[a] = False
        mx = datvec[mask].max() if mask.any() else -1.0e30
        v = np.sort(mx - datvec[a])
        keys.append(v)

    import functools

    def cmp(i, j):
        x, y = keys[i], keys[j]
        out = -1 if len(x) < len(y) else (1 if len(x) > len(y) else 0)
        for a, b in zip(x, y):
            if a < b: return -1
            if a > b: return 1
        return out

    order = sorted(range(cm), key=functools.cmp_to_key(cmp))
    obs = np.array([keys[i][0] for i in order])
    ptr, idx = [0], []
    for i in order:
        idx += list(ass_idx[ass_ptr[i]:ass_ptr[i + 1]])
        ptr.append(len(idx))
    return (np.array(order, dtype=np.int32), obs, np.array(ptr, dtype=np.int32), np.array(idx, dtype=np.int32))


def exact_rounded_contract(L: np.ndarray, counts: np.ndarray, r: float) -> np.ndarray:
    """Big-integer restatement of the fixed-point path (CGPU_SUM_INT8): the replicate sums
    sum_j c[b,j]*L[k,j] are formed EXACTLY (Python integers), rounded once to fp64, then
    divided by r like src/makermt.c:69.  Small cases only."""
    from fractions import Fraction
    L = np.asarray(L, dtype=np.float64)
    K, n = L.shape
    # common scale: every double is an integer multiple of 2^-1074
    fr = [[Fraction(float(x)) for x in row] for row in L]
    den = 1
    for row in fr:
        for f in row:
            den = max(den, f.denominator)
    M = np.array([[int(f * den) for f in row] for row in fr], dtype=object)  # exact: den is a power of two
    C = np.asarray(counts).astype(object)
    T = C.dot(M.T)  # (B, K) exact integers
    out = np.empty((K, C.shape[0]))
    rr = np.float64(r)
    for b in range(C.shape[0]):
        for k in range(K):
            x = np.float64(float(Fraction(int(T[b, k]), den)))  # one correctly rounded conversion
            out[k, b] = x / rr
    return out
