"""ctypes binding of libconselgpu.so (include/consel_gpu.h) + the small amount of
host logic the reference keeps on the host (scale set-up, src/makermt.c:159-168).

Nothing here computes replicates or statistics: every such call goes through the
C ABI into CUDA kernels and raises CgpuError when that is impossible.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
HEADER = ROOT / "include" / "consel_gpu.h"

RNG_MT19937, RNG_PHILOX = 0, 1
SUM_STRICT, SUM_FAST, SUM_INT8 = 0, 1, 2

_lib = None


class CgpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class RowSum(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("cnt", "lo", "hi", "min1", "min2", "max1", "max2")]


ROWSUM_DTYPE = np.dtype([(k, np.float64) for k in ("cnt", "lo", "hi", "min1", "min2", "max1", "max2")])


def lib_path() -> Path:
    return PKG / "libconselgpu.so"


def load() -> C.CDLL:
    """dlopen the in-tree library; never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise CgpuError(-2, f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(consel_b200 has no CPU fallback)")
    lib = C.CDLL(str(p))
    lib.cgpu_last_error.restype = C.c_char_p
    lib.cgpu_launch_count.restype = C.c_longlong
    if hasattr(lib, "cgpu_smooth_count"):
        lib.cgpu_smooth_count.restype = C.c_double
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    """every function the public header declares (used by the CPU-side ABI test)"""
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cgpu_[a-z0-9_]+)\s*\(", txt)))


def _check(rc: int) -> None:
    if rc != 0:
        raise CgpuError(rc, load().cgpu_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def init(device: int = -1) -> None:
    _check(load().cgpu_init(device))


def set_shard(rank: int, world: int) -> None:
    _check(load().cgpu_set_shard(rank, world))


def shard_range(bb: int) -> tuple[int, int]:
    b0, b1 = C.c_int(), C.c_int()
    _check(load().cgpu_shard_range(int(bb), C.byref(b0), C.byref(b1)))
    return b0.value, b1.value


def set_option(name: str, value: int) -> None:
    _check(load().cgpu_set_option(name.encode(), int(value)))


def launch_count() -> int:
    return int(load().cgpu_launch_count())


def last_timing() -> dict:
    ms = (C.c_double * 12)()
    _check(load().cgpu_last_timing_ex(ms, 12))
    keys = ("upload", "rng", "resample", "contract", "download", "total", "gemm_i8", "fx_combine", "mt_jump", "mt_draw",
            "n_gemm", "n_hist")
    return dict(zip(keys, [float(x) for x in ms]))


def scale_setup(L: np.ndarray, rr):
    """Host part of genrmt (src/makermt.c:159-168): datvec by sequential row sums,
    bn = (int)(rr*n) with the fp64 product truncated, rr1 = bn/n."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    datvec = np.zeros(K)
    # sequential left-to-right sums like the reference loop (np.sum pairs terms): host helper
    load().cgpu_host_rowsums(_dp(L), K, n, _dp(datvec))
    rr = np.asarray(rr, dtype=np.float64)
    bn = (rr * np.float64(n)).astype(np.int32)
    rr1 = bn.astype(np.float64) / np.float64(n)
    return datvec, bn, rr1


def _ptr_array(arrs):
    return (C.POINTER(C.c_double) * len(arrs))(*[_dp(a) for a in arrs])


def scaleboot_all(L: np.ndarray, bn, bb, seed: int, stream_offset: int = 0, rng_mode: int = RNG_MT19937,
                  sum_mode: int = SUM_STRICT, out=None):
    """cgpu_scaleboot_all: host matrix in, host replicate blocks out: list of K x bb[s]
    arrays (K x this rank's columns when sharded)."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    bn = np.ascontiguousarray(bn, dtype=np.int32)
    bb = np.ascontiguousarray(bb, dtype=np.int32)
    if out is None:
        out = [np.zeros((K, shard_range(int(b))[1] - shard_range(int(b))[0]), dtype=np.float64) for b in bb]
    _check(load().cgpu_scaleboot_all(_dp(L), K, n, len(bn), _ip(bn), _ip(bb), C.c_ulong(seed),
                                     C.c_ulonglong(stream_offset), rng_mode, sum_mode, _ptr_array(out)))
    return out


class Matrix:
    """Device-resident site-log-likelihood matrix (cgpu_matrix)."""

    def __init__(self, L: np.ndarray):
        L = np.ascontiguousarray(L, dtype=np.float64)
        self.K, self.n = L.shape
        self._h = C.c_void_p()
        _check(load().cgpu_matrix_upload(_dp(L), self.K, self.n, C.byref(self._h)))

    def scaleboot(self, bn, bb, seed: int, stream_offset: int = 0, rng_mode: int = RNG_MT19937,
                  sum_mode: int = SUM_STRICT) -> "Reps":
        bn = np.ascontiguousarray(bn, dtype=np.int32)
        bb = np.ascontiguousarray(bb, dtype=np.int32)
        h = C.c_void_p()
        _check(load().cgpu_scaleboot_dev(self._h, len(bn), _ip(bn), _ip(bb), C.c_ulong(seed),
                                         C.c_ulonglong(stream_offset), rng_mode, sum_mode, C.byref(h)))
        return Reps(h, self.K, bb)

    def info(self) -> dict:
        K, n, NP, S, F, P = (C.c_int() for _ in range(6))
        mp = C.c_longlong()
        _check(load().cgpu_matrix_info(self._h, C.byref(K), C.byref(n), C.byref(NP), C.byref(S), C.byref(F), C.byref(mp), C.byref(P)))
        return dict(K=K.value, n=n.value, NP=NP.value, slices=S.value, frac_bits=F.value, mpad=mp.value, patterns=P.value)

    def free(self):
        if self._h:
            load().cgpu_matrix_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Reps:
    """Device-resident replicate blocks and the statistics derived from them (cgpu_reps)."""

    def __init__(self, handle, K: int, bb):
        self._h = handle
        self.K = K
        self.bb = np.ascontiguousarray(bb, dtype=np.int32)
        self.S = len(self.bb)
        self.cm = 0

    @classmethod
    def upload(cls, reps) -> "Reps":
        reps = [np.ascontiguousarray(r, dtype=np.float64) for r in reps]
        K = reps[0].shape[0]
        bb = np.array([r.shape[1] for r in reps], dtype=np.int32)
        h = C.c_void_p()
        _check(load().cgpu_reps_upload(K, len(reps), _ip(bb), _ptr_array(reps), C.byref(h)))
        return cls(h, K, bb)

    @classmethod
    def upload_ptrs(cls, K: int, bb, ptrs) -> "Reps":
        """Blocks given by raw addresses (host, or device memory of this process — the copy
        kind is inferred): K x bb[s] doubles each, row-major."""
        bb = np.ascontiguousarray(bb, dtype=np.int32)
        arr = (C.c_void_p * len(bb))(*[C.c_void_p(int(p)) for p in ptrs])
        h = C.c_void_p()
        _check(load().cgpu_reps_upload(int(K), len(bb), _ip(bb), C.cast(arr, C.POINTER(C.POINTER(C.c_double))), C.byref(h)))
        return cls(h, int(K), bb)

    @classmethod
    def upload_stats(cls, stats) -> "Reps":
        """A handle over statistic blocks that already exist (the payload of a `.rep` file,
        consel -R): cm x bb[s] each.  No replicate blocks behind it: stats_build / mctest raise."""
        stats = [np.ascontiguousarray(t, dtype=np.float64) for t in stats]
        cm = stats[0].shape[0]
        bb = np.array([t.shape[1] for t in stats], dtype=np.int32)
        h = C.c_void_p()
        _check(load().cgpu_stats_upload(cm, len(stats), _ip(bb), _ptr_array(stats), C.byref(h)))
        r = cls(h, 0, bb)
        r.cm = cm
        return r

    def _local_cols(self):
        return [shard_range(int(b))[1] - shard_range(int(b))[0] for b in self.bb]

    def download(self, out=None):
        if out is None:
            out = [np.zeros((self.K, nb), dtype=np.float64) for nb in self._local_cols()]
        _check(load().cgpu_reps_download(self._h, _ptr_array(out)))
        return out

    def download_cols(self, s: int, col0: int, ncols: int) -> np.ndarray:
        out = np.zeros((self.K, ncols), dtype=np.float64)
        _check(load().cgpu_reps_download_cols(self._h, int(s), int(col0), int(ncols), _dp(out)))
        return out

    def stats_build(self, ass_ptr, ass_idx) -> None:
        ass_ptr = np.ascontiguousarray(ass_ptr, dtype=np.int32)
        ass_idx = np.ascontiguousarray(ass_idx, dtype=np.int32)
        self.cm = len(ass_ptr) - 1
        _check(load().cgpu_stats_build(self._h, self.cm, _ip(ass_ptr), _ip(ass_idx)))

    def stats_build_rescaled(self, i0: int, rs, ass_ptr, ass_idx) -> "Reps":
        """consel -f: statistics of len(rs) target scales derived from scale i0 (cgpu_stats_build_rescaled)."""
        rs = np.ascontiguousarray(rs, dtype=np.float64)
        ass_ptr = np.ascontiguousarray(ass_ptr, dtype=np.int32)
        ass_idx = np.ascontiguousarray(ass_idx, dtype=np.int32)
        h = C.c_void_p()
        _check(load().cgpu_stats_build_rescaled(self._h, int(i0), len(rs), _dp(rs), len(ass_ptr) - 1, _ip(ass_ptr),
                                                _ip(ass_idx), C.byref(h)))
        r = Reps(h, 0, np.full(len(rs), int(self.bb[i0]), dtype=np.int32))
        r.cm = len(ass_ptr) - 1
        return r

    def stats_download(self, sorted_rows: bool = False):
        out = [np.zeros((self.cm, nb), dtype=np.float64) for nb in self._local_cols()]
        _check(load().cgpu_stats_download(self._h, int(sorted_rows), _ptr_array(out)))
        return out

    def count_query(self, thr: np.ndarray) -> np.ndarray:
        """thr: (cm, nq) thresholds -> structured array (cm, S, nq) of row summaries."""
        thr = np.ascontiguousarray(thr, dtype=np.float64)
        thr = thr.reshape(thr.shape[0], -1)
        nq = thr.shape[1]
        if self.cm and thr.shape[0] != self.cm:
            raise CgpuError(-1, f"count_query: {thr.shape[0]} threshold rows for {self.cm} items")
        out = np.zeros((thr.shape[0], self.S, nq), dtype=ROWSUM_DTYPE)
        _check(load().cgpu_count_query(self._h, nq, _dp(thr), out.ctypes.data_as(C.POINTER(RowSum))))
        return out

    def count_query_rows(self, row0: int, thr: np.ndarray) -> np.ndarray:
        """The same for items row0 .. row0+len(thr)-1 only (after one whole-range query)."""
        thr = np.ascontiguousarray(thr, dtype=np.float64)
        thr = thr.reshape(thr.shape[0], -1)
        nq = thr.shape[1]
        out = np.zeros((thr.shape[0], self.S, nq), dtype=ROWSUM_DTYPE)
        _check(load().cgpu_count_query_rows(self._h, int(row0), int(thr.shape[0]), nq, _dp(thr),
                                            out.ctypes.data_as(C.POINTER(RowSum))))
        return out

    def order_stat(self, s: int, rank) -> np.ndarray:
        rank = np.ascontiguousarray(rank, dtype=np.int32)
        val = np.zeros(self.cm)
        _check(load().cgpu_order_stat(self._h, int(s), _ip(rank), _dp(val)))
        return val

    def mctest(self, i0: int, datvec, ass_ptr, ass_idx, varadd: float = 1.0):
        datvec = np.ascontiguousarray(datvec, dtype=np.float64)
        ass_ptr = np.ascontiguousarray(ass_ptr, dtype=np.int32)
        ass_idx = np.ascontiguousarray(ass_idx, dtype=np.int32)
        cm = len(ass_ptr) - 1
        outs = [np.zeros(cm) for _ in range(4)]
        _check(load().cgpu_mctest(self._h, int(i0), _dp(datvec), C.c_double(varadd), cm, _ip(ass_ptr),
                                  _ip(ass_idx), *[_dp(o) for o in outs]))
        return tuple(outs)

    def devptr(self, s: int):
        p = C.c_void_p()
        rows, cols, ld = C.c_int(), C.c_int(), C.c_int()
        _check(load().cgpu_reps_devptr(self._h, int(s), C.byref(p), C.byref(rows), C.byref(cols), C.byref(ld)))
        return p.value, rows.value, cols.value, ld.value

    def free(self):
        if self._h:
            load().cgpu_reps_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def kernel_times() -> dict:
    """{kernel: (launches, total_ms, algorithmic work)} recorded since the last call (needs set_option("ktime", 1))."""
    buf = C.create_string_buffer(1 << 16)
    _check(load().cgpu_kernel_times(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms, work = line.rsplit(" ", 3)
        out[name] = (int(cnt), float(ms), float(work))
    return out


def bench_peaks(sustained_seconds: float = 0.0) -> dict:
    """roofline denominators measured on this device: int8 tcgen05 issue rate (burst, and back to back
    for `sustained_seconds` when given), FP64 DFMA rate"""
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    _check(load().cgpu_bench_i8_peak(C.byref(a)))
    _check(load().cgpu_bench_fp64_peak(C.byref(b)))
    out = {"int8_tops": a.value, "fp64_tflops": b.value}
    if sustained_seconds > 0:
        _check(load().cgpu_bench_i8_peak_sustained(C.c_double(sustained_seconds), C.byref(c)))
        out["int8_tops_sustained"] = c.value
    return out


def smooth_count(rowsum, B: int, t: float) -> float:
    q = RowSum(*[float(rowsum[k]) for k in ("cnt", "lo", "hi", "min1", "min2", "max1", "max2")])
    return float(load().cgpu_smooth_count(C.byref(q), int(B), C.c_double(t)))


def resample_counts(n: int, bn, bb, seed: int, s: int, rng_mode: int = RNG_MT19937, stream_offset: int = 0):
    """Site-count matrix (nb x n, uint8) of scale s for this rank's replicates."""
    bn = np.ascontiguousarray(bn, dtype=np.int32)
    bb = np.ascontiguousarray(bb, dtype=np.int32)
    b0, b1 = shard_range(int(bb[s]))
    out = np.zeros((b1 - b0, n), dtype=np.uint8)
    _check(load().cgpu_resample_counts(int(n), len(bn), _ip(bn), _ip(bb), C.c_ulong(seed),
                                       C.c_ulonglong(stream_offset), rng_mode, int(s),
                                       out.ctypes.data_as(C.POINTER(C.c_ubyte))))
    return out


def mt19937_stream(seed: int, skip: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint32)
    _check(load().cgpu_mt19937_stream(C.c_ulong(seed), C.c_ulonglong(skip), C.c_ulonglong(count),
                                      out.ctypes.data_as(C.POINTER(C.c_uint))))
    return out


def makermt(L: np.ndarray, rr, bb, seed: int, rng_mode: int = RNG_MT19937, sum_mode: int = SUM_STRICT) -> dict:
    """What `makermt` computes before it writes `.rmt` (src/makermt.c:109-224):
    dict(datvec, rr (= recomputed rr1), bb, reps)."""
    datvec, bn, rr1 = scale_setup(L, rr)
    reps = scaleboot_all(L, bn, bb, seed, 0, rng_mode, sum_mode)
    return dict(datvec=datvec, rr=rr1, bb=np.asarray(bb, dtype=np.int32), reps=reps, bn=bn)
