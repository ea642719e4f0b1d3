"""In-process makermt -> consel (no .rmt round trip, the shape the reference
ships as makerep, src/makerep.c:304-376).

The replicate blocks stay on the device; statistics, count queries and the KH/SH tests
run through the C ABI; the per-item AU fit runs in the host library libconselhost.so
(consel_b200/host/audriver.cpp, the same lockstep driver the `consel` program uses).
With a library communicator (cgpu_comm_init, one process per GPU) every rank holds a column
shard of every scale; the collectives (one all-gather of the scale nearest r=1, one grouped
reduce per AU round, one all-gather of the results) are NCCL calls made inside the native
libraries (consel_b200/csrc/comm.cu, consel_b200/host/pipeline.cpp).

Host logic here mirrors src/consel.c: read_assoc (item order, :459-541), do_batest
(:1311-1340), find_i0 (:1778-1789), getseval (:2151-2158) and the .pv column layout of
write_pv (:1611-1683): [bp, pp, kh, sh, wkh, wsh, au, np].
"""
from __future__ import annotations

import ctypes as C
import functools
import threading
import time

import numpy as np

from . import api

_host = None
QUERY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(api.RowSum))
DIST_QUERY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(api.RowSum), C.c_int, C.POINTER(C.c_int))


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        p = api.PKG / "libconselhost.so"
        if not p.exists():
            raise api.CgpuError(-2, f"{p} is missing: run __graft_entry__.build()")
        api.load()  # libconselhost.so resolves cgpu_smooth_count from libconselgpu.so
        _host = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    return _host


def find_i0(rr) -> int:
    return int(np.argmin(np.abs(np.asarray(rr, dtype=np.float64) - 1.0)))


def identity_assoc(K: int):
    return np.arange(K + 1, dtype=np.int32), np.arange(K, dtype=np.int32)


def item_order(datvec, ass_ptr, ass_idx, nosort: bool = False):
    """read_assoc: observed statistic vectors (max over the complement minus each member,
    ascending) and the item order by compdvec (length first, then lexicographic)."""
    datvec = np.asarray(datvec, dtype=np.float64)
    cm = len(ass_ptr) - 1
    ass_ptr = np.asarray(ass_ptr)
    ass_idx = np.asarray(ass_idx)
    if cm and np.all(np.diff(ass_ptr) == 1):
        # singleton associations (the default identity): max over "all but i" from the top two
        a = int(np.argmax(datvec))
        second = np.max(np.delete(datvec, a)) if len(datvec) > 1 else -1.0e30
        trees = ass_idx[ass_ptr[:-1]]
        obs_all = np.where(trees == a, second, datvec[a]) - datvec[trees]
        order = np.arange(cm) if nosort else np.argsort(obs_all, kind="stable")
        order = order.astype(np.int32)
        return order, obs_all[order], np.arange(cm + 1, dtype=np.int32), trees[order].astype(np.int32)
    keys = []
    for e in range(cm):
        a = np.asarray(ass_idx[ass_ptr[e]:ass_ptr[e + 1]])
        mask = np.ones(len(datvec), bool)
        mask[a] = False
        mx = datvec[mask].max() if mask.any() else -1.0e30
        mx = max(mx, -1.0e30)
        keys.append(np.sort(mx - datvec[a]))

    def cmp(i, j):
        x, y = keys[i], keys[j]
        out = -1 if len(x) < len(y) else (1 if len(x) > len(y) else 0)
        for a, b in zip(x, y):
            if a < b:
                return -1
            if a > b:
                return 1
        return out

    order = list(range(cm)) if nosort else sorted(range(cm), key=functools.cmp_to_key(cmp))
    obs = np.array([keys[i][0] for i in order])
    ptr, idx = [0], []
    for i in order:
        idx += [int(v) for v in ass_idx[ass_ptr[i]:ass_ptr[i + 1]]]
        ptr.append(len(idx))
    return np.array(order, dtype=np.int32), obs, np.array(ptr, dtype=np.int32), np.array(idx, dtype=np.int32)


def comm_init_from_torch() -> tuple[int, int]:
    """Create the library's own NCCL communicator from an initialised torch.distributed job:
    rank 0 makes the unique id (cgpu_comm_unique_id), the launcher's process group only
    broadcasts those 128 bytes; every collective of the data path then runs inside
    libconselgpu.so.  Returns (rank, world).  No-op for single-process runs."""
    try:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0, 1
    except Exception:
        return 0, 1
    lib = api.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    cur_r, cur_w = C.c_int(), C.c_int()
    lib.cgpu_comm_info(C.byref(cur_r), C.byref(cur_w))
    if (cur_r.value, cur_w.value) == (rank, world):
        return rank, world
    buf = C.create_string_buffer(128)
    if rank == 0:
        api._check(lib.cgpu_comm_unique_id(buf))
    box = [buf.raw]
    dist.broadcast_object_list(box, src=0)
    api._check(lib.cgpu_comm_init(rank, world, box[0]))
    return rank, world


def comm_info() -> tuple[int, int]:
    r, w = C.c_int(), C.c_int()
    api.load().cgpu_comm_info(C.byref(r), C.byref(w))
    return r.value, w.value


def consel(reps: api.Reps, datvec, rr, bb, ass=None, alpha=(0.05, 0.10, 0.5, 0.90, 0.95), threshold: float = 0.0,
           kappa: float = 1.0, fitmode: int = 1, do_mc: bool = True, do_au: bool = True, confidence: bool = True,
           nosort: bool = False, varadd: float = 1.0) -> dict:
    """Everything `consel foo` computes from foo.rmt, from device-resident replicate blocks:
    a thin wrapper over chost_consel_run (consel_b200/host/pipeline.cpp).  With a library
    communicator (cgpu_comm_init / comm_init_from_torch) the call is collective and every rank
    returns the full table."""
    datvec = np.ascontiguousarray(datvec, dtype=np.float64)
    rr = np.ascontiguousarray(rr, dtype=np.float64)
    bb = np.ascontiguousarray(bb, dtype=np.int32)
    K, kk = len(datvec), len(rr)
    ptr0, idx0 = identity_assoc(K) if ass is None else ass
    order, obs, ptr, idx = item_order(datvec, ptr0, idx0, nosort)
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    cm = len(order)
    i0 = find_i0(rr)

    w = np.exp(datvec - datvec.max())  # do_batest
    w = w / np.add.accumulate(w)[-1]
    if np.all(np.diff(ptr) == 1):
        pp = w[idx[ptr[:-1]]]
    else:
        pp = np.array([np.add.accumulate(w[idx[ptr[i]:ptr[i + 1]]])[-1] for i in range(cm)])

    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    pv, se = np.zeros((cm, 8)), np.zeros((cm, 8))
    cnt, ax7 = np.zeros((cm, kk)), np.zeros((cm, 7))
    ci6 = np.zeros((6, cm, len(alpha)))
    ret = np.zeros(cm, dtype=np.int32)
    tm = np.zeros(8)
    lib = host_lib()
    lib.chost_consel_run.restype = C.c_int
    rc = lib.chost_consel_run(reps._h, K, api._dp(datvec), kk, api._dp(rr), api._ip(bb), cm, api._ip(ptr), api._ip(idx),
                              C.c_double(threshold), C.c_double(varadd), int(fitmode), C.c_double(kappa), int(do_mc),
                              int(do_au), len(alpha), api._dp(alpha), api._dp(pv), api._dp(se), api._dp(cnt), api._dp(ax7),
                              api._dp(ci6) if confidence else None, api._ip(ret), api._dp(tm))
    if rc:
        raise api.CgpuError(rc, api.load().cgpu_last_error().decode())
    reps.cm = cm
    pv[:, 1] = pp
    timing = {"mctest_s": float(tm[0]), "stats_s": float(tm[1])}
    out = dict(item=order, stat=obs, pv=pv, se=se, cnt=cnt, i0=i0, timing=timing)
    if do_au and kk >= 2:
        out.update(ax=ax7, ci=ci6 if confidence else None, alpha=alpha, ret=ret)
        timing["au_s"] = float(tm[2])
    return out


def _rowsums_local(L: np.ndarray, comm):
    """this rank's share of the row sums (all of them without a communicator): host threads only, no device or
    collective call, so it may run beside a resampling call.  Returns (values, rows per rank, world)."""
    rank, world = comm
    K, n = L.shape
    lib = api.load()
    if world == 1 or K < 4 * world:
        datvec = np.zeros(K)
        lib.cgpu_host_rowsums(api._dp(L), K, n, api._dp(datvec))
        return datvec, K, 1
    per = -(-K // world)
    r0, r1 = min(K, rank * per), min(K, rank * per + per)
    mine = np.zeros(per)
    if r1 > r0:
        lib.cgpu_host_rowsums(api._dp(L[r0:r1]), r1 - r0, n, api._dp(mine))  # a row range of a C-contiguous matrix is contiguous
    return mine, per, world


def _rowsums_gather(part, K: int, _allgather=None) -> np.ndarray:
    """completes datvec on every rank: one small all-gather of the per-rank shares (collective)"""
    mine, per, world = part
    if world == 1:
        return mine
    everyone = np.zeros(per * world)
    if _allgather is not None:
        _allgather(mine, everyone)
    else:
        api._check(api.load().cgpu_comm_allgather_host(mine.ctypes.data_as(C.c_void_p), everyone.ctypes.data_as(C.c_void_p),
                                                       C.c_longlong(per * 8)))
    return everyone[:K].copy()


def scale_setup(L: np.ndarray, rr, _comm=None, _allgather=None):
    """api.scale_setup with the row sums of a sharded run split over the ranks: one process per GPU shares the
    host's cores and memory system, so every rank sums ceil(K/world) rows (sequentially, the reference's
    rounding order: src/makermt.c:159-164) and one small all-gather completes datvec on every rank --
    instead of `world` processes each streaming the whole matrix.  Collective when a communicator exists.
    (_comm / _allgather: test seams -- (rank, world) and a stand-in for cgpu_comm_allgather_host.)"""
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    datvec = _rowsums_gather(_rowsums_local(L, _comm if _comm is not None else comm_info()), K, _allgather)
    rr = np.asarray(rr, dtype=np.float64)
    bn = (rr * np.float64(n)).astype(np.int32)  # bn = (int)(rr*n), rr1 = bn/n: src/makermt.c:165-168, as api.scale_setup
    return datvec, bn, bn.astype(np.float64) / np.float64(n)


def run(L, rr, bb, seed: int, rng_mode: int = api.RNG_MT19937, sum_mode: int = api.SUM_STRICT, matrix=None, **kw) -> dict:
    """makermt + consel in one process: L (K x n) -> p-value table, nothing touches a file.
    The row sums (datvec; host threads) are formed while the device resamples: consel needs them first."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    K, n = L.shape
    rr_a = np.asarray(rr, dtype=np.float64)
    bn = (rr_a * np.float64(n)).astype(np.int32)
    rr1 = bn.astype(np.float64) / np.float64(n)
    comm = comm_info()
    box = {}

    def sums():
        try:
            box["part"] = _rowsums_local(L, comm)
        except BaseException as e:  # re-raised on the calling thread
            box["error"] = e
    th = threading.Thread(target=sums)
    th.start()
    try:
        t0 = time.perf_counter()
        mat = matrix if matrix is not None else api.Matrix(L)
        reps = mat.scaleboot(bn, bb, seed, 0, rng_mode, sum_mode)
        t_boot = time.perf_counter() - t0
    finally:
        th.join()
    if "error" in box:
        raise box["error"]
    datvec = _rowsums_gather(box["part"], K)
    res = consel(reps, datvec, rr1, bb, **kw)
    res["timing"]["makermt_s"] = t_boot
    res["timing"]["device_ms"] = api.last_timing()
    res.update(datvec=datvec, rr=rr1, bb=np.asarray(bb, dtype=np.int32))
    reps.free()
    if matrix is None:
        mat.free()
    return res
