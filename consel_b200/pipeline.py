"""In-process makermt -> consel (no .rmt round trip, the shape the reference
ships as makerep, src/makerep.c:304-376).

The replicate blocks stay on the device; statistics, count queries and the KH/SH tests
run through the C ABI; the per-item AU fit runs in the host library libconselhost.so
(consel_b200/host/audriver.cpp, the same lockstep driver the `consel` program uses).
With torch.distributed initialised and cgpu_set_shard(rank, world) in effect, every rank
holds a column shard of every scale and the only collectives are the small combines of
consel_b200.dist.

Host logic here mirrors src/consel.c: read_assoc (item order, :459-541), do_batest
(:1311-1340), find_i0 (:1778-1789), getseval (:2151-2158) and the .pv column layout of
write_pv (:1611-1683): [bp, pp, kh, sh, wkh, wsh, au, np].
"""
from __future__ import annotations

import ctypes as C
import functools
import time

import numpy as np

from . import api, dist as cdist

_host = None
QUERY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(api.RowSum))


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


def _distributed() -> bool:
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def _dist_order_stat(reps: api.Reps, s: int, rank_k: np.ndarray) -> np.ndarray:
    """k-th smallest of sharded rows: bisection on the 64 bits of the order-preserving
    integer image of a double, one combined count query per bit (exact)."""
    cm = reps.cm
    res = np.zeros(cm, dtype=np.uint64)

    def unkey(k):
        neg = (k >> np.uint64(63)) == 0
        u = np.where(neg, ~k, k & np.uint64(0x7FFFFFFFFFFFFFFF))
        return u.view(np.float64)

    for bit in range(63, -1, -1):
        cand = res | (np.uint64(1) << np.uint64(bit))
        # count(key < cand) == count(v < unkey(cand)) == count(v <= prev_double(unkey(cand)))
        t = np.nextafter(unkey(cand), -np.inf)
        t = np.where(np.isnan(t), np.inf, t)
        q = cdist.combine_rowsums(reps.count_query(t.reshape(cm, 1)))
        cnt = q["cnt"][:, s, 0]
        res = np.where(cnt <= rank_k, cand, res)
    return unkey(res)


def consel(reps: api.Reps, datvec, rr, bb, ass=None, alpha=(0.05, 0.10, 0.5, 0.90, 0.95), threshold: float = 0.0,
           kappa: float = 1.0, fitmode: int = 1, do_mc: bool = True, do_au: bool = True, confidence: bool = True,
           nosort: bool = False, varadd: float = 1.0) -> dict:
    """Everything `consel foo` computes from foo.rmt, from device-resident replicate blocks."""
    datvec = np.asarray(datvec, dtype=np.float64)
    rr = np.ascontiguousarray(rr, dtype=np.float64)
    bb = np.ascontiguousarray(bb, dtype=np.int32)
    K, kk = len(datvec), len(rr)
    ptr0, idx0 = identity_assoc(K) if ass is None else ass
    order, obs, ptr, idx = item_order(datvec, ptr0, idx0, nosort)
    cm = len(order)
    i0 = find_i0(rr)
    nb0 = int(bb[i0])
    dist_mode = _distributed()
    timing = {}

    w = np.exp(datvec - datvec.max())  # do_batest
    w = w / np.add.accumulate(w)[-1]
    if np.all(np.diff(ptr) == 1):
        pp = w[idx[ptr[:-1]]]
    else:
        pp = np.array([np.add.accumulate(w[idx[ptr[i]:ptr[i + 1]]])[-1] for i in range(cm)])

    pv = np.zeros((cm, 8))
    se = np.zeros((cm, 8))
    pv[:, 1] = pp
    t0 = time.perf_counter()
    one = None  # the whole scale-i0 block on this rank (distributed runs): KH/SH tests and the median need it
    whole_d = None
    if dist_mode and (do_mc or (do_au and kk >= 2)):
        import torch
        import torch.distributed as dist
        if dist.get_backend() == "nccl":  # all-gather straight from the library's device block
            p, rows, cols, ld = reps.devptr(i0)
            whole_d = cdist.gather_columns_device(p, rows, cols, ld, nb0)
            torch.cuda.synchronize()
            api.set_shard(0, 1)
            one = api.Reps.upload_ptrs(K, [nb0], [whole_d.data_ptr()])
        else:  # CPU process groups (tests): through the host
            blocks = reps.download()
            whole = cdist.gather_columns(blocks[i0], nb0)
            api.set_shard(0, 1)
            one = api.Reps.upload([whole])
        api.set_shard(dist.get_rank(), dist.get_world_size())
    if do_mc:
        if dist_mode:  # sequential row sums over all replicates of scale i0
            api.set_shard(0, 1)
            kh, sh, wkh, wsh = one.mctest(0, datvec, ptr, idx, varadd)  # returns after the library stream drained
            api.set_shard(dist.get_rank(), dist.get_world_size())
        else:
            kh, sh, wkh, wsh = reps.mctest(i0, datvec, ptr, idx, varadd)
        pv[:, 2], pv[:, 3], pv[:, 4], pv[:, 5] = kh, sh, wkh, wsh
    timing["mctest_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    reps.stats_build(ptr, idx)
    q0 = reps.count_query(np.full((cm, 1), threshold))
    if dist_mode:
        q0 = cdist.combine_rowsums(q0)
    timing["stats_s"] = time.perf_counter() - t0
    sm0 = np.zeros(cm)
    api.load().cgpu_smooth_count_many.restype = None
    r0_i0 = np.ascontiguousarray(q0[:, i0, 0])
    api.load().cgpu_smooth_count_many(r0_i0.ctypes.data_as(C.POINTER(api.RowSum)), cm, nb0, C.c_double(threshold),
                                      api._dp(sm0))
    bp = 1.0 - (1.0 - sm0 / nb0)
    pv[:, 0] = bp
    for c in (0, 2, 3, 4, 5):
        se[:, c] = np.sqrt(pv[:, c] * (1.0 - pv[:, c]) / nb0)
    out = dict(item=order, stat=obs, pv=pv, se=se, cnt=q0["cnt"][:, :, 0].copy(), i0=i0, timing=timing)

    if do_au and kk >= 2:
        t0 = time.perf_counter()
        rank_k = np.full(cm, nb0 // 2, dtype=np.int32)
        if dist_mode:  # the median of the whole scale-i0 rows, from the gathered block
            api.set_shard(0, 1)
            one.stats_build(ptr, idx)
            med = one.order_stat(0, rank_k)
            api.set_shard(dist.get_rank(), dist.get_world_size())
        else:
            med = reps.order_stat(i0, rank_k)
        r0 = np.ascontiguousarray(q0[:, :, 0])
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        pv2, se2, ax7 = np.zeros((cm, 2)), np.zeros((cm, 2)), np.zeros((cm, 7))
        ci6 = np.zeros((6, cm, len(alpha)))
        ret = np.zeros(cm, dtype=np.int32)
        lib = host_lib()
        if dist_mode:
            # Every rank fits only ITS items (the per-item fits are independent host work; with
            # one process per GPU the host cores are shared) but counts over its replicate shard
            # for ALL items.  One protocol round = (1) agree on the threshold count of the round
            # and on whether anybody is still iterating, (2) all-gather the thresholds of all
            # items, (3) local count query, (4) all-gather the (count, lo, hi) triples and reduce
            # them (sum / max / min: exact), (5) every rank takes the rows of its items.  A rank
            # whose items have converged keeps serving rounds until all ranks are done.
            import torch
            import torch.distributed as dist
            world, rank = dist.get_world_size(), dist.get_rank()
            dev = cdist._device_for(None)
            lo_r = [cm * r // world for r in range(world + 1)]
            my0, my1 = lo_r[rank], lo_r[rank + 1]
            cml = my1 - my0
            edges = {f: np.ascontiguousarray(q0[f][:, :, 0]) for f in ("min1", "min2", "max1", "max2")}  # combined once
            HUGE = 1.0e300

            def protocol_round(nq_loc: int, thr_loc):
                """thr_loc: (cml, nq_loc) or None.  Returns (active_anywhere, rows for my items or None)."""
                hdr = torch.tensor([float(nq_loc), 1.0 if nq_loc > 0 else 0.0], dtype=torch.float64, device=dev)
                dist.all_reduce(hdr, op=dist.ReduceOp.MAX)
                nq_r = int(hdr[0].item())
                if hdr[1].item() == 0.0 or nq_r == 0:
                    return False, None
                mine = np.full((cml, nq_r), HUGE)
                if nq_loc > 0:
                    mine[:, :nq_loc] = thr_loc
                parts = [torch.empty((lo_r[r + 1] - lo_r[r], nq_r), dtype=torch.float64, device=dev) for r in range(world)]
                me = torch.from_numpy(mine).to(dev)
                if len({p.shape for p in parts}) == 1:
                    dist.all_gather(parts, me)
                else:
                    cdist._uneven_gather(parts, me, None)
                thr_all = torch.cat(parts, dim=0).cpu().numpy()
                q = reps.count_query(thr_all)  # (cm, S, nq_r) over this rank's replicate columns
                # all-to-all: rank p receives, from every rank, the triples of ITS items only
                trip = torch.from_numpy(np.stack([q["cnt"], q["lo"], q["hi"]], axis=1)).to(dev)  # (cm, 3, S, nq_r)
                width = 3 * kk * nq_r
                recv = torch.empty((world * cml, width), dtype=torch.float64, device=dev)
                dist.all_to_all_single(recv, trip.reshape(cm, width), output_split_sizes=[cml] * world,
                                       input_split_sizes=[lo_r[r + 1] - lo_r[r] for r in range(world)])
                g = recv.reshape(world, cml, 3, kk, nq_r)
                cnt = g[:, :, 0].sum(dim=0)
                lo_ = g[:, :, 1].max(dim=0).values
                hi_ = g[:, :, 2].min(dim=0).values
                if nq_loc == 0:
                    return True, None
                rows = np.zeros((cml, kk, nq_r), dtype=api.ROWSUM_DTYPE)
                rows["cnt"], rows["lo"], rows["hi"] = cnt.cpu().numpy(), lo_.cpu().numpy(), hi_.cpu().numpy()
                for f in ("min1", "min2", "max1", "max2"):
                    rows[f] = edges[f][my0:my1, :, None]
                return True, rows[:, :, :nq_loc]

            def py_query(ctx, nq, thr_p, out_p):
                thr = np.ctypeslib.as_array(thr_p, shape=(cml, nq)).copy()
                _, rows = protocol_round(nq, thr)
                rows = np.ascontiguousarray(rows)
                C.memmove(out_p, rows.ctypes.data, rows.nbytes)
                return 0
            cb = QUERY_FN(py_query)
            rc = 0
            loc = dict(pv2=np.zeros((cml, 2)), se2=np.zeros((cml, 2)), ax7=np.zeros((cml, 7)),
                       ci6=np.zeros((6, cml, len(alpha))), ret=np.zeros(cml, dtype=np.int32))
            if cml > 0:
                med_l = np.ascontiguousarray(med[my0:my1])
                r0_l = np.ascontiguousarray(r0[my0:my1])
                rc = lib.chost_au_run(cml, kk, api._dp(rr), api._ip(bb), i0, C.c_double(threshold), api._dp(med_l),
                                      r0_l.ctypes.data_as(C.POINTER(api.RowSum)), len(alpha), api._dp(alpha), fitmode,
                                      C.c_double(kappa), cb, None, api._dp(loc["pv2"]), api._dp(loc["se2"]), api._dp(loc["ax7"]),
                                      api._dp(loc["ci6"]) if confidence else None, api._ip(loc["ret"]))
            while protocol_round(0, None)[0]:  # serve the ranks that are still iterating
                pass
            if rc:
                raise api.CgpuError(rc, api.load().cgpu_last_error().decode())
            gathered = [None] * world
            dist.all_gather_object(gathered, loc)
            pv2 = np.concatenate([g_["pv2"] for g_ in gathered])
            se2 = np.concatenate([g_["se2"] for g_ in gathered])
            ax7 = np.concatenate([g_["ax7"] for g_ in gathered])
            ci6 = np.concatenate([g_["ci6"] for g_ in gathered], axis=1)
            ret = np.concatenate([g_["ret"] for g_ in gathered])
        else:
            cb = C.cast(api.load().cgpu_count_query, QUERY_FN)
            ctx = reps._h
            # row-range queries let the driver fit one half of the items while the device counts for the other
            lib.chost_set_query_rows.argtypes = [C.c_void_p]
            lib.chost_set_query_rows.restype = None
            lib.chost_set_query_rows(C.cast(api.load().cgpu_count_query_rows, C.c_void_p))
            rc = lib.chost_au_run(cm, kk, api._dp(rr), api._ip(bb), i0, C.c_double(threshold), api._dp(med),
                                  r0.ctypes.data_as(C.POINTER(api.RowSum)), len(alpha), api._dp(alpha), fitmode,
                                  C.c_double(kappa), cb, ctx, api._dp(pv2), api._dp(se2), api._dp(ax7),
                                  api._dp(ci6) if confidence else None, api._ip(ret))
            if rc:
                raise api.CgpuError(rc, api.load().cgpu_last_error().decode())
        pv[:, 6:8], se[:, 6:8] = pv2, se2
        out.update(ax=ax7, ci=ci6 if confidence else None, alpha=alpha, ret=ret)
        timing["au_s"] = time.perf_counter() - t0
    if one is not None:
        one.free()
    return out


def run(L, rr, bb, seed: int, rng_mode: int = api.RNG_MT19937, sum_mode: int = api.SUM_STRICT, matrix=None, **kw) -> dict:
    """makermt + consel in one process: L (K x n) -> p-value table, nothing touches a file."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    datvec, bn, rr1 = api.scale_setup(L, rr)
    t0 = time.perf_counter()
    mat = matrix if matrix is not None else api.Matrix(L)
    reps = mat.scaleboot(bn, bb, seed, 0, rng_mode, sum_mode)
    t_boot = time.perf_counter() - t0
    res = consel(reps, datvec, rr1, bb, **kw)
    res["timing"]["makermt_s"] = t_boot
    res["timing"]["device_ms"] = api.last_timing()
    res.update(datvec=datvec, rr=rr1, bb=np.asarray(bb, dtype=np.int32))
    reps.free()
    if matrix is None:
        mat.free()
    return res
