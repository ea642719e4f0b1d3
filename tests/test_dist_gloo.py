"""N>1 host logic on CPU: world_size-2 (and 3) gloo groups drive the SAME native lockstep the
multi-GPU run uses (audriver's collective rounds through chost_au_run_dist, consel_b200/host/
pipeline.cpp) with a numpy + gloo statement of cgpu_count_query_dist as the query callback, and
compare every AU / NP p-value, fit diagnostic and confidence limit bit for bit with the
single-process driver (chost_au_run) on the reference-written statistics of tests/golden/g100.rep.
"""
import ctypes as C
import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from consel_b200 import dist as cdist
from consel_b200 import pipeline
from consel_b200.api import ROWSUM_DTYPE, RowSum

GOLD = Path(__file__).resolve().parent / "golden"
ALPHA = np.array([0.05, 0.10, 0.5, 0.90, 0.95])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _summary(rows: np.ndarray, thr: np.ndarray) -> np.ndarray:
    """numpy statement of cgpu_rowsum (count, lo, hi only) for rows (cm, B) and thresholds (cm, nq)"""
    cm, nq = thr.shape
    q = np.zeros((cm, nq), dtype=ROWSUM_DTYPE)
    for i in range(cm):
        v = rows[i]
        for k in range(nq):
            t = thr[i, k]
            le, gt = v[v <= t], v[v > t]
            q[i, k]["cnt"] = len(le)
            q[i, k]["lo"] = le.max() if len(le) else -1e30
            q[i, k]["hi"] = gt.min() if len(gt) else 1e30
    return q


def _edges(rows: np.ndarray) -> np.ndarray:
    v = np.sort(rows, axis=1)
    return np.stack([v[:, 0], v[:, 1], v[:, -1], v[:, -2]], axis=1)  # min1 min2 max1 max2


def _load_stats():
    from oracle import oracle as orc
    rep = orc.read_rep(GOLD / "g100.rep")
    stats = [np.ascontiguousarray(s) for s in rep["stats"]]
    return stats, np.asarray(rep["rr"], dtype=np.float64), np.asarray(rep["bb"], dtype=np.int32)


def _fill(q, edges_s):
    for f, c in zip(("min1", "min2", "max1", "max2"), range(4)):
        q[f] = edges_s[:, c][:, None]
    return q


def _au_inputs(stats, rr, bb):
    i0 = pipeline.find_i0(rr)
    cm, kk = stats[0].shape[0], len(stats)
    med = np.sort(stats[i0], axis=1)[:, bb[i0] // 2].copy()
    edges = [_edges(s) for s in stats]
    r0 = np.zeros((cm, kk, 1), dtype=ROWSUM_DTYPE)
    for s in range(kk):
        r0[:, s, :] = _fill(_summary(stats[s], np.zeros((cm, 1))), edges[s])
    return i0, med, edges, np.ascontiguousarray(r0[:, :, 0])


def _single(stats, rr, bb):
    """all items in one process through chost_au_run with a plain numpy query"""
    lib = pipeline.host_lib()
    i0, med, edges, r0 = _au_inputs(stats, rr, bb)
    cm, kk = r0.shape

    def query(ctx, nq, thr_p, out_p):
        thr = np.ctypeslib.as_array(thr_p, shape=(cm, nq))
        out = np.zeros((cm, kk, nq), dtype=ROWSUM_DTYPE)
        for s in range(kk):
            out[:, s, :] = _fill(_summary(stats[s], thr), edges[s])
        C.memmove(out_p, out.ctypes.data, out.nbytes)
        return 0
    cb = pipeline.QUERY_FN(query)
    pv2, se2, ax7 = np.zeros((cm, 2)), np.zeros((cm, 2)), np.zeros((cm, 7))
    ci6, ret = np.zeros((6, cm, len(ALPHA))), np.zeros(cm, dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.chost_au_run(cm, kk, dp(rr), bb.ctypes.data_as(C.POINTER(C.c_int)), i0, C.c_double(0.0), dp(med),
                          r0.ctypes.data_as(C.POINTER(RowSum)), len(ALPHA), dp(ALPHA), 1, C.c_double(1.0), cb, None,
                          dp(pv2), dp(se2), dp(ax7), dp(ci6), ret.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return pv2, se2, ax7, ci6, ret


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CHOST_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stats, rr, bb = _load_stats()
        i0, med, edges, r0 = _au_inputs(stats, rr, bb)
        cm, kk = r0.shape
        per = (cm + world - 1) // world
        row0, nrows = cdist.item_range(cm, rank, world)
        shards = []
        for s in range(kk):
            b0, b1 = cdist.shard_range(int(bb[s]), rank, world)
            shards.append(stats[s][:, b0:b1])
        rounds = [0]

        def dquery(ctx, nq, thr_p, out_p, active, any_p):
            # the gloo statement of cgpu_count_query_dist(all_rows=0): thresholds of every rank's items
            # all-gathered (padded to `per` rows), every item counted over this rank's columns,
            # count summed / lo maxed / hi minned, the rank keeps the rows of its own items
            rounds[0] += 1
            mine = np.full((per, nq), 1.0e300)
            if nrows:
                mine[:nrows] = np.ctypeslib.as_array(thr_p, shape=(nrows, nq))
            parts = [torch.empty((per, nq), dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(mine))
            thr_all = torch.cat(parts).numpy()[:cm]
            flags = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(flags, torch.tensor([1.0 if active else 0.0], dtype=torch.float64))
            any_p[0] = int(any(f.item() != 0.0 for f in flags))
            loc = np.zeros((cm, kk, nq), dtype=ROWSUM_DTYPE)
            for s in range(kk):
                loc[:, s, :] = _summary(shards[s], thr_all)
            cnt = torch.from_numpy(np.ascontiguousarray(loc["cnt"]))
            lo = torch.from_numpy(np.ascontiguousarray(loc["lo"]))
            hi = torch.from_numpy(np.ascontiguousarray(loc["hi"]))
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            dist.all_reduce(lo, op=dist.ReduceOp.MAX)
            dist.all_reduce(hi, op=dist.ReduceOp.MIN)
            out = np.zeros((nrows, kk, nq), dtype=ROWSUM_DTYPE)
            out["cnt"], out["lo"], out["hi"] = cnt.numpy()[row0:row0 + nrows], lo.numpy()[row0:row0 + nrows], hi.numpy()[row0:row0 + nrows]
            for s in range(kk):
                for f, c in zip(("min1", "min2", "max1", "max2"), range(4)):
                    out[f][:, s, :] = edges[s][row0:row0 + nrows, c][:, None]
            if nrows:
                C.memmove(out_p, out.ctypes.data, out.nbytes)
            return 0

        cb = pipeline.DIST_QUERY_FN(dquery)
        lib = pipeline.host_lib()
        n1 = max(nrows, 1)
        pv2, se2, ax7 = np.zeros((n1, 2)), np.zeros((n1, 2)), np.zeros((n1, 7))
        ci6, ret = np.zeros((6, n1, len(ALPHA))), np.zeros(n1, dtype=np.int32)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        med_l = np.ascontiguousarray(med[row0:row0 + nrows]) if nrows else np.zeros(1)
        r0_l = np.ascontiguousarray(r0[row0:row0 + nrows]) if nrows else np.zeros((1, kk), dtype=ROWSUM_DTYPE)
        rc = lib.chost_au_run_dist(nrows, kk, dp(rr), bb.ctypes.data_as(C.POINTER(C.c_int)), i0, C.c_double(0.0), dp(med_l),
                                   r0_l.ctypes.data_as(C.POINTER(RowSum)), len(ALPHA), dp(ALPHA), 1, C.c_double(1.0), cb, None,
                                   dp(pv2), dp(se2), dp(ax7), dp(ci6), ret.ctypes.data_as(C.POINTER(C.c_int)))
        # every rank must have entered the same number of collective rounds
        nr = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(nr, torch.tensor([float(rounds[0])], dtype=torch.float64))
        same_rounds = len({int(t.item()) for t in nr}) == 1
        out_q.put((rank, rc, row0, nrows, pv2[:nrows].copy(), se2[:nrows].copy(), ax7[:nrows].copy(),
                   ci6[:, :nrows].copy() if nrows else ci6[:, :0].copy(), ret[:nrows].copy(), same_rounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_collective_au_lockstep_equals_single_process(world):
    stats, rr, bb = _load_stats()
    want = _single(stats, rr, bb)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = 0
    for rank, rc, row0, nrows, pv2, se2, ax7, ci6, ret, same_rounds in res:
        assert rc == 0 and same_rounds
        assert (row0, nrows) == cdist.item_range(stats[0].shape[0], rank, world)
        sl = slice(row0, row0 + nrows)
        assert np.array_equal(pv2, want[0][sl]) and np.array_equal(se2, want[1][sl])
        assert np.array_equal(ax7, want[2][sl], equal_nan=True)
        assert np.array_equal(ci6, want[3][:, sl], equal_nan=True)
        assert np.array_equal(ret, want[4][sl])
        covered += nrows
    assert covered == stats[0].shape[0]


def test_shard_and_item_ranges_partition():
    import consel_b200 as cg
    for bb in (0, 1, 7, 100, 10007):
        for world in (1, 2, 3, 8):
            cols, items = [], []
            for rank in range(world):
                b0, b1 = cdist.shard_range(bb, rank, world)
                cols += list(range(b0, b1))
                r0, n = cdist.item_range(bb, rank, world)
                items += list(range(r0, r0 + n))
            assert cols == list(range(bb)) and items == list(range(bb))
    lib = cg.load()
    b0, b1 = C.c_int(), C.c_int()
    assert lib.cgpu_shard_range(100, C.byref(b0), C.byref(b1)) == 0 and (b0.value, b1.value) == (0, 100)
    assert lib.cgpu_comm_item_range(10, C.byref(b0), C.byref(b1)) == 0 and (b0.value, b1.value) == (0, 10)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_row_sums_equal_the_sequential_sums(world):
    """pipeline.scale_setup: every rank sums its own row range, one all-gather completes datvec; the result is
    api.scale_setup's (the reference's left-to-right sums) bit for bit, for K not divisible by the world size"""
    import consel_b200 as cg
    rng = np.random.default_rng(world)
    L = -rng.exponential(4.0, size=(37, 333)) - 1.0
    rr = [0.5, 1.0, 1.4]
    want = cg.scale_setup(L, rr)
    per = -(-37 // world)
    shares = {}
    for rank in range(world):  # pass 1: what every rank would send
        pipeline.scale_setup(L, rr, _comm=(rank, world), _allgather=lambda mine, out, r=rank: shares.__setitem__(r, mine.copy()))
    assert all(len(v) == per for v in shares.values())
    gathered = np.concatenate([shares[r] for r in range(world)])
    for rank in range(world):  # pass 2: every rank receives the same gathered vector
        got = pipeline.scale_setup(L, rr, _comm=(rank, world), _allgather=lambda mine, out: out.__setitem__(slice(None), gathered))
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
