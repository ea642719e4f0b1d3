"""N>1 host logic on CPU: world_size-2 (and 3) gloo groups exercise the replicate
sharding arithmetic and the exact combine of per-rank row summaries (consel_b200/dist.py)
against the single-rank answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from consel_b200 import dist as cdist
from consel_b200.api import ROWSUM_DTYPE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _summary(rows: np.ndarray, thr: np.ndarray) -> np.ndarray:
    """numpy statement of cgpu_rowsum for rows (cm, B) and thresholds (cm, nq)"""
    cm, nq = thr.shape
    q = np.zeros((cm, nq), dtype=ROWSUM_DTYPE)
    for i in range(cm):
        v = np.sort(rows[i])
        pad_lo = np.concatenate([v, [1e30, 1e30]])
        pad_hi = np.concatenate([[-1e30, -1e30], v])
        for k in range(nq):
            t = thr[i, k]
            le, gt = v[v <= t], v[v > t]
            q[i, k] = (len(le), le.max() if len(le) else -1e30, gt.min() if len(gt) else 1e30,
                       pad_lo[0], pad_lo[1], pad_hi[-1], pad_hi[-2])
    return q


def _worker(rank, world, port, seed, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    cm, B, nq = 7, 53, 3
    rows = rng.standard_normal((cm, B))
    rows[2, :] = 0.25          # a constant row: ties everywhere
    rows[3, 10] = rows[3, 40]  # a tie across shards
    thr = np.stack([np.zeros(cm), rows[:, 5], np.full(cm, 9.0)], axis=1)
    b0, b1 = cdist.shard_range(B, rank, world)
    local = _summary(rows[:, b0:b1], thr)
    combined = cdist.combine_rowsums(local)
    whole = _summary(rows, thr)
    ok = all(np.array_equal(combined[f], whole[f]) for f in cdist.FIELDS)
    full = cdist.gather_columns(rows[:, b0:b1], B)
    ok = ok and np.array_equal(full, rows)
    ok = ok and cdist.max_over_ranks(float(rank)) == float(world - 1)
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rowsum_combine_and_gather_match_single_rank(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 123, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_shard_ranges_partition_every_scale():
    import consel_b200 as cg
    for bb in (0, 1, 7, 100, 10007):
        for world in (1, 2, 3, 8):
            cols = []
            for rank in range(world):
                b0, b1 = cdist.shard_range(bb, rank, world)
                assert 0 <= b0 <= b1 <= bb
                cols += list(range(b0, b1))
            assert cols == list(range(bb))
    # the C library uses the same arithmetic (no GPU needed for these two calls)
    lib = cg.load()
    assert lib.cgpu_set_shard(2, 3) == 0
    assert cg.shard_range(100) == cdist.shard_range(100, 2, 3)
    assert lib.cgpu_set_shard(0, 1) == 0
    assert lib.cgpu_set_shard(3, 3) != 0
