"""One-process-per-GPU plumbing (torch.distributed) for the sharded path.

Replicates are sharded by column range over ranks (cgpu_set_shard); the makermt side
needs no data-path collective.  The consel side combines the per-rank row summaries of a
count query with ONE small collective per fit round: the integer count adds, `lo` is a
max, `hi` a min and the two smallest / two largest values merge — all exact and order
independent, so the combined summary is bit-identical to the one a single device would
compute.  Works with any backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

FIELDS = ("cnt", "lo", "hi", "min1", "min2", "max1", "max2")


def shard_range(bb: int, rank: int, world: int) -> tuple[int, int]:
    """Column range of `rank`: mirrors cgpu_shard_range (consel_b200/csrc/capi.cu)."""
    return bb * rank // world, bb * (rank + 1) // world


def _dist():
    import torch.distributed as dist
    return dist


def _device_for(group=None):
    import torch
    dist = _dist()
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def combine_rowsums(q: np.ndarray, group=None) -> np.ndarray:
    """All-reduce a structured array of row summaries (any shape, dtype api.ROWSUM_DTYPE)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    if world == 1:
        return q
    dev = _device_for(group)
    # ONE collective per fit round: every rank contributes its packed (7, N) summary block and
    # reduces the gathered copies itself (sum / max / min / two-smallest / two-largest are all
    # exact and order independent)
    flat = np.stack([np.ascontiguousarray(q[f], dtype=np.float64).ravel() for f in FIELDS])  # (7, N)
    t = torch.from_numpy(flat).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    g = torch.stack(parts).cpu().numpy()  # (world, 7, N)
    mins = np.sort(np.concatenate([g[:, 3], g[:, 4]]), axis=0)[:2]
    maxs = np.sort(np.concatenate([g[:, 5], g[:, 6]]), axis=0)[-2:][::-1]
    res = [g[:, 0].sum(axis=0), g[:, 1].max(axis=0), g[:, 2].min(axis=0), mins[0], mins[1], maxs[0], maxs[1]]
    out = np.empty_like(q)
    for f, v in zip(FIELDS, res):
        out[f] = v.reshape(q.shape)
    return out


def gather_columns(local: np.ndarray, bb: int, group=None) -> np.ndarray:
    """All-gather rank-local column blocks (rows x nb_rank) into the full (rows x bb) block,
    e.g. the replicate block of the scale nearest r=1 that the KH/SH tests need whole."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    if world == 1:
        return local
    dev = _device_for(group)
    rows = local.shape[0]
    parts = []
    for r in range(world):
        b0, b1 = shard_range(bb, r, world)
        parts.append(torch.empty((rows, b1 - b0), dtype=torch.float64, device=dev))
    mine = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    dist.all_gather(parts, mine, group=group) if len({p.shape for p in parts}) == 1 else _uneven_gather(parts, mine, group)
    return torch.cat(parts, dim=1).cpu().numpy()


class _DevBlock:
    """A rows x cols fp64 device block exposed through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, rows: int, cols: int, ld: int):
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": (ld * 8, 8)}


def gather_columns_device(ptr: int, rows: int, cols: int, ld: int, bb: int, group=None):
    """gather_columns for a block that already lives on this rank's GPU (cgpu_reps_devptr):
    NCCL all-gather straight from the library's memory, result a contiguous rows x bb device tensor."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    dev = _device_for(group)
    mine = torch.as_tensor(_DevBlock(ptr, rows, cols, ld), device=dev).contiguous()
    if world == 1:
        return mine
    parts = []
    for r in range(world):
        b0, b1 = shard_range(bb, r, world)
        parts.append(torch.empty((rows, b1 - b0), dtype=torch.float64, device=dev))
    dist.all_gather(parts, mine, group=group) if len({p.shape for p in parts}) == 1 else _uneven_gather(parts, mine, group)
    return torch.cat(parts, dim=1).contiguous()


def _uneven_gather(parts, mine, group):
    dist = _dist()
    rank = dist.get_rank(group)
    for r, p in enumerate(parts):
        if r == rank:
            p.copy_(mine)
        dist.broadcast(p, src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def max_over_ranks(x: float, group=None) -> float:
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=_device_for(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
