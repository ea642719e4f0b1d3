"""Partition arithmetic of the one-process-per-GPU path (host-side mirrors, no collectives here).

Replicates are sharded by column range over the ranks (cgpu_set_shard / cgpu_comm_init); items
(rows of the statistic blocks) are owned in equal contiguous ranges for the KH/SH loops and the
AU fits (cgpu_comm_item_range).  The collectives themselves are NCCL calls inside the native
libraries: consel_b200/csrc/comm.cu (communicator, cgpu_count_query_dist, cgpu_reps_gather_scale)
and consel_b200/host/pipeline.cpp (chost_consel_run).  tests/test_dist_gloo.py drives the same
host protocol (audriver's collective lockstep) over gloo on CPU.
"""
from __future__ import annotations


def shard_range(bb: int, rank: int, world: int) -> tuple[int, int]:
    """Column range of `rank`: mirrors cgpu_shard_range (consel_b200/csrc/capi.cu)."""
    return bb * rank // world, bb * (rank + 1) // world


def item_range(cm: int, rank: int, world: int) -> tuple[int, int]:
    """(row0, nrows) of the items `rank` owns: mirrors cgpu_comm_item_range (csrc/comm.cu):
    equal ranges of ceil(cm/world) rows, the tail ranks may own fewer (or none)."""
    per = (cm + world - 1) // world
    r0 = min(cm, rank * per)
    return r0, min(cm, r0 + per) - r0
