"""Sharded makermt+consel over N GPUs must reproduce the single-GPU table bit for bit.

    python tools/dist_check.py N            # spawns N ranks itself (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_PORT)
    torchrun --nproc-per-node N tools/dist_check.py

No torch.distributed: the ranks meet through cgpu_comm_init_env (TCP rendezvous of the NCCL id in
libconselgpu.so) and every collective is issued by the native libraries.  Each rank first computes
the single-GPU answer alone (shard 0/1), then the sharded one; rank r compares its copy of the full
table.  tests/test_gpu_dist.py runs this under pytest when the box has two or more GPUs."""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def rank_main():
    import consel_b200 as cg
    from consel_b200 import api, pipeline

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    cg.init(local)
    rng = np.random.Generator(np.random.PCG64(20261019))
    cases = [(60, 2000, 4000, [0.5, 0.7, 0.9, 1.0, 1.1, 1.3], None), (37, 300, 1001, [0.6, 1.0, 1.4], "groups")]
    refs = []
    for K, n, B, RR, assoc in cases:  # single-GPU answers first (no communicator yet)
        base = -2.3 - 4.07 * rng.exponential(size=n)
        L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
        ass = None
        if assoc:  # a few multi-tree items next to singletons (edge associations)
            groups = [[0, 3, 5], [1], [2, 4], [6, 7, 8, 9], [10]] + [[k] for k in range(11, K)]
            ass = (np.cumsum([0] + [len(g) for g in groups]).astype(np.int32), np.concatenate(groups).astype(np.int32))
        for sm in (cg.SUM_STRICT, cg.SUM_INT8):
            refs.append((L, RR, B, ass, sm, pipeline.run(L, RR, [B] * len(RR), 11, sum_mode=sm, ass=ass)))
    api._check(cg.load().cgpu_comm_init_env())
    assert pipeline.comm_info() == (rank, world)
    ok_all = True
    for L, RR, B, ass, sm, ref in refs:
        t = time.perf_counter()
        res = pipeline.run(L, RR, [B] * len(RR), 11, sum_mode=sm, ass=ass)
        dt = time.perf_counter() - t
        ok = all(np.array_equal(res[k], ref[k]) for k in ("item", "cnt", "pv", "se", "ret")) and \
            np.array_equal(res["ax"], ref["ax"], equal_nan=True) and np.array_equal(res["ci"], ref["ci"], equal_nan=True)
        print(f"rank {rank}: K={L.shape[0]} n={L.shape[1]} B={B} sum_mode {sm} {'groups' if ass else 'identity'}: "
              f"sharded over {world} GPUs == single GPU: {ok}  ({dt*1e3:.1f} ms)", flush=True)
        ok_all = ok_all and ok
    cg.load().cgpu_comm_destroy()
    sys.exit(0 if ok_all else 1)


def main():
    if "RANK" in os.environ:
        rank_main()
        return
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    port = 20000 + os.getpid() % 20000
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)], env=env))
    rcs = [p.wait(timeout=900) for p in procs]
    print("dist_check exit codes", rcs)
    sys.exit(max(rcs))


if __name__ == "__main__":
    main()
