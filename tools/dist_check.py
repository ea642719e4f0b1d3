"""torchrun --nproc-per-node N tools/dist_check.py : sharded makermt+consel over N GPUs must
reproduce the single-GPU table (run by hand under `gpurun --gpus N`; development aid)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cg.init(local)
rng = np.random.Generator(np.random.PCG64(20261019))
K, n, B = 60, 2000, 4000
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n
RR = [0.5, 0.7, 0.9, 1.0, 1.1, 1.3]
for sm in (cg.SUM_STRICT, cg.SUM_INT8):
    cg.set_shard(rank, world)
    t = time.perf_counter()
    res = pipeline.run(L, RR, [B] * len(RR), 11, sum_mode=sm)
    dt = time.perf_counter() - t
    if rank == 0:
        # single-GPU answer computed by rank 0 alone (other ranks idle at the barrier below)
        cg.set_shard(0, 1)
        save = pipeline._distributed
        pipeline._distributed = lambda: False
        ref = pipeline.run(L, RR, [B] * len(RR), 11, sum_mode=sm)
        pipeline._distributed = save
        ok = all(np.array_equal(res[k], ref[k]) for k in ("item", "cnt")) and np.array_equal(res["pv"], ref["pv"]) \
            and np.array_equal(res["ax"], ref["ax"]) and np.array_equal(res["ci"], ref["ci"])
        print(f"sum_mode {sm}: sharded over {world} GPUs == single GPU: {ok}  ({dt*1e3:.1f} ms)", flush=True)
        assert ok
    dist.barrier()
dist.destroy_process_group()
