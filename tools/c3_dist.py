"""torchrun --nproc-per-node N tools/c3_dist.py [K n B]: makermt+consel of C3 (K=1000 trees x n=100000
sites, 10 scales x 10000 replicates), replicates sharded over the N ranks (strong scaling); wall time
of pipeline.run on rank 0 (development aid; the north star's "well under a second on 8 GPUs")."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import consel_b200 as cg
from consel_b200 import pipeline

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cg.init(local)
cg.set_shard(rank, world)
K, n, B = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1000, 100000, 10000)
rng = np.random.Generator(np.random.PCG64(20261019))
base = -2.3 - 4.07 * rng.exponential(size=n)
L = base[None, :] + 0.106 * rng.standard_normal((K, n), dtype=np.float64) - 0.5 * np.arange(K)[:, None] / n
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
mat = cg.Matrix(L)
for it in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = pipeline.run(L, RR, [B] * 10, 7, sum_mode=cg.SUM_INT8, matrix=mat)
    wall = time.perf_counter() - t0
    if rank == 0:
        print(f"iter {it}: world {world} K={K} n={n} B={B}: makermt+consel wall {wall*1e3:.1f} ms",
              {k: round(v * 1e3, 1) for k, v in res["timing"].items() if isinstance(v, float)}, flush=True)
if rank == 0:
    print("au[:3]", res["pv"][:3, 6], "bp[:3]", res["pv"][:3, 0])
if world > 1:
    dist.destroy_process_group()
