"""One rank of a multi-process run of the EMULATED library (tests/test_emu_dist.py): the sharded
makermt -> consel pipeline exactly as a one-process-per-GPU job runs it -- cgpu_comm_init_env (TCP rendezvous
of the communicator id on MASTER_ADDR:MASTER_PORT+17, RANK / WORLD_SIZE from the environment), replicate columns
sharded, the library's own collectives (comm.cu over the stand-in libnccl.so.2) -- on the reference's example
data.  Every rank writes its result table to <out>.<rank>.npz.  No torch in this process: the library must find
the stand-in NCCL, not a framework's copy."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    libpath, out, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    import consel_b200 as cg
    from consel_b200 import api, pipeline
    from oracle import oracle as orc
    lib = C.CDLL(libpath)
    lib.cgpu_last_error.restype = C.c_char_p
    lib.cgpu_launch_count.restype = C.c_longlong
    lib.cgpu_smooth_count.restype = C.c_double
    api._lib, pipeline._host = lib, lib
    cg.init()
    api._check(lib.cgpu_comm_init_env())
    rank, world = pipeline.comm_info()
    gold = ROOT / "tests" / "golden"
    L = orc.read_lnf(gold / "mam15.lnf")
    rr = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
    bb = [100, 101, 99, 100, 103, 100, 97, 100, 100, 102]  # not all divisible by the world size
    if mode == "tinyb":  # fewer replicates than ranks at some scales: a rank's share of a scale can be EMPTY
        bb = [2, 1, 3, 2, 1, 2, 4, 1, 2, 2]
    ass = None
    if mode == "edges":
        ass = orc.read_ass(gold / "mam15.ass")
    if mode.startswith("rand"):  # tools/emu_dist_worlds.py: random shapes, scale sets, replicate counts, associations
        rng = np.random.default_rng(int(mode[4:]))
        K, n = int(rng.integers(2, 40)), int(rng.integers(5, 400))
        L = (-2.3 - 4.07 * rng.exponential(size=n))[None, :] + rng.choice([0.05, 0.3]) * rng.standard_normal((K, n))
        kk = int(rng.integers(2, 9))
        rr = np.sort(np.round(rng.uniform(0.4, 1.8, size=kk), 2))
        rr[int(rng.integers(0, kk))] = 1.0
        rr = [float(x) for x in np.sort(rr)]
        bb = [int(x) for x in rng.integers(1, 150, size=kk)]
        if rng.random() < 0.5:
            cm = int(rng.integers(1, 10))
            sets = [np.sort(rng.choice(K, size=int(rng.integers(1, K)), replace=False)) for _ in range(cm)]
            ass = (np.cumsum([0] + [len(x) for x in sets]).astype(np.int32), np.concatenate(sets).astype(np.int32))
    res = pipeline.run(L, rr, bb, 1234, sum_mode=cg.SUM_STRICT, ass=ass)
    np.savez(f"{out}.{rank}.npz", rank=rank, world=world, item=res["item"], stat=res["stat"], pv=res["pv"], se=res["se"], cnt=res["cnt"],
             ax=res["ax"], ci=res["ci"], ret=res["ret"], datvec=res["datvec"])
    lib.cgpu_comm_destroy()


if __name__ == "__main__":
    main()
