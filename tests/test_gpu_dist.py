"""Multi-GPU parity (needs two or more GPUs; skipped on a one-GPU box): makermt+consel with the
replicates sharded over the ranks and every collective issued by the native libraries (NCCL
communicator of libconselgpu.so, chost_consel_run) equals the single-GPU table bit for bit.
Plus the single-rank behaviour of the same entry points."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_sharded_pipeline_equals_single_gpu():
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU on this box")
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "dist_check.py"), "2"], capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-3000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("== single GPU: True") == 8  # 2 ranks x 2 cases x 2 sum modes


def test_makermt_and_consel_programs_on_two_gpus(tmp_path, gold):
    """the command-line programs as one process per GPU (RANK / WORLD_SIZE / MASTER_PORT, no launcher
    framework): .rmt/.vt/.pv/.ci/.cnt byte-equal to the single-GPU run of the same programs"""
    import os
    import shutil
    if _ngpu() < 2:
        pytest.skip("one GPU on this box")
    bindir = ROOT / "consel_b200" / "bin"
    for d in ("one", "two"):
        (tmp_path / d).mkdir()
        shutil.copy(gold / "mam15.lnf", tmp_path / d / "mam15.lnf")

    def run(prog, args, cwd, world):
        procs = []
        port = 21000 + os.getpid() % 20000
        for r in range(world):
            env = dict(os.environ)
            if world > 1:
                env.update(RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            else:
                for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                    env.pop(k, None)
            procs.append(subprocess.Popen([str(bindir / prog)] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        return outs

    for d, world in (("one", 1), ("two", 2)):
        run("makermt", ["-s", "1234", "-b", "0.3", "--paml", "mam15", "x"], tmp_path / d, world)
        outs = run("consel", ["-c", "x"], tmp_path / d, world)
        if world > 1:
            assert outs[1] == ""  # rank 0 speaks for the job
    for ext in ("rmt", "vt", "pv", "ci", "cnt"):
        a, b = (tmp_path / "one" / f"x.{ext}").read_bytes(), (tmp_path / "two" / f"x.{ext}").read_bytes()
        assert a == b, ext


@pytest.fixture(scope="module")
def cg():
    import consel_b200 as cg
    cg.init()
    cg.set_shard(0, 1)
    return cg


def test_collective_entry_points_without_communicator(cg, orc, gold):
    """world = 1: cgpu_count_query_dist == cgpu_count_query, gather_scale copies, item range is everything"""
    from consel_b200 import api
    lib = cg.load()
    r, w = C.c_int(-1), C.c_int(-1)
    assert lib.cgpu_comm_info(C.byref(r), C.byref(w)) == 0 and (r.value, w.value) == (0, 1)
    rmt = orc.read_rmt(gold / "g100.rmt")
    K = len(rmt["datvec"])
    ptr, idx = orc.identity_assoc(K)
    _, _, sptr, sidx = orc.order_items(rmt["datvec"], ptr, idx)
    reps = cg.Reps.upload(rmt["reps"])
    reps.stats_build(sptr, sidx)
    thr = np.random.default_rng(5).normal(size=(reps.cm, 3))
    want = reps.count_query(thr)
    got = np.zeros_like(want)
    anyf = C.c_int(-1)
    api._check(lib.cgpu_count_query_dist(reps._h, 0, 3, api._dp(np.ascontiguousarray(thr)), got.ctypes.data_as(C.POINTER(api.RowSum)), 1, C.byref(anyf)))
    assert anyf.value == 1 and got.tobytes() == want.tobytes()
    h = C.c_void_p()
    api._check(lib.cgpu_reps_gather_scale(reps._h, 2, C.byref(h)))
    whole = api.Reps(h, K, rmt["bb"][2:3])
    assert np.array_equal(whole.download()[0], rmt["reps"][2])
    whole.free()


@pytest.mark.parametrize("name", ["g100", "g100e"])
def test_mctest_item_ranges_compose(cg, orc, gold, name):
    """cgpu_mctest_rows over a partition of the items == cgpu_mctest (what every rank of a sharded run computes)"""
    from consel_b200 import api
    lib = cg.load()
    rmt = orc.read_rmt(gold / "g100.rmt")
    K = len(rmt["datvec"])
    ptr, idx = orc.read_ass(gold / "mam15.ass") if name == "g100e" else orc.identity_assoc(K)
    _, _, sptr, sidx = orc.order_items(rmt["datvec"], ptr, idx)
    sptr = np.ascontiguousarray(sptr, dtype=np.int32)
    sidx = np.ascontiguousarray(sidx, dtype=np.int32)
    cm = len(sptr) - 1
    i0 = orc.find_i0(rmt["rr"])
    reps = cg.Reps.upload(rmt["reps"])
    want = reps.mctest(i0, rmt["datvec"], sptr, sidx)
    dat = np.ascontiguousarray(rmt["datvec"], dtype=np.float64)
    for cuts in ([0, cm], [0, 1, cm // 2, cm], [0, cm // 3, cm // 3, cm - 1, cm]):
        outs = [np.full(cm, -7.0) for _ in range(4)]
        for e0, e1 in zip(cuts[:-1], cuts[1:]):
            api._check(lib.cgpu_mctest_rows(reps._h, int(i0), api._dp(dat), C.c_double(1.0), cm, api._ip(sptr), api._ip(sidx),
                                            int(e0), int(e1), *[api._dp(o) for o in outs]))
        for a, b in zip(outs, want):
            assert np.array_equal(a, b), cuts
