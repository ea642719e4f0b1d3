"""The N>1 path of the PRODUCT on the CPU: world_size 2 and 3, one process per rank, each running the emulated
library (tests/emu/build_emu_lib.py) with the real comm.cu over a stand-in NCCL (tests/emu/fake_nccl.cpp:
collectives between processes through shared memory).  What a one-process-per-GPU job does is done here
unchanged: cgpu_comm_init_env with its TCP rendezvous, MT19937 jump-ahead to the rank's own replicate columns,
the split row sums with their all-gather, cgpu_reps_gather_scale, the KH/SH tests split over items
(cgpu_mctest_rows), the collective count queries of the lockstep AU driver (cgpu_count_query_dist: all-gather of
thresholds, grouped reduce-scatter of count / max / min), the final all-gather of the per-item results.  Every
rank must return the table the single-process run returns, bit for bit (tests/test_gpu_dist.py checks the same
on two real GPUs)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    """a MASTER_PORT whose rendezvous port (MASTER_PORT + 17, consel_b200/csrc/comm.cu) is free as well"""
    for _ in range(64):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        if port + 17 > 65535:
            continue
        try:
            with socket.socket() as t:
                t.bind(("0.0.0.0", port + 17))
            return port
        except OSError:
            continue
    raise RuntimeError("no free port pair")


def _run_ranks(lib_path: Path, out: Path, world: int, mode: str):
    env = dict(os.environ)
    env.update(WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), CHOST_THREADS="2",
               LD_LIBRARY_PATH=f"{lib_path.parent}:{env.get('LD_LIBRARY_PATH', '')}")
    env.pop("LOCAL_RANK", None)
    procs = []
    for rank in range(world):
        e = dict(env, RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "emu" / "dist_worker.py"), str(lib_path), str(out), mode],
                                      env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            log, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(log)
    for rank, p in enumerate(procs):
        assert p.returncode == 0, f"rank {rank}:\n{logs[rank][-3000:]}"
    return [np.load(f"{out}.{rank}.npz") for rank in range(world)]


@pytest.mark.parametrize("world,mode", [(2, "identity"), (3, "identity"), (2, "edges"), (3, "tinyb")])
def test_sharded_pipeline_equals_single_process_bit_for_bit(emu_lib, tmp_path, world, mode):
    one = _run_ranks(emu_lib._path, tmp_path / "one", 1, mode)[0]
    many = _run_ranks(emu_lib._path, tmp_path / "many", world, mode)
    assert int(one["world"]) == 1
    for rank, got in enumerate(many):
        assert (int(got["rank"]), int(got["world"])) == (rank, world)
        for key in ("item", "stat", "pv", "se", "cnt", "ax", "ci", "ret", "datvec"):
            assert np.array_equal(got[key], one[key], equal_nan=True), (rank, key)


def test_makermt_and_consel_programs_as_two_ranks(emu_lib, emu_bin, gold, tmp_path):
    """the command-line programs as one process per rank (RANK / WORLD_SIZE / MASTER_PORT, no launcher framework):
    .rmt/.vt/.pv/.ci/.cnt byte-equal to the single-process run of the same programs -- what
    tests/test_gpu_dist.py::test_makermt_and_consel_programs_on_two_gpus checks on two GPUs"""
    import shutil
    for d in ("one", "two"):
        (tmp_path / d).mkdir()
        shutil.copy(gold / "mam15.lnf", tmp_path / d / "mam15.lnf")

    def run(prog, args, cwd, world):
        procs = []
        port = _free_port()
        for r in range(world):
            env = dict(os.environ, LD_LIBRARY_PATH=f"{emu_lib._path.parent}:{os.environ.get('LD_LIBRARY_PATH', '')}", CHOST_THREADS="2")
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                env.pop(k, None)
            if world > 1:
                env.update(RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([str(emu_bin / prog)] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        return outs

    for d, world in (("one", 1), ("two", 2)):
        run("makermt", ["-s", "1234", "-b", "0.05", "--paml", "mam15", "x"], tmp_path / d, world)
        outs = run("consel", ["-c", "x"], tmp_path / d, world)
        if world > 1:
            assert outs[1] == ""  # rank 0 speaks for the job
    for ext in ("rmt", "vt", "pv", "ci", "cnt"):
        a, b = (tmp_path / "one" / f"x.{ext}").read_bytes(), (tmp_path / "two" / f"x.{ext}").read_bytes()
        assert a == b, ext
