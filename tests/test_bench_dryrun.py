"""bench.py's host-side glue, end to end, without a GPU: the device library and the pipeline are
replaced by stand-ins with the shapes of consel_b200.api / consel_b200.pipeline, and run_ours() must
walk through every leg (timed steps, spot check, side modes, pipeline, host-buffer arm, rooflines,
cpu_baseline) and print ONE JSON line carrying every key the measurement contract names.  Nothing
here measures anything; it guards the report code that only runs AFTER the GPU work of a bench run.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


class _Reps:
    def __init__(self, owner, bn):
        self.owner, self.bn = owner, bn

    def free(self):
        pass

    def download_cols(self, s, col0, ncols):
        # what the device must return: the correctly rounded exact sums over the reference's draws
        o = self.owner
        skip = sum(int(o.bb[t]) * int(self.bn[t]) for t in range(s)) + col0 * int(self.bn[s])
        y = o.orc.mt_stream(7, skip, ncols * int(self.bn[s]))
        idx = o.orc.site_indices(y, o.L.shape[1]).reshape(ncols, int(self.bn[s]))
        counts = np.stack([np.bincount(idx[b], minlength=o.L.shape[1]) for b in range(ncols)])
        return bench.exact_sums(o.L, counts, float(self.bn[s]) / o.L.shape[1])


class _Matrix:
    def __init__(self, L, orc):
        self.L, self.orc, self.bb = np.asarray(L), orc, None

    def scaleboot(self, bn, bb, seed, stream_offset=0, rng_mode=0, sum_mode=0):
        self.bb = np.asarray(bb)
        return _Reps(self, np.asarray(bn))

    def info(self):
        K, n = self.L.shape
        return dict(K=K, n=n, NP=(n + 127) // 128 * 128, slices=8, frac_bits=50, mpad=128, patterns=0)

    def free(self):
        pass


@pytest.mark.parametrize("clocks_low", [False, True])
def test_run_ours_prints_the_contract_line(monkeypatch, capsys, orc, clocks_low):
    import torch

    import consel_b200 as cg
    from consel_b200 import pipeline

    # ---- torch.cuda stand-ins
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(multi_processor_count=148))
    real_empty = torch.empty

    def empty(*a, **kw):
        kw.pop("device", None)
        kw.pop("pin_memory", None)
        return real_empty(*a, **kw)
    monkeypatch.setattr(torch, "empty", empty)

    # ---- device library stand-ins (shapes of consel_b200/api.py)
    timing = dict(upload=0.0, rng=0.6, resample=0.4, contract=0.0, download=0.0, total=2.0, gemm_i8=0.9, fx_combine=0.0,
                  mt_jump=0.2, mt_draw=0.4, n_gemm=10.0, n_hist=10.0)
    launches = [0]

    def launch_count():
        launches[0] += 31
        return launches[0]
    monkeypatch.setattr(cg, "init", lambda d=-1: None)
    monkeypatch.setattr(cg, "set_shard", lambda r, w: None)
    monkeypatch.setattr(cg, "shard_range", lambda b: (0, int(b)))
    monkeypatch.setattr(cg, "bench_peaks", lambda sustained_seconds=0.0: dict(
        int8_tops=4500.0, fp64_tflops=33.0, **({"int8_tops_sustained": 3800.0} if sustained_seconds > 0 else {})))
    monkeypatch.setattr(cg, "Matrix", lambda L: _Matrix(L, orc))
    monkeypatch.setattr(cg, "last_timing", lambda: dict(timing))
    monkeypatch.setattr(cg, "launch_count", launch_count)
    monkeypatch.setattr(cg, "set_option", lambda k, v: None)
    monkeypatch.setattr(cg, "kernel_times", lambda: {"col_top2_kernel": (11, 0.3, 1.3e8), "pair_weight_kernel": (1, 0.2, 3.0e7),
                                                     "unknown_kernel": (1, 0.1, 1.0)})

    def scaleboot_all(L, bn, bb, seed, stream_offset=0, rng_mode=0, sum_mode=0, out=None):
        for o in out:
            o[:] = 1.0
        return out
    monkeypatch.setattr(cg, "scaleboot_all", scaleboot_all)

    def run(L, rr, bb, seed, rng_mode=0, sum_mode=0, matrix=None, **kw):
        K = np.asarray(L).shape[0]
        pv = np.full((K, 8), 0.5)
        return dict(item=np.arange(K), pv=pv, timing={"mctest_s": 0.001, "stats_s": 0.001, "au_s": 0.002, "makermt_s": 0.002,
                                                       "device_ms": dict(timing)})
    monkeypatch.setattr(pipeline, "run", run)
    monkeypatch.setattr(pipeline, "comm_init_from_torch", lambda: (0, 1))

    class _Sampler:
        def __init__(self, idx):
            pass

        def start(self):
            pass

        def stop(self):
            return {"sm_mhz": 1400.0 if clocks_low else 1965.0, "sm_max_mhz": 1965.0,
                    "reasons": ["sw_power_cap"] if clocks_low else [], "samples": 9}
    monkeypatch.setattr(bench, "ClockSampler", _Sampler)
    # the CPU baseline leg is the real oracle; keep its sample small here
    monkeypatch.setattr(bench, "CPU_BASELINE_ACC", 15 * 3414 * 10 * 50.0)  # 50 replicates per scale
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)

    args = types.SimpleNamespace(gpus=1, steps=2, warmup=3, impl="ours", config="C1")
    bench.run_ours(args)
    lines = [t for t in capsys.readouterr().out.splitlines() if t.strip()]
    assert len(lines) == 1, "bench.py must print exactly one JSON line"
    d = json.loads(lines[0])

    # ---- the base contract
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == bench.METRIC and d["unit"] == "acc/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["ms_per_step"] == pytest.approx(2.0) and d["value"] == pytest.approx(15 * 3414 * 1e5 / 2.0e-3)
    assert d["gpu_launches"] > 0
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 15 * 3414 * 8 and e2e["d2h_bytes_per_step"] == 15 * 10000 * 8 * 10 and e2e["value"] > 0
    # ---- roofline of the dominant kernel + every kernel + the step-level floor
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernels", "step"):
        assert key in r, key
    assert r["bound"] in ("hbm", "tensor") and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["kernel"].startswith("gemm_i8")  # the largest share in the stand-in timing
    # power-capped steps are graded against the sustained int8 rate, full-clock steps against the burst rate
    assert r["peak"] == (3800.0 if clocks_low else 4500.0)
    names = [k["kernel"] for k in r["kernels"]]
    assert any("mt_draw_kernel" in n for n in names) and any("col_top2_kernel" in n for n in names)
    assert not any("unknown_kernel" in n for n in names)
    assert 0.0 < r["step"]["frac"] and r["step"]["ms_per_step"] == pytest.approx(2.0)
    # ---- cpu baseline, spot check, side modes, pipeline
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if (orc.REF_DIR / "libmakermt_ref.so").exists() else "port")
    assert cb["cores"] == 1 and cb["value"] > 0 and "50 replicates per scale" in cb["sample"]
    assert d["spot_check"]["ok"] is True and d["spot_check"]["bit_equal_to_exact_rounded_sums"] is True
    assert set(d["modes"]) == {"strict_mt", "fast_mt", "int8_philox"}
    p = d["pipeline"]
    assert "error" not in p, p
    assert p["items"] == 15 and p["host_matrix_table_identical"] is True and p["makermt_consel_from_host_matrix_ms"] > 0


def test_reference_arm_non_zero_ranks_do_nothing(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "8")
    bench.run_reference(types.SimpleNamespace(gpus=8, steps=1, warmup=0, impl="reference", config="C3"))
    assert capsys.readouterr().out == ""


def test_auxiliary_legs_cannot_cost_the_headline_line(monkeypatch, capsys, orc):
    """side modes, the host-buffer arm, the sustained-peak measurement and the CPU baseline fail: the line is still
    printed, with the headline value and the failures named"""
    import torch

    import consel_b200 as cg
    from consel_b200 import pipeline
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(multi_processor_count=148))
    real_empty = torch.empty

    def empty(*a, **kw):
        kw.pop("device", None)
        kw.pop("pin_memory", None)
        return real_empty(*a, **kw)
    monkeypatch.setattr(torch, "empty", empty)
    timing = dict(upload=0.0, rng=0.6, resample=0.4, contract=0.0, download=0.0, total=2.0, gemm_i8=0.9, fx_combine=0.0,
                  mt_jump=0.2, mt_draw=0.4, n_gemm=10.0, n_hist=10.0)

    class Mat(_Matrix):
        def scaleboot(self, bn, bb, seed, stream_offset=0, rng_mode=0, sum_mode=0):
            if (rng_mode, sum_mode) != (cg.RNG_MT19937, cg.SUM_INT8):
                raise cg.CgpuError(-3, "side mode broke")
            return super().scaleboot(bn, bb, seed, stream_offset, rng_mode, sum_mode)

    def peaks(sustained_seconds=0.0):
        if sustained_seconds > 0:
            raise cg.CgpuError(-3, "no sustained peak")
        return dict(int8_tops=4500.0, fp64_tflops=33.0)

    def boom(*a, **k):
        raise cg.CgpuError(-3, "host-buffer call broke")
    monkeypatch.setattr(cg, "init", lambda d=-1: None)
    monkeypatch.setattr(cg, "set_shard", lambda r, w: None)
    monkeypatch.setattr(cg, "shard_range", lambda b: (0, int(b)))
    monkeypatch.setattr(cg, "bench_peaks", peaks)
    monkeypatch.setattr(cg, "Matrix", lambda L: Mat(L, orc))
    monkeypatch.setattr(cg, "last_timing", lambda: dict(timing))
    monkeypatch.setattr(cg, "launch_count", lambda: 31)
    monkeypatch.setattr(cg, "set_option", lambda k, v: None)
    monkeypatch.setattr(cg, "kernel_times", lambda: {})
    monkeypatch.setattr(cg, "scaleboot_all", boom)
    monkeypatch.setattr(pipeline, "run", boom)
    monkeypatch.setattr(bench, "ClockSampler", lambda idx: types.SimpleNamespace(start=lambda: None, stop=lambda: None))
    monkeypatch.setattr(orc, "makermt", boom)
    monkeypatch.setattr(bench, "_ref_worker", boom)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    # the side leg of the default workload (its own process on the GPU box) is walked too: here the default is C1
    # and the child is a stub, what is checked is the parent's branch (buffers released, result attached)
    monkeypatch.setattr(bench, "DEFAULT_CONFIG", "C1")
    monkeypatch.setattr(bench, "side_c4_from", lambda a, w=1: {"status": "stub", "variants": []})
    bench.run_ours(types.SimpleNamespace(gpus=1, steps=2, warmup=3, impl="ours", config="C1"))
    lines = [t for t in capsys.readouterr().out.splitlines() if t.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["side_c4"] == {"status": "stub", "variants": []}
    assert d["value"] == pytest.approx(15 * 3414 * 1e5 / 2.0e-3) and d["spot_check"]["ok"] is True
    assert "error" in d["e2e"] and "error" in d["cpu_baseline"] and "error" in d["pipeline"]
    assert all("error" in m for m in d["modes"].values())
    assert d["roofline"]["peak"] == 4500.0 and d["clocks"] is None


# ---------------------------------------------------------------------------------------------- the C4 side leg
def test_side_c4_leg_on_the_emulated_library(monkeypatch, capsys, emu_cg):
    """bench.run_side_c4 end to end on the CPU (emulated device library, a few trees / replicates, 300000 sites so
    that the rows are longer than every shared-memory histogram): both variants print a line, the default one is
    bit-equal to the exact sums over the oracle's draws, the bucketed one identical to the default."""
    import types

    import torch

    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(emu_cg, "init", lambda d=-1: None)
    bench.run_side_c4(types.SimpleNamespace(steps=1, side_reps=3, side_sites=300000, side_trees=3))
    lines = [json.loads(t) for t in capsys.readouterr().out.splitlines() if t.startswith("{")]
    assert [x["variant"] for x in lines] == ["setup", "global_atomics (default)", "bucketed (hist_bucket=1)"]
    assert "error" not in lines[1] and "error" not in lines[2], lines
    assert lines[1]["spot_check"]["bit_equal_to_exact_rounded_sums"] is True
    assert lines[2]["identical_to_default_histogram"] is True
    assert lines[1]["value"] > 0 and lines[2]["phases_ms_per_step_rank0"]["n_hist"] > 0


def test_side_c4_collector_keeps_what_was_printed(monkeypatch, tmp_path):
    """the parent side: a child that prints one variant and then dies (or hangs past the limit) costs nothing but
    its own missing lines"""
    import subprocess as sp
    import types

    real_popen = sp.Popen
    script = tmp_path / "child.py"

    def fake(mode):
        def popen(cmd, **kw):
            script.write_text("import json, sys, time\nprint(json.dumps({'side': 'C4', 'variant': 'a', 'value': 1.0}), flush=True)\n"
                              + ("time.sleep(60)\n" if mode == "hang" else "sys.stderr.write('boom\\n'); sys.exit(3)\n"))
            return real_popen([sys.executable, str(script)], **kw)
        return popen
    monkeypatch.setenv("BENCH_SIDE_TIMEOUT_S", "3")
    for mode, word in (("die", "exit status 3: boom"), ("hang", "time limit")):
        monkeypatch.setattr(bench.subprocess, "Popen", fake(mode))
        out = bench.side_c4_from(types.SimpleNamespace(), 8 if mode == "hang" else 1)
        assert word in out["status"], out
        assert out["variants"] == [{"side": "C4", "variant": "a", "value": 1.0}]


def test_side_c4_leg_as_two_ranks_on_the_emulated_library(emu_lib, tmp_path):
    """the multi-rank form of the leg (the 8-rank job the 8-GPU bench starts), here as 2 ranks over gloo on the
    CPU: every rank resamples its own replicate columns, rank 0 prints, the spot check and the bucketed == default
    comparison hold"""
    import os
    import socket
    import subprocess as sp

    worker = tmp_path / "worker.py"
    worker.write_text(f"""
import ctypes as C, sys, types
sys.path.insert(0, {str(ROOT)!r})
import torch
torch.cuda.set_device = lambda d: None
import consel_b200 as cg
from consel_b200 import api
lib = C.CDLL({str(emu_lib._path)!r})
lib.cgpu_last_error.restype = C.c_char_p
lib.cgpu_launch_count.restype = C.c_longlong
lib.cgpu_smooth_count.restype = C.c_double
api._lib = lib
real_init = cg.init
import bench
bench_cg_init = lambda d=-1: real_init()
cg.init = bench_cg_init
bench.run_side_c4(types.SimpleNamespace(steps=1, side_reps=6, side_sites=300000, side_trees=3))
""")
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), LOCAL_WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(sp.Popen([sys.executable, str(worker)], env=env, stdout=sp.PIPE, stderr=sp.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert not [t for t in outs[1][0].splitlines() if t.startswith("{")], "only rank 0 prints"
    lines = [json.loads(t) for t in outs[0][0].splitlines() if t.startswith("{")]
    assert [x["variant"] for x in lines] == ["setup", "global_atomics (default)", "bucketed (hist_bucket=1)"]
    assert lines[0]["n_gpus"] == 2 and lines[0]["replicates_per_scale_per_gpu"] == 3
    assert "error" not in lines[1] and "error" not in lines[2], lines
    assert lines[1]["spot_check"]["bit_equal_to_exact_rounded_sums"] is True
    assert lines[2]["identical_to_default_histogram"] is True
