#!/usr/bin/env python
"""bench.py — tree-site accumulations per second of the multiscale RELL bootstrap.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): acc/s = K*n*sum_s(B_s) / t, where one "step" is one full
makermt resampling pass (all scales) over one synthetic site-log-likelihood matrix.
Workload = BASELINE.json configs[1] ("C2"): K=105 trees x n=3414 sites, 10 scales
r=0.5..1.4, 100000 replicates per scale PER GPU (replicates are sharded over ranks with
no data-path collective, so N GPUs resample N x 100000 replicates per scale: weak
scaling).  Headline mode = the north star's parity contract: the reference's own seeded
MT19937 index stream (site counts bit-exact) + exact fixed-point sums on the tcgen05 int8
tensor cores (replicate sums = the correctly rounded exact sums, <= 1e-12 relative to
`makermt -s 7`).  The line also carries `modes`: the bit-identical strict fp64 mode, the
fast fp64 mode and the counter-based (Philox) resampler, each timed the same way.

  value  device-resident: L already in HBM, replicate blocks left in HBM
         (cgpu_scaleboot_dev), timed with CUDA events on the library's stream.
  e2e    through the drop-in C-ABI call cgpu_scaleboot_all with HOST buffers:
         H2D of L, compute, D2H of the replicate blocks, all inside the timed region.

--impl reference times the UNMODIFIED reference program (oracle/_ref/makermt, built
from /root/reference/src by oracle/Makefile) on the host cores with the reference's
own parallel recipe (one process per scale, doc: src/program.tex:1089-1109), on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "bootstrap tree-site accumulations/sec (K*n*B)"
UNIT = "acc/s"
RR = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
K, N_SITES, B_PER_SCALE = 105, 3414, 100000


def synth_matrix(k: int, n: int, seed: int = 20261019) -> np.ndarray:
    """mam15-like site log-likelihoods."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    return np.ascontiguousarray(base[None, :] + 0.106 * rng.standard_normal((k, n)) - 0.5 * np.arange(k)[:, None] / n)


def config(world: int) -> dict:
    return {
        "workload": "C2 (BASELINE.json configs[1]): synthetic K=105 trees x n=3414 sites, 10 scales "
                    "r=0.5..1.4 x 100000 replicates per scale per GPU",
        "K": K, "n": N_SITES, "scales": len(RR), "replicates_per_scale_per_gpu": B_PER_SCALE,
        "replicates_per_scale_total": B_PER_SCALE * world,
        "rng_mode": "mt19937 reference stream (seed 7), device jump-ahead: site counts bit-exact",
        "sum_mode": "int8: exact fixed-point sums on tcgen05 tensor cores, one rounding (<=1e-12 rel. to the "
                    "reference; `modes.strict_mt` is the bit-identical fp64 mode)",
        "l2": "per-scale working set (index buffer 1.9 GB + counts 0.35 GB + replicate block 84 MB) "
              "exceeds the 126 MB L2; a 256 MB buffer is also written between timed steps",
    }


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        # NVML polls in well under a millisecond, so even a 50 ms timed region yields tens of
        # samples; nvidia-smi (one process per sample) is the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self._halt.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append(["nvml", str(sm), str(mx), "", ""] + ["Active" if r & b else "Not Active" for b in bits.values()])
                self._halt.wait(0.002)
            return
        except Exception:
            pass
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------- ours
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import consel_b200 as cg

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cg.init(local)
    cg.set_shard(rank, world)

    L = synth_matrix(K, N_SITES)
    _, bn, _ = cg.scale_setup(L, RR)
    bb = np.full(len(RR), B_PER_SCALE * world, dtype=np.int32)
    acc_per_step_total = float(K) * N_SITES * float(bb.sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm
    mat = cg.Matrix(L)
    flush_ctr = [0]

    def timed_steps(rng_mode, sum_mode, steps, warmup):
        for w in range(warmup):
            mat.scaleboot(bn, bb, 7, rng_mode=rng_mode, sum_mode=sum_mode).free()
        dev, wall, ph = 0.0, 0.0, {}
        for step in range(steps):
            flush_ctr[0] += 1
            flush.fill_(flush_ctr[0] & 0xFF)
            barrier()
            t0 = time.perf_counter()
            reps = mat.scaleboot(bn, bb, 7, rng_mode=rng_mode, sum_mode=sum_mode)  # returns after the stream drained
            torch.cuda.synchronize()
            wall += (time.perf_counter() - t0) * 1e3
            tm = cg.last_timing()
            dev += tm["total"]
            for k_, v in tm.items():
                ph[k_] = ph.get(k_, 0.0) + v / steps
            reps.free()
        barrier()
        return max_over_ranks(dev), max_over_ranks(wall), ph

    HEAD = (cg.RNG_MT19937, cg.SUM_INT8)
    timed_steps(*HEAD, 0, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = cg.launch_count()
    dev_ms, wall_ms, phases = timed_steps(*HEAD, args.steps, 0)
    launches = cg.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    modes = {}
    side_steps = max(1, min(args.steps, 3))
    for name, rm, sm in (("strict_mt", cg.RNG_MT19937, cg.SUM_STRICT), ("fast_mt", cg.RNG_MT19937, cg.SUM_FAST),
                         ("int8_philox", cg.RNG_PHILOX, cg.SUM_INT8)):
        d, w_, ph = timed_steps(rm, sm, side_steps, 2)
        modes[name] = {"ms_per_step": d / side_steps, "value": acc_per_step_total * side_steps / (d * 1e-3),
                       "phases_ms_per_step": ph}

    # ---- the metric's second half: makermt+consel wall time, in one process (no .rmt round trip),
    # same workload; sharded over the ranks when N>1 (row summaries combined over NCCL)
    from consel_b200 import pipeline
    pipe = None
    try:
        pipeline.run(L, RR, bb, 7, sum_mode=cg.SUM_INT8, matrix=mat)  # warm-up: host pool, stat buffers
        pw = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            res = pipeline.run(L, RR, bb, 7, sum_mode=cg.SUM_INT8, matrix=mat)
            pw.append((time.perf_counter() - t0) * 1e3)
        barrier()
        pipe = {"makermt_consel_wall_ms": max_over_ranks(float(np.median(pw))),
                "phases_ms": {k_: round(v * 1e3, 2) for k_, v in res["timing"].items() if isinstance(v, float)},
                "what": "pipeline.run: MT19937 + int8 resampling, KH/SH/wKH/wSH tests, BP, AU fit and confidence "
                        "limits for all K items; replicate blocks stay on the device"}
    except Exception as e:  # the headline line must still be printed
        pipe = {"error": f"{type(e).__name__}: {e}"}

    # ---- end-to-end arm: host buffers through cgpu_scaleboot_all
    Lh = torch.empty((K, N_SITES), dtype=torch.float64, pin_memory=True)
    Lh.numpy()[:] = L
    b0, b1 = cg.shard_range(int(bb[0]))
    out_t = [torch.empty((K, b1 - b0), dtype=torch.float64, pin_memory=True) for _ in bb]  # rank-local blocks
    outs = [t.numpy() for t in out_t]
    h2d = int(L.nbytes)
    d2h = int(K * (b1 - b0) * 8 * len(RR))
    e2e_steps = max(1, min(args.steps, 3))
    cg.scaleboot_all(Lh.numpy(), bn, bb, 7, sum_mode=cg.SUM_INT8, out=outs)  # warm-up (page-locks, pool growth)
    e2e_ms = 0.0
    for step in range(e2e_steps):
        flush.fill_(step & 0xFF)
        barrier()
        t0 = time.perf_counter()
        cg.scaleboot_all(Lh.numpy(), bn, bb, 7, sum_mode=cg.SUM_INT8, out=outs)
        checksum = float(outs[0][0, 0])  # the step's result is read on the host
        e2e_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    assert np.isfinite(checksum)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines, measured live: every kernel of the step with its own bound; the top-level
    # `roofline` object is the kernel with the largest share of the step
    import ctypes as C
    tf = C.c_double()
    cg.api._check(cg.load().cgpu_bench_fp64_peak(C.byref(tf)))
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    traffic_db = {}
    tp = ROOT / "profiles" / "kernel_traffic.json"
    if tp.exists():
        try:
            traffic_db = json.loads(tp.read_text())
        except Exception:
            traffic_db = {}
    b_loc = b1 - b0
    S_SLICES = 8   # fixed-point slices of this matrix (57-bit data: see DESIGN.md 4.4)
    M_PAD = (S_SLICES * K + 127) // 128 * 128
    NP = (N_SITES + 127) // 128 * 128
    draws = float(sum(int(x) for x in bn)) * b_loc       # this rank's draws per step
    idx_b = 2.0 if N_SITES <= 65536 else 4.0
    nS = len(RR)

    def entry(name, kernel, ms_step, launches, bound, alg_per_launch, peak, unit, algorithmic, note=None):
        t_launch = max(ms_step, 1e-9) / launches * 1e-3
        ach = alg_per_launch / t_launch / (1e12 if bound == "tensor" else 1e9)
        e = {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
             "launches_per_step": launches, "us_per_launch": t_launch * 1e6, "algorithmic": algorithmic,
             "share_of_step": ms_step / max(phases["total"], 1e-9),
             "traffic": traffic_db.get(name)}
        if note:
            e["note"] = note
        return e

    kernels = [
        entry("mt_draw_kernel", "mt_draw_kernel<int path, uint16, 3 chunks/CTA> (MT19937 block regeneration + tempering + site index)",
              phases["mt_draw"], 1, "hbm", idx_b * draws, hbm_peak, "GB/s",
              "2 B written per draw (uint16 site index), all scales of the step in one launch",
              "integer-issue bound, not HBM bound: ~0.82 warp-instruction per draw at ~64 % of the issue slots "
              "(profiles/); the HBM fraction is reported because the contract asks for one of hbm|tensor"),
        entry("hist_group_kernel", "hist_group_kernel<uint16 idx, smem atomics, byte rows>", phases["resample"], nS, "hbm",
              (idx_b * draws + float(NP) * b_loc * nS) / nS, hbm_peak, "GB/s",
              "2 B read per draw + NP count bytes written per replicate, one launch per scale",
              "shared-memory atomic bound (one atomic per draw)"),
        entry("gemm_i8_persistent_kernel", "gemm_i8_persistent_kernel (tcgen05.mma kind::i8, 128x256x128 tiles, TMA SW128 ring, "
              "2 TMEM accumulator stages, exact int128 recombination + rounding + /r fused into the epilogue)", phases["gemm_i8"], nS, "tensor", 2.0 * S_SLICES * K * N_SITES * b_loc,
              2.0 * bf16_peak, "TOP/s (int8)",
              "2*S int8-ops per tree-site accumulation (SURVEY 8d), S=8 slices, x K*n*B_s per launch",
              "peak = 2 x bf16_tflops of MEASURED_PEAKS.json (dense int8 issues at twice the bf16 rate on B200; the file has no int8 figure)"),
        entry("fx_combine_kernel", "fx_combine_kernel (exact int128 recombination of the S slice sums, one rounding, /r)",
              phases["fx_combine"], nS, "hbm", 4.0 * M_PAD * b_loc + 8.0 * K * b_loc, hbm_peak, "GB/s",
              "4 B x Mpad int32 partial sums read + 8 B x K doubles written per replicate"),
        entry("mt_jump_kernel", "mt_jump_kernel (x^J mod phi applied to a 624-word window; position list in smem)",
              phases["mt_jump"], 1, "hbm", 2.0 * 2496.0 * draws / 2 ** 22, hbm_peak, "GB/s",
              "2.5 KB read + 2.5 KB written per chunk state; the work is 624 x ~10^4 shared-memory word reads per jump",
              "shared-memory bandwidth bound (~126 us per jump per SM against a 103 us floor at 128 B/clk); "
              "launches_per_step=1 stands for the ~10 levels of the jump tree"),
    ]
    # the draw kernel is bound by instruction issue, which the hbm|tensor vocabulary cannot express:
    # add the issue-slot view (warp-instructions per draw from profiles/r01_final_draw_equalfill_ncu_full.json)
    try:
        prof = json.loads((ROOT / "profiles" / "r01_final_draw_equalfill_ncu_full.json").read_text())["launches"][0]
        wi_per_draw = prof["warp_instructions"] / 3.243e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 4.0 * torch.cuda.get_device_properties(local).multi_processor_count * sm_mhz * 1e6
        ach = wi_per_draw * draws / (phases["mt_draw"] * 1e-3)
        kernels[0]["issue_view"] = {"achieved": ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s",
                                    "frac": ach / issue_peak, "warp_instructions_per_draw": wi_per_draw,
                                    "source": "instruction count from the ncu capture, time from this run's CUDA events"}
    except Exception:
        pass
    kernels = [e for e in kernels if e["share_of_step"] > 1e-6]  # e.g. fx_combine runs inside the GEMM epilogue when S divides 32
    dominant = max(kernels, key=lambda e: e["share_of_step"])
    roofline = dict(dominant)
    roofline["peak_source"] = peak_src
    roofline["step_shares"] = {k_: v / max(phases["total"], 1e-9) for k_, v in phases.items() if k_ not in ("total", "rng")}
    roofline["fp64_dfma_peak_tflops_measured_live"] = tf.value
    roofline["kernels"] = kernels

    # ---- CPU baseline: the oracle port, single thread, bounded sample
    from oracle import oracle as orc
    b_cpu = 3000
    t0 = time.perf_counter()
    orc.makermt(L, bn, [b_cpu] * len(RR), 7)
    t_cpu = time.perf_counter() - t0
    cpu_baseline = {"value": K * N_SITES * b_cpu * len(RR) / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"same matrix and scales, {b_cpu} replicates per scale ({t_cpu:.1f} s of oracle/oracle.c)"}

    line = {
        "metric": METRIC, "value": acc_per_step_total * args.steps / (dev_ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config(world),
        "timing": "CUDA events on the library stream, first to last event of each step, max over ranks",
        "wall_ms_per_step": wall_ms / args.steps, "phases_ms_per_step": phases,
        "e2e": {"value": acc_per_step_total * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "api": "cgpu_scaleboot_all(CGPU_RNG_MT19937, CGPU_SUM_INT8): pinned host L in, pinned host replicate blocks out"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "modes": modes, "pipeline": pipe,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------ reference
def run_reference(args) -> None:
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import oracle as orc
    exe = orc.ref_bin("makermt")
    if exe is None:
        try:
            orc.build(ref=True)
        except Exception:
            pass
        exe = orc.ref_bin("makermt")
    L = synth_matrix(K, N_SITES)
    cores = os.cpu_count() or 1
    # the reference's parallel recipe is one process per scale; to use every host core each
    # scale is further split into `parts` independent processes (own seed), and the
    # 10*parts jobs run through a pool of `cores` concurrent processes
    parts = max(1, -(-2 * cores // len(RR))) if cores > len(RR) else 1
    nproc = len(RR) * parts
    # ~10 s per job at ~1.5e9 acc/s per core
    b_each = max(50, int(9e9 / (K * N_SITES)))
    cfg = config(world)
    with tempfile.TemporaryDirectory(prefix="consel_ref_") as td:
        td = Path(td)
        kind = "reference"
        if exe is not None:
            orc.write_mt(td / "in.mt", L)
            for s, r in enumerate(RR):
                for p in range(parts):
                    orc.write_pa(td / f"p{s}_{p}.pa", [r], [b_each])

            def step():
                jobs = [(s, p) for s in range(len(RR)) for p in range(parts)]
                running = []
                while jobs or running:
                    while jobs and len(running) < cores:
                        s, p = jobs.pop()
                        running.append(subprocess.Popen(
                            [str(exe), "-s", str(151 + s * parts + p), "-p", f"p{s}_{p}", "in", f"o{s}_{p}"],
                            cwd=td, stdout=subprocess.DEVNULL))
                    still = []
                    for pr in running:
                        rc = pr.poll()
                        if rc is None:
                            still.append(pr)
                        elif rc != 0:
                            raise RuntimeError("reference makermt failed")
                    running = still
                    time.sleep(0.005)
            total_b = b_each * nproc
        else:
            kind = "port"
            _, bn, _ = orc.scale_setup(L, RR)

            def step():
                orc.makermt(L, bn, [b_each] * len(RR), 7)
            total_b = b_each * len(RR)
            nproc = 1
        for _ in range(min(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        t = time.perf_counter() - t0
    value = K * N_SITES * total_b * args.steps / t
    sample = (f"{nproc} `makermt -s SEED -p one_scale.pa` processes, at most {cores} at a time (reference binary, -O2), "
              f"{b_each} replicates each, same matrix read from a 17-digit .mt text file; wall clock "
              f"includes process start and text parse") if kind == "reference" else \
             f"oracle port, 1 thread, {b_each} replicates per scale"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(cores, nproc) if kind == "reference" else 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
