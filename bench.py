#!/usr/bin/env python
"""bench.py — tree-site accumulations per second of the multiscale RELL bootstrap, and the
makermt+consel wall time, on the configurations BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C1|C2|C3|C4|C5] [--impl ours|reference]

Metric (BASELINE.json): acc/s = K*n*sum_s(B_s) / t, one "step" = one full makermt resampling pass
(all ten scales) over one site-log-likelihood matrix.  Workloads:

    C1  example/mam15.lnf (real data): K=15 x n=3414, 10 x 10^4 replicates
    C2  synthetic K=105 x n=3414, 10 x 10^5 replicates PER GPU (weak scaling: the r01 headline)
    C3  synthetic K=1000 x n=10^5, 10 x 10^4 replicates, sharded over the ranks (strong scaling)
        -- the DEFAULT: the configuration the north star is quoted on, and it fits one GPU
    C4  synthetic K=100 x n=10^7, 10 x 10^4 replicates over 8 GPUs; with N < 8 ranks the run keeps
        the per-GPU share of the 8-GPU job (10^4 * N/8 replicates per scale) and says so
    C5  synthetic K=10^4 x n=5*10^4, 10 x 10^4 replicates (consel-kernel stress), strong scaling

Headline mode = the north star's parity contract: the reference's own seeded MT19937 index stream
(site counts bit-exact) + exact fixed-point sums on the tcgen05 int8 tensor cores (the correctly
rounded exact sums).  A few replicate columns of the timed run are checked in-run against the CPU
oracle (`spot_check`).

  value      device-resident: L already in HBM, replicate blocks left in HBM (cgpu_scaleboot_dev),
             CUDA events on the library's stream, max over ranks.
  e2e        through the drop-in C-ABI call cgpu_scaleboot_all with HOST buffers: H2D of L, compute,
             D2H of this rank's replicate blocks, all inside the timed region.
  pipeline   the metric's second half: makermt+consel wall time in one process (chost_consel_run:
             KH/SH/wKH/wSH, BP, AU + confidence limits for every item), p-value table on the host;
             sharded over the ranks through the library's NCCL communicator when N > 1.
  roofline   the kernel with the largest share of the step, `roofline.kernels` every kernel of the
             makermt AND consel sides with its own bound, `roofline.step` the whole step against
             SURVEY 8(d)'s floor (2*S*K*n*sum(B) int8-ops at the int8 peak MEASURED in this run).

  side_c4    (N=1, default workload) one GPU's share of the 8-GPU genome-scale job C4, measured in a
             process of its own after every other leg: default histogram and the bucketed one.

--impl reference times the reference's own CPU code on the host cores: the unmodified makermt
binary (one process per scale and part, src/program.tex:1089-1109) when the matrix is small enough
to pass as a .mt text file, else the reference's scaleboot() compiled unmodified from
/root/reference/src/makermt.c into oracle/_ref/libmakermt_ref.so and called in-process (a C3 .mt
file is 2.4 GB of text that every one of the processes would parse for minutes); plus the
reference consel binary on the .rmt those replicates make (`pipeline`).
"""
from __future__ import annotations

import argparse
import json
import math
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
CONFIGS = {
    "C1": dict(K=15, n=3414, B=10000, scaling="strong", what="example/mam15.lnf (real data, BASELINE.json configs[0])"),
    "C2": dict(K=105, n=3414, B=100000, scaling="weak", what="synthetic mam105-style, replicate-heavy (BASELINE.json configs[1])"),
    "C3": dict(K=1000, n=100000, B=10000, scaling="strong", what="synthetic phylogenomic (BASELINE.json configs[2])"),
    "C4": dict(K=100, n=10000000, B=10000, scaling="strong", what="synthetic genome-scale (BASELINE.json configs[3], an 8-GPU job)"),
    "C5": dict(K=10000, n=50000, B=10000, scaling="strong", what="synthetic many-trees, consel-kernel stress (BASELINE.json configs[4])"),
}
DEFAULT_CONFIG = "C3"


def synth_matrix(k: int, n: int, seed: int = 20261019) -> np.ndarray:
    """mam15-like site log-likelihoods (row blocks keep the peak host memory at one copy)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = np.empty((k, n))
    step = max(1, (1 << 24) // n)
    for k0 in range(0, k, step):
        k1 = min(k, k0 + step)
        L[k0:k1] = base[None, :] + 0.106 * rng.standard_normal((k1 - k0, n)) - 0.5 * np.arange(k0, k1)[:, None] / n
    return L


def read_paml_lnf(path: Path) -> np.ndarray:
    """PAML .lnf (the makermt --paml input, src/freadmat.c:110-154): `m n p`, per tree an id line and p
    pattern lines `idx count lnL ...`; every pattern's lnL is repeated `count` times."""
    rows = [t.split() for t in path.read_text().splitlines() if t.strip()]
    m, n, npat = (int(x) for x in rows[0][:3])
    L = np.empty((m, n))
    p = 1
    for i in range(m):
        p += 1
        col = 0
        for _ in range(npat):
            cnt = int(rows[p][1])
            L[i, col:col + cnt] = float(rows[p][2])
            col += cnt
            p += 1
    return L


def workload(name: str, world: int) -> dict:
    c = dict(CONFIGS[name])
    if name == "C4" and world < 8:
        c["B_total"] = max(8, c["B"] * world // 8)
        c["note"] = (f"C4 is an 8-GPU job (9.5e11 draws): with {world} rank(s) the run keeps the per-GPU share of it, "
                     f"{c['B_total']} replicates per scale in total")
    else:
        c["B_total"] = c["B"] * world if c["scaling"] == "weak" else c["B"]
    return c


def make_matrix(name: str, c: dict) -> np.ndarray:
    if name == "C1":
        return read_paml_lnf(ROOT / "tests" / "golden" / "mam15.lnf")
    return synth_matrix(c["K"], c["n"])


def config_json(name: str, c: dict, world: int) -> dict:
    K, n = c["K"], c["n"]
    NP = (n + 127) // 128 * 128
    idx_b = 2 if n <= 65536 else 4
    per_scale = c["B_total"] // world
    ws = (idx_b * n * per_scale + NP * per_scale + 8 * K * per_scale) / 1e9
    out = {
        "workload": f"{name}: {c['what']}: K={K} trees x n={n} sites, 10 scales r=0.5..1.4 x {c['B_total']} replicates per scale"
                    + (" in total (sharded over the ranks)" if c["scaling"] == "strong" else f" ({c['B']} per GPU)"),
        "K": K, "n": n, "scales": len(RR), "replicates_per_scale_total": c["B_total"],
        "replicates_per_scale_per_gpu": per_scale,
        "rng_mode": "mt19937 reference stream (seed 7), device jump-ahead: site counts bit-exact",
        "sum_mode": "int8: exact fixed-point sums on tcgen05 tensor cores, one rounding (<=1e-12 rel. to the reference; "
                    "`modes.strict_mt` is the bit-identical fp64 mode)",
        "l2": f"inputs larger than L2: a scale's working set (indices + count rows + replicate block) is ~{ws:.2f} GB "
              "per rank against the 126 MB L2; a 256 MB buffer is also written between timed steps",
    }
    if "note" in c:
        out["note"] = c["note"]
    return out


class ClockSampler(threading.Thread):
    """NVML clocks / throttle reasons DURING the timed region (B200_PROFILING.md); nvidia-smi fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.rows = []
        self._halt = threading.Event()
        # NVML queries take a driver lock; polled too often they delay the library's own launches
        self.period = float(os.environ.get("BENCH_SAMPLER_MS", "25")) * 1e-3

    def run(self):
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
                self._halt.wait(self.period)
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
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def exact_sums(Lrows: np.ndarray, counts: np.ndarray, r: float) -> np.ndarray:
    """Correctly rounded exact sums sum_j c[b,j]*L[k,j] (math.fsum over the drawn multiset), then /r:
    what CGPU_SUM_INT8 must return bit for bit (checker side of the in-run spot check)."""
    out = np.empty((Lrows.shape[0], counts.shape[0]))
    for b in range(counts.shape[0]):
        c = counts[b].astype(np.int64)
        for k in range(Lrows.shape[0]):
            out[k, b] = math.fsum(np.repeat(Lrows[k], c)) / np.float64(r)
    return out


# ----------------------------------------------------------------------------- ours
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import consel_b200 as cg
    from consel_b200 import pipeline

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cg.init(local)
    if world > 1:
        pipeline.comm_init_from_torch()  # the library's own NCCL communicator (sets the shard too)
    else:
        cg.set_shard(0, 1)

    live = cg.bench_peaks()  # burst int8 / fp64 rates, measured on the idle device before the workload heats it
    name = args.config
    c = workload(name, world)
    K, N_SITES = c["K"], c["n"]
    L = make_matrix(name, c)
    _, bn, _ = cg.scale_setup(L, RR)
    bb = np.full(len(RR), c["B_total"], dtype=np.int32)
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
    keep = {}

    def timed_steps(rng_mode, sum_mode, steps, warmup, keep_last=False):
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
            if keep_last and step == steps - 1:
                keep["reps"] = reps
            else:
                reps.free()
        barrier()
        return max_over_ranks(dev), max_over_ranks(wall), ph

    HEAD = (cg.RNG_MT19937, cg.SUM_INT8)
    timed_steps(*HEAD, 0, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = cg.launch_count()
    dev_ms, wall_ms, phases = timed_steps(*HEAD, args.steps, 0, keep_last=True)
    launches = cg.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- in-run spot check of the timed run's own output against the CPU oracle (rank 0: the first
    # two replicates of scales 0 and 1 -- scale 1 starts bb[0]*bn[0] draws into the MT19937 stream)
    spot = None
    if rank == 0:
        try:  # an auxiliary leg must never cost the run its headline line
            from oracle import oracle as orc
            t0 = time.perf_counter()
            ok_counts, ok_exact, worst_rel = True, True, 0.0
            ktrees = np.unique(np.linspace(0, K - 1, 3).astype(int))
            skip = 0
            for s in (0, 1):
                nrep = 2
                y = orc.mt_stream(7, skip, nrep * int(bn[s]))
                idx = orc.site_indices(y, N_SITES).reshape(nrep, int(bn[s]))
                counts = np.stack([np.bincount(idx[b], minlength=N_SITES) for b in range(nrep)])
                got = keep["reps"].download_cols(s, 0, nrep)
                r_s = float(bn[s]) / float(N_SITES)
                want = exact_sums(L[ktrees], counts, r_s)
                ok_exact = ok_exact and bool(np.array_equal(got[ktrees], want))
                seq = orc.contract_counts(L[ktrees], counts, r_s)  # the reference's sequentially rounded sums
                worst_rel = max(worst_rel, float(np.max(np.abs(got[ktrees] - seq) / np.abs(seq))))
                skip += int(bb[s]) * int(bn[s])
                if skip > 3_000_000_000:
                    break
            spot = {"ok": bool(ok_exact and worst_rel <= 1e-9), "bit_equal_to_exact_rounded_sums": ok_exact,
                    "max_rel_diff_to_reference_sequential_sums": worst_rel,
                    "what": f"replicates 0,1 of scales 0 and 1 of the LAST timed step, trees {ktrees.tolist()}: MT19937 draws and site "
                            "indices by the oracle (oracle/oracle.c), exact sums by math.fsum", "seconds": time.perf_counter() - t0}
        except Exception as e:
            spot = {"ok": False, "error": f"{type(e).__name__}: {e}"}
    keep.pop("reps").free()

    modes = {}
    heavy = acc_per_step_total / world > 2e12
    side = [("int8_philox", cg.RNG_PHILOX, cg.SUM_INT8)]
    # (the fp64 modes parallelise over trees and replicates, not over sites: at n = 10^7 with K = 100 a batch is
    # ONE CTA walking 10^7 sites -- minutes per step; they are side lines, C4 reports the int8 modes only)
    if acc_per_step_total / world <= 2e13 and N_SITES <= 1_000_000:
        side = [("strict_mt", cg.RNG_MT19937, cg.SUM_STRICT), ("fast_mt", cg.RNG_MT19937, cg.SUM_FAST)] + side
    for mname, rm, sm in side:
        is_fp64 = sm != cg.SUM_INT8
        st_ = 1 if (heavy and is_fp64) else max(1, min(args.steps, 3))
        try:
            d, w_, ph = timed_steps(rm, sm, st_, 1 if (heavy and is_fp64) else 2)
            modes[mname] = {"ms_per_step": d / st_, "value": acc_per_step_total * st_ / (d * 1e-3), "steps": st_,
                            "phases_ms_per_step": ph}
        except Exception as e:  # side lines: the headline must still be printed
            modes[mname] = {"error": f"{type(e).__name__}: {e}"}

    # pinned host copy of L (the host-buffer arms below read it)
    Lh = torch.empty((K, N_SITES), dtype=torch.float64, pin_memory=True)
    Lh.numpy()[:] = L

    # ---- the metric's second half: makermt+consel wall time in one process, same workload
    pipe = None
    ktimes = {}
    try:
        pipeline.run(L, RR, bb, 7, sum_mode=cg.SUM_INT8, matrix=mat)  # warm-up: host pool, stat buffers
        pw, res = [], None
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            res = pipeline.run(L, RR, bb, 7, sum_mode=cg.SUM_INT8, matrix=mat)
            pw.append((time.perf_counter() - t0) * 1e3)
        barrier()
        cg.set_option("ktime", 1)  # one more pass with per-kernel CUDA events on the consel side
        pipeline.run(L, RR, bb, 7, sum_mode=cg.SUM_INT8, matrix=mat)
        ktimes = cg.kernel_times()
        cg.set_option("ktime", 0)
        # the same from HOST buffers: L (pinned) uploaded inside the timed region, p-value table back on the host
        pe = []
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            res_h = pipeline.run(Lh.numpy(), RR, bb, 7, sum_mode=cg.SUM_INT8)
            pe.append((time.perf_counter() - t0) * 1e3)
        barrier()
        same_table = bool(np.array_equal(res_h["pv"], res["pv"], equal_nan=True))
        pipe = {"makermt_consel_wall_ms": max_over_ranks(float(np.median(pw))),
                "makermt_consel_from_host_matrix_ms": max_over_ranks(float(min(pe))), "host_matrix_table_identical": same_table,
                "phases_ms": {k_: round(v * 1e3, 2) for k_, v in res["timing"].items() if isinstance(v, float)},
                "items": int(len(res["item"])), "au_first3": [float(x) for x in res["pv"][:3, 6]],
                "what": "pipeline.run -> chost_consel_run: MT19937 + int8 resampling, KH/SH/wKH/wSH tests, BP, AU fit and "
                        "confidence limits for all K items; replicate blocks never leave the device, the p-value table is "
                        "returned to the host; `makermt_consel_from_host_matrix_ms` also uploads L from pinned host memory "
                        "inside the timed region" + (f"; replicates sharded over {world} ranks, collectives issued by the "
                                                  "library's NCCL communicator" if world > 1 else "")}
    except Exception as e:  # the headline line must still be printed
        pipe = {"error": f"{type(e).__name__}: {e}"}

    # ---- end-to-end arm: host buffers through cgpu_scaleboot_all
    b0, b1 = cg.shard_range(int(bb[0]))
    out_t = [torch.empty((K, b1 - b0), dtype=torch.float64, pin_memory=True) for _ in bb]  # rank-local blocks
    outs = [t.numpy() for t in out_t]
    h2d = int(L.nbytes)
    d2h = int(K * (b1 - b0) * 8 * len(RR))
    e2e_steps = max(1, min(args.steps, 3))
    e2e_error = None
    e2e_ms = 0.0
    try:
        cg.scaleboot_all(Lh.numpy(), bn, bb, 7, sum_mode=cg.SUM_INT8, out=outs)  # warm-up (page-locks, pool growth)
        checksum = 0.0
        for step in range(e2e_steps):
            flush.fill_(step & 0xFF)
            barrier()
            t0 = time.perf_counter()
            cg.scaleboot_all(Lh.numpy(), bn, bb, 7, sum_mode=cg.SUM_INT8, out=outs)
            checksum = float(outs[0][0, 0])  # the step's result is read on the host
            e2e_ms += (time.perf_counter() - t0) * 1e3
        barrier()
        e2e_ms = max_over_ranks(e2e_ms)
        if not np.isfinite(checksum):
            e2e_error = "the host-buffer call returned a non-finite replicate sum"
    except Exception as e:  # the headline line must still be printed
        e2e_error = f"{type(e).__name__}: {e}"
    if world > 1:
        # every collective of the run is behind us: all ranks tear the communicators down TOGETHER (NCCL's
        # destroy waits for its peers); rank 0 then finishes the report alone
        barrier()
        cg.load().cgpu_comm_destroy()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- rooflines, measured live: every kernel of the step with its own bound
    try:
        live.update({k_: v for k_, v in cg.bench_peaks(sustained_seconds=2.0).items() if k_ == "int8_tops_sustained"})
    except Exception:
        live["int8_tops_sustained"] = live["int8_tops"]
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    # a tensor-bound kernel inside a long step runs at the clock the power limit allows: grade it against the
    # sustained rate there (B200_PROFILING.md), against the burst rate when the steps are short
    # (evidence: the clock sampler saw the SMs below 95 % of their maximum during the timed steps)
    long_step = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] < 0.95 * clocks["sm_max_mhz"])
    i8_peak = live["int8_tops_sustained"] if long_step else live["int8_tops"]
    fp64_peak = live["fp64_tflops"]
    traffic_db = {}
    tp = ROOT / "profiles" / "kernel_traffic.json"
    if tp.exists():
        try:
            traffic_db = json.loads(tp.read_text()).get(name, {})
        except Exception:
            traffic_db = {}
    info = mat.info()
    b_loc = b1 - b0
    S_SLICES = info["slices"]
    M_PAD = info["mpad"]
    NP = info["NP"]
    draws = float(sum(int(x) for x in bn)) * b_loc       # this rank's draws per step
    idx_b = 2.0 if N_SITES <= 65536 else 4.0
    nS = len(RR)
    total_ms = max(phases["total"], 1e-9)

    def entry(kname, kernel, ms_step, nlaunch, bound, alg_per_launch, peak, unit, algorithmic, note=None, base_ms=total_ms):
        t_launch = max(ms_step, 1e-9) / max(nlaunch, 1) * 1e-3
        scale = 1e12 if bound != "hbm" else 1e9
        ach = alg_per_launch / t_launch / scale
        e = {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
             "launches_per_step": nlaunch, "us_per_launch": t_launch * 1e6, "algorithmic": algorithmic,
             "share_of_step": ms_step / base_ms, "traffic": traffic_db.get(kname)}
        if note:
            e["note"] = note
        return e

    if NP <= 14080:
        hist_name = "hist_group_kernel<smem atomics, byte rows, one warp per replicate>"
    elif NP <= 225280:
        hist_name = "hist_wide_kernel<one replicate per CTA, 16-32 warps on one shared byte histogram>"
    else:
        hist_name = "hist_global_kernel<global atomics>"
    gemm_launches = max(1, int(round(phases.get("n_gemm", nS))))
    hist_launches = max(1, int(round(phases.get("n_hist", nS))))
    kernels = [
        entry("mt_draw_kernel", f"mt_draw_kernel<{'uint16' if idx_b == 2 else 'uint32'} indices, 3 chunks/CTA> (MT19937 block regeneration + tempering + site index)",
              phases["mt_draw"], 1, "hbm", idx_b * draws, hbm_peak, "GB/s",
              f"{int(idx_b)} B written per draw (site index), all scales of the step in one launch",
              "integer-issue bound, not HBM bound (~0.8 warp-instruction per draw, profiles/); the HBM fraction is reported "
              "because the contract asks for one of hbm|tensor"),
        entry("hist_kernel", hist_name, phases["resample"], hist_launches, "hbm",
              (idx_b * draws + float(NP) * b_loc * nS) / hist_launches, hbm_peak, "GB/s",
              f"{int(idx_b)} B read per draw + NP count bytes written per replicate, one launch per (scale, batch)",
              "shared-memory atomic bound (one atomic per draw)"),
        entry("gemm_i8_persistent_kernel", "gemm_i8 kernel (tcgen05.mma kind::i8, TMA SW128 ring, TMEM accumulator stages, exact "
              "recombination + rounding + /r fused into the epilogue)", phases["gemm_i8"], gemm_launches, "tensor",
              2.0 * S_SLICES * K * N_SITES * b_loc * nS / gemm_launches, i8_peak, "TOP/s (int8)",
              f"2*S int8-ops per tree-site accumulation (SURVEY 8d), S={S_SLICES} slices, x K*n*B_s per launch",
              "peak = tcgen05.mma kind::i8 issue rate from smem-resident operands measured in this run: "
              + ("cgpu_bench_i8_peak_sustained (2 s back to back, random operand bytes: the timed steps ran power-capped, "
                 "see `clocks`)" if long_step else "cgpu_bench_i8_peak (burst, idle device: the timed steps kept full clocks)")),
        entry("fx_combine_kernel", "fx_combine_kernel (exact int128 recombination of the S slice sums, one rounding, /r)",
              phases["fx_combine"], nS, "hbm", 4.0 * M_PAD * b_loc + 8.0 * K * b_loc, hbm_peak, "GB/s",
              "4 B x Mpad int32 partial sums read + 8 B x K doubles written per replicate"),
        entry("mt_jump_kernel", "mt_jump_kernel (x^J mod phi applied to a 624-word window; position list in smem)",
              phases["mt_jump"], 1, "hbm", 2.0 * 2496.0 * draws / 2 ** 22, hbm_peak, "GB/s",
              "2.5 KB read + 2.5 KB written per chunk state; the work is 624 x ~10^4 shared-memory word reads per jump",
              "shared-memory bandwidth bound; launches_per_step=1 stands for the ~10 levels of the jump tree"),
    ]
    # the draw kernel is bound by instruction issue, which the hbm|tensor vocabulary cannot express: add the
    # issue-slot view (warp-instructions per draw from the ncu capture profiles/r01_final_draw_equalfill_ncu_full.json;
    # the kernel has not changed since)
    try:
        prof = json.loads((ROOT / "profiles" / "r01_final_draw_equalfill_ncu_full.json").read_text())["launches"][0]
        wi_per_draw = prof["warp_instructions"] / 3.243e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 4.0 * torch.cuda.get_device_properties(local).multi_processor_count * sm_mhz * 1e6
        ach_i = wi_per_draw * draws / (max(phases["mt_draw"], 1e-9) * 1e-3)
        kernels[0]["issue_view"] = {"achieved": ach_i / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s",
                                    "frac": ach_i / issue_peak, "warp_instructions_per_draw": wi_per_draw,
                                    "source": "instruction count from the ncu capture, time and SM clock from this run"}
    except Exception:
        pass
    kernels = [e for e in kernels if e["share_of_step"] > 1e-6]  # e.g. fx_combine runs inside the GEMM epilogue when S divides 8
    # consel side: per-kernel CUDA-event times of one pipeline pass (cgpu_kernel_times)
    consel_alg = {  # kernel -> (bound, what the library's per-launch work figure counts)
        "col_top2_kernel": ("hbm", "8 B x K x B_s read per scale (one pass over the replicate block)"),
        "stat_build_kernel": ("hbm", "8 B x K x B_s read + 8 B x cm x B_s written per scale"),
        "rowsum_multi_kernel<1>": ("hbm", "8 B per (queried item, replicate) over all scales per one-threshold query"),
        "rowsum_multi/sorted_kernel": ("hbm", "8 B per (queried item, replicate) over all scales per pass of <= 8 thresholds"),
        "center_rows_kernel": ("hbm", "8 B read + 8 B written per (tree, replicate) of scale i0; the row mean is a sequential chain of B rounded adds"),
        "pair_weight_kernel": ("fp64", "pairs x B x (sub, mul, add) fp64 instructions"),
        "wsh_single_kernel": ("fp64", "members x (K-1) x B x (sub, mul, compare) fp64 instructions"),
        "order_stat_kernel": ("hbm", "8 B per element of the scale-i0 statistic rows (8 radix passes from L1/L2)"),
        "rowsum_kernel<kh>": ("hbm", "two replicate rows read per item"),
        "rowsum_kernel<sh>": ("hbm", "one replicate row read per member"),
        "rowsum_kernel<plain>": ("hbm", "one statistic row read per member"),
    }
    pipe_ms = (pipe or {}).get("makermt_consel_wall_ms") or 1e9
    for kname, (cnt, ms, work) in ktimes.items():
        if kname not in consel_alg or cnt == 0:
            continue
        bound, what = consel_alg[kname]
        alg = work / cnt
        if bound == "hbm":
            kernels.append(entry(kname, kname + " (consel side)", ms, cnt, "hbm", alg, hbm_peak, "GB/s", what,
                                 "share_of_step is the share of the makermt+consel pipeline wall time", base_ms=pipe_ms))
        else:
            e = entry(kname, kname + " (consel side)", ms, cnt, "tensor", alg, fp64_peak / 2.0, "T fp64-instr/s", what,
                      "FP64-pipe bound: peak = measured DFMA rate / 2 flops per instruction; `bound` says tensor only because "
                      "the contract's vocabulary is hbm|tensor", base_ms=pipe_ms)
            e["bound_detail"] = "fp64"
            kernels.append(e)
    step_kernels = [e for e in kernels if "(consel side)" not in e["kernel"]]
    dominant = max(step_kernels, key=lambda e: e["share_of_step"])
    roofline = dict(dominant)
    roofline["peak_source"] = peak_src + " (HBM); int8 and fp64 peaks measured live by cgpu_bench_i8_peak / cgpu_bench_fp64_peak"
    roofline["step_shares"] = {k_: v / total_ms for k_, v in phases.items() if k_ not in ("total", "rng")}
    roofline["int8_peak_tops_measured_live"] = {"burst": live["int8_tops"], "sustained_2s": live["int8_tops_sustained"], "used": i8_peak}
    for e_ in kernels:
        if e_["kernel"].startswith("gemm_i8"):
            e_["frac_of_burst_peak"] = e_["achieved"] / live["int8_tops"]
            e_["frac_of_sustained_peak"] = e_["achieved"] / live["int8_tops_sustained"]
    roofline["fp64_dfma_peak_tflops_measured_live"] = fp64_peak
    floor_ms = 2.0 * S_SLICES * K * N_SITES * b_loc * nS / (i8_peak * 1e12) * 1e3
    roofline["step"] = {"floor_ms": floor_ms, "ms_per_step": dev_ms / args.steps, "frac": floor_ms / (dev_ms / args.steps),
                        "what": "SURVEY 8(d): 2*S*K*n*sum(B) int8-ops per rank at the measured int8 peak, against the whole "
                                "device-resident step (draws + histograms + GEMMs)"}
    roofline["kernels"] = kernels

    # ---- CPU baseline: the reference's own scaleboot() on ONE host core, bounded sample (~10-20 s)
    cpu_baseline = cpu_baseline_leg(L, bn, K, N_SITES)

    # ---- side leg (N=1, default workload only): the per-GPU share of C4 in a process of its own
    side_c4 = None
    if world in (1, 8) and name == DEFAULT_CONFIG and os.environ.get("BENCH_SIDE", "1") != "0":
        try:
            mat.free()
            del flush, Lh, out_t, outs
            torch.cuda.empty_cache()
            # the library keeps freed blocks in the device's default stream-ordered pool (release threshold = max):
            # hand them back to the driver, the child process needs ~40 GB of its own
            import ctypes
            cu = ctypes.CDLL("libcuda.so.1")
            pool = ctypes.c_void_p()
            torch.cuda.synchronize()
            if cu.cuDeviceGetDefaultMemPool(ctypes.byref(pool), ctypes.c_int(local)) == 0:
                cu.cuMemPoolTrimTo(pool, ctypes.c_size_t(0))
        except Exception:
            pass
        side_c4 = side_c4_from(args, world)

    line = {
        "metric": METRIC, "value": acc_per_step_total * args.steps / (dev_ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic" if name != "C1" else "example/mam15.lnf",
        "config": config_json(name, c, world),
        "timing": "CUDA events on the library stream, first to last event of each step, max over ranks",
        "wall_ms_per_step": wall_ms / args.steps, "phases_ms_per_step": phases,
        "e2e": ({"error": e2e_error, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h} if e2e_error else
                {"value": acc_per_step_total * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                 "api": "cgpu_scaleboot_all(CGPU_RNG_MT19937, CGPU_SUM_INT8): pinned host L in, this rank's pinned host replicate blocks out"}),
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "spot_check": spot, "modes": modes, "pipeline": pipe,
    }
    if side_c4 is not None:
        line["side_c4"] = side_c4
    print(json.dumps(line))


# ------------------------------------------------------------------------ CPU baseline leg
CPU_BASELINE_ACC = 2.4e10  # tree-site accumulations of the sample: ~20 s of one core


def cpu_baseline_leg(L: np.ndarray, bn, K: int, n: int) -> dict:
    """single-threaded CPU baseline on a bounded sample of the same workload (same matrix, all ten scales): the
    UNMODIFIED reference's scaleboot() (src/makermt.c:35-75, compiled into oracle/_ref/libmakermt_ref.so) driven
    the way genrmt drives it (one seeded MT19937 stream, scale after scale) when that library was built
    (`kind: reference`), else the oracle's C restatement (`kind: port`).  Never raises."""
    try:
        from oracle import oracle as orc
        b_cpu = max(1, int(CPU_BASELINE_ACC / (K * n * len(RR))))
        so = orc.REF_DIR / "libmakermt_ref.so"
        if so.exists():
            global _REF_L
            _REF_L = np.ascontiguousarray(L, dtype=np.float64)
            try:
                t_cpu = 0.0
                for s_ in range(len(RR)):
                    dt, _ = _ref_worker((str(so), K, n, int(bn[s_]), b_cpu, 7), seed_it=(s_ == 0))
                    t_cpu += dt
            finally:
                _REF_L = None
            kind, what = "reference", "unmodified src/makermt.c:35-75 scaleboot(), -O2, oracle/_ref/libmakermt_ref.so"
        else:
            t0 = time.perf_counter()
            orc.makermt(L, bn, [b_cpu] * len(RR), 7)
            t_cpu = time.perf_counter() - t0
            kind, what = "port", "oracle/oracle.c"
        return {"value": K * n * b_cpu * len(RR) / t_cpu, "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"same matrix and scales, {b_cpu} replicates per scale ({t_cpu:.1f} s of {what})"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}", "unit": UNIT, "cores": 1, "kind": "port"}


# ------------------------------------------------------------------------ side leg: C4
def synth_matrix_threads(k: int, n: int, seed: int = 20261019) -> np.ndarray:
    """synth_matrix's distribution, filled by row blocks on all host threads (one PCG64 stream per block):
    10^9 entries in seconds instead of half a minute.  Side leg only -- the headline uses synth_matrix."""
    from concurrent.futures import ThreadPoolExecutor
    base = -2.3 - 4.07 * np.random.Generator(np.random.PCG64(seed)).exponential(size=n)
    L = np.empty((k, n))
    step = max(1, (1 << 25) // n)

    def fill(k0):
        k1 = min(k, k0 + step)
        blk = L[k0:k1]
        np.random.Generator(np.random.PCG64([seed, k0])).standard_normal(out=blk)  # releases the GIL
        blk *= 0.106
        blk += base[None, :]
        blk -= 0.5 * np.arange(k0, k1)[:, None] / n
    share = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))  # one process per GPU shares the host's cores
    with ThreadPoolExecutor(max_workers=max(1, min(32, (os.cpu_count() or 4) // share))) as ex:
        list(ex.map(fill, range(0, k, step)))
    return L


def run_side_c4(args) -> None:
    """`--side c4` (started by run_ours as its OWN process or torchrun job, after every leg of the default
    line): the genome-scale job C4 (K=100 x n=10^7, BASELINE.json configs[3]) -- all of it (10^4 replicates per
    scale) when started as 8 ranks, one GPU's share of it (1250 per scale) as a single process -- reference
    MT19937 stream + int8 sums, once with the default histogram of rows that long (32-bit global atomics) and
    once with the bucketed histogram (csrc/hist_bucket.cuh, cgpu_set_option("hist_bucket", 1)).  The makermt
    side has no data-path collective: every rank jumps the stream to its own replicate columns; the ranks only
    agree on the step time (max over ranks, gloo).  One JSON line per variant, flushed by rank 0 as soon as it
    exists: the bucketed kernels were written after the round's GPU budget was spent (pinned on the CPU by
    emulated execution only), so whatever they do on hardware must not cost the lines before them."""
    import torch

    import consel_b200 as cg

    t_start = time.perf_counter()
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    cg.init(local)
    cg.set_shard(rank, world)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def say(d: dict) -> None:
        if rank == 0:
            print(json.dumps(d), flush=True)

    c = workload("C4", world)
    if args.side_reps > 0:
        c["B_total"] = args.side_reps
    K = c["K"] if args.side_trees <= 0 else args.side_trees  # (the size switches exist for the CPU test of this leg)
    n = c["n"] if args.side_sites <= 0 else args.side_sites
    L = synth_matrix_threads(K, n)
    bn = (np.asarray(RR, dtype=np.float64) * np.float64(n)).astype(np.int32)  # bn = (int)(rr*n): src/makermt.c:165-168
    bb = np.full(len(RR), c["B_total"], dtype=np.int32)
    b0, b1 = cg.shard_range(int(bb[-1]))
    acc = float(K) * n * float(bb.sum())
    draws = float(np.sum(bn.astype(np.float64) * bb))
    t0 = time.perf_counter()
    mat = cg.Matrix(L)
    t_up = time.perf_counter() - t0
    say({"side": "C4", "variant": "setup", "n_gpus": world, "K": K, "n": n, "replicates_per_scale_total": int(c["B_total"]),
         "replicates_per_scale_per_gpu": b1 - b0, "draws_per_step_total": draws, "matrix_s": round(t0 - t_start, 2),
         "upload_s": round(t_up, 2)})
    first = None
    for label, opt in (("global_atomics (default)", 0), ("bucketed (hist_bucket=1)", 1)):
        cg.set_option("hist_bucket", opt)
        out = {"side": "C4", "variant": label, "n_gpus": world}
        late = max_over_ranks(time.perf_counter() - t_start) > float(os.environ.get("BENCH_SIDE_SECOND_VARIANT_BY_S", "80"))
        if opt and late:  # the leg is a side line: it must not stretch the bench run
            out["skipped"] = "the default variant took too long for a second one inside the leg's time budget"
            say(out)
            continue
        try:
            mat.scaleboot(bn, bb, 7, rng_mode=cg.RNG_MT19937, sum_mode=cg.SUM_INT8).free()  # warm-up (fixed-point slices, pools)
            ms, ph = [], {}
            cols = None
            for step in range(args.steps):
                max_over_ranks(0.0)  # the ranks start a step together
                reps = mat.scaleboot(bn, bb, 7, rng_mode=cg.RNG_MT19937, sum_mode=cg.SUM_INT8)
                tm = cg.last_timing()
                ms.append(max_over_ranks(tm["total"]))
                for k_, v in tm.items():
                    ph[k_] = ph.get(k_, 0.0) + v / args.steps
                if step == args.steps - 1:  # the first two and the last two replicate columns of this rank
                    cols = [reps.download_cols(s, 0, 2) for s in (0, len(RR) - 1)] + \
                           [reps.download_cols(len(RR) - 1, max(0, b1 - b0 - 2), min(2, b1 - b0))]
                reps.free()
            t = float(np.mean(ms))
            out.update({"ms_per_step": t, "value": acc / (t * 1e-3), "unit": UNIT, "steps": args.steps,
                        "timing": "CUDA events on the library stream (cgpu_last_timing), max over ranks",
                        "draws_per_s_per_gpu_through_draw_and_histogram": draws / world / ((ph["rng"] + ph["resample"]) * 1e-3),
                        "phases_ms_per_step_rank0": ph})
            if first is None:
                first = cols
                if rank == 0:
                    # the oracle's stream and site indices for replicate 0 of scale 0, exact sums by math.fsum
                    from oracle import oracle as orc
                    y = orc.mt_stream(7, 0, int(bn[0]))
                    cnt = np.bincount(orc.site_indices(y, n), minlength=n)[None, :]
                    kt = np.array([0, K - 1])
                    want = exact_sums(L[kt], cnt, float(bn[0]) / float(n))
                    out["spot_check"] = {"bit_equal_to_exact_rounded_sums": bool(np.array_equal(cols[0][kt, :1], want)),
                                         "what": "replicate 0 of scale 0, first and last tree: MT19937 draws and site indices by "
                                                 "the oracle, exact sums by math.fsum"}
            else:
                same = all(np.array_equal(a_, b_) for a_, b_ in zip(first, cols))
                out["identical_to_default_histogram"] = bool(max_over_ranks(0.0 if same else 1.0) == 0.0)
        except Exception as e:
            out["error"] = f"{type(e).__name__}: {e}"
        say(out)
    cg.set_option("hist_bucket", -1)
    if world > 1:
        dist.destroy_process_group()


def _free_port() -> int:
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return int(so.getsockname()[1])


def side_c4_from(parent_args, world: int = 1) -> dict:
    """runs `bench.py --side c4` as its own process (world 1) or its own torchrun job (world 8: every other rank
    of the parent job has exited by now) with a time limit and collects its lines; never raises"""
    limit = float(os.environ.get("BENCH_SIDE_TIMEOUT_S", "150"))
    me = str(Path(__file__).resolve())
    if world > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), me, "--side", "c4", "--steps", "2"]
    else:
        cmd = [sys.executable, me, "--side", "c4", "--steps", "2"]
    scrub = ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK",
             "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS")
    env = {k_: v for k_, v in os.environ.items() if k_ not in scrub and not k_.startswith("TORCHELASTIC_")}
    t0 = time.perf_counter()
    text, status = "", "ok"
    try:
        # its own process group: on a time-out the whole job (torchrun agent and ranks) is killed, not just its head
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=str(ROOT),
                             start_new_session=True)
        try:
            text, err = p.communicate(timeout=limit)
            if p.returncode != 0:
                tail = [t for t in err.strip().splitlines() if "Error" in t or "error" in t] or err.strip().splitlines() or [""]
                status = f"exit status {p.returncode}: {tail[-1].strip()[:300]}"
        except subprocess.TimeoutExpired:
            import signal
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except Exception:
                p.kill()
            text, _ = p.communicate()
            status = f"time limit of {limit:.0f} s reached: the variants printed so far are kept"
    except Exception as e:
        status = f"{type(e).__name__}: {e}"
    variants = []
    for ln in (text or "").splitlines():
        if ln.startswith("{"):
            try:
                variants.append(json.loads(ln))
            except ValueError:
                pass
    return {"what": "C4 (BASELINE.json configs[3]: K=100 x n=10^7, 10 x 10^4 replicates, an 8-GPU job), MT19937 stream + int8 sums, "
                    + ("the whole job as 8 ranks" if world == 8 else "one GPU's share of it (1250 replicates per scale)")
                    + ", in a process of its own after the default line was measured; `bucketed` is the first hardware run of "
                      "csrc/hist_bucket.cuh",
            "status": status, "seconds": round(time.perf_counter() - t0, 1), "variants": variants}


# ------------------------------------------------------------------------ reference
_REF_L = None  # the matrix of the in-process reference arm: set before the pool forks, inherited copy-on-write


def _ref_worker(args, seed_it: bool = True):
    """one host process of the in-process reference arm: scaleboot() of the unmodified makermt.c
    (seed_it False: the MT19937 stream continues where the previous call left it, as between the scales of genrmt)"""
    import ctypes as C
    so, K, n, bn, b_each, seed = args
    lib = C.CDLL(so)
    L = _REF_L
    rows = (C.POINTER(C.c_double) * K)(*[C.cast(L.ctypes.data + 8 * n * k, C.POINTER(C.c_double)) for k in range(K)])
    out = np.zeros((K, b_each))
    orow = (C.POINTER(C.c_double) * K)(*[C.cast(out.ctypes.data + 8 * b_each * k, C.POINTER(C.c_double)) for k in range(K)])
    lib.smrand.argtypes = [C.c_ulong]
    lib.scaleboot.restype = C.c_void_p
    if seed_it:
        lib.smrand(int(seed))
    t0 = time.perf_counter()
    lib.scaleboot(rows, orow, K, n, int(bn), int(b_each))
    dt = time.perf_counter() - t0
    return dt, out


def run_reference(args) -> None:
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import oracle as orc
    name = args.config
    c = workload(name, max(world, args.gpus))
    K, N_SITES = c["K"], c["n"]
    exe = orc.ref_bin("makermt")
    so = orc.REF_DIR / "libmakermt_ref.so"
    if exe is None or not so.exists():
        try:
            orc.build(ref=True)
        except Exception:
            pass
        exe = orc.ref_bin("makermt")
    L = make_matrix(name, c)
    cores = os.cpu_count() or 1
    cfg = config_json(name, c, max(world, args.gpus))
    _, bn, rr1 = orc.scale_setup(L, RR)
    use_cli = exe is not None and K * N_SITES <= 2_000_000
    # one job per host core (at least one per scale), job j on scale j % 10; a job is ~6 s of one core, so a
    # step is one round of ~6-9 s and the driver's --steps 20 run stays within a few minutes
    nproc = max(len(RR), cores)
    parts = -(-nproc // len(RR))
    jobs_sp = [(j % len(RR), j // len(RR)) for j in range(nproc)]
    b_each = max(2, int((6e9 if K * N_SITES > 2_000_000 else 9e9) / (K * N_SITES)))
    pipe_ref = None
    with tempfile.TemporaryDirectory(prefix="consel_ref_") as td:
        td = Path(td)
        if use_cli:
            kind = "reference"
            orc.write_mt(td / "in.mt", L)
            for s_, p in jobs_sp:
                orc.write_pa(td / f"p{s_}_{p}.pa", [RR[s_]], [b_each])

            def step():
                jobs = list(jobs_sp)
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
                return None
            sample = (f"{nproc} `makermt -s SEED -p one_scale.pa` processes, at most {cores} at a time (unmodified reference binary, -O2), "
                      f"{b_each} replicates each, same matrix read from a 17-digit .mt text file; wall clock includes process start and text parse")
        elif so.exists():
            kind = "reference"
            from multiprocessing import get_context
            global _REF_L
            _REF_L = np.ascontiguousarray(L, dtype=np.float64)  # read-only in the workers: the forked pages are shared
            ctx = get_context("fork")
            pool = ctx.Pool(min(cores, nproc))
            keep = {}

            def step():
                jobs = [(str(so), K, N_SITES, int(bn[s_]), b_each, 151 + s_ * parts + p) for s_, p in jobs_sp]
                res = pool.map(_ref_worker, jobs, chunksize=1)
                keep["reps"] = [res[s_][1] for s_ in range(len(RR))]  # part 0 of every scale: the consel sample
                return None
            sample = (f"{nproc} calls of scaleboot() (unmodified /root/reference/src/makermt.c:35-75 compiled -O2 into "
                      f"oracle/_ref/libmakermt_ref.so), one per (scale, part), at most {cores} host processes at a time, {b_each} replicates "
                      "each, matrix shared by the forked workers (no .mt text parse: the file would be "
                      f"{K * N_SITES * 24 / 1e9:.1f} GB per process)")
        else:
            kind = "port"
            nproc = 1
            parts = 1

            def step():
                orc.makermt(L, bn, [b_each] * len(RR), 7)
            sample = f"oracle port (oracle/oracle.c), 1 thread, {b_each} replicates per scale"
        total_b = b_each * (nproc if kind == "reference" else len(RR))
        for _ in range(min(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        t = time.perf_counter() - t0
        value = K * N_SITES * total_b * args.steps / t
        # ---- the reference's consel on the replicates the sample made (the metric's second half)
        cexe = orc.ref_bin("consel")
        try:
            if cexe is not None and kind == "reference":
                if use_cli:
                    subprocess.run([str(orc.ref_bin("catrep") or "catrep"), "--no_sort", "-m"] +
                                   [f"o{s}_0" for s in range(len(RR))] + ["all"], cwd=td, stdout=subprocess.DEVNULL, check=True)
                    bsamp = b_each
                else:
                    datvec, _, _ = orc.scale_setup(L, RR)
                    bsamp = keep["reps"][0].shape[1]
                    orc.write_rmt(td / "all.rmt", datvec, rr1, [bsamp] * len(RR), keep["reps"])
                t0 = time.perf_counter()
                subprocess.run([str(cexe), "all"], cwd=td, stdout=subprocess.DEVNULL, check=True, timeout=1500)
                t_consel = time.perf_counter() - t0
                t_makermt_full = K * N_SITES * float(c["B_total"]) * len(RR) / value
                pipe_ref = {"consel_wall_s_at_sample": t_consel, "replicates_per_scale_in_sample": int(bsamp),
                            "makermt_wall_s_full_size_on_all_cores": t_makermt_full,
                            "makermt_consel_wall_ms_extrapolated": (t_makermt_full + t_consel * c["B_total"] / bsamp) * 1e3,
                            "what": "unmodified reference consel (single-threaded, as shipped) on a .rmt of the sampled replicates; "
                                    "the full-size figure scales both parts linearly in the replicate count (SURVEY 8d)"}
        except Exception as e:
            pipe_ref = {"error": f"{type(e).__name__}: {e}"}
        if not use_cli and kind == "reference":
            pool.close()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": c["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic" if name != "C1" else "example/mam15.lnf", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(cores, nproc) if kind == "reference" else 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pipeline": pipe_ref,
    }))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--side", default="", choices=["", "c4"], help=argparse.SUPPRESS)  # run_ours starts this itself
    ap.add_argument("--side-reps", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--side-sites", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--side-trees", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.side == "c4":
        run_side_c4(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
