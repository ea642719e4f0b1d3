"""Condense an .ncu-rep (ncu --set full) into the JSON summary kept under profiles/.

usage: python tools/ncu_summary.py capture.ncu-rep out.json [note]
Reads the raw page through `ncu -i ... --page raw --csv`, keeps per launch: duration,
grid, registers, waves, DRAM bytes, issue/pipe utilisation and stall reasons per issue.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__shared_mem_per_block_static": "smem_static",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "pipe_tensor_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_cycles_pct",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_imma_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__cycles_elapsed.avg.per_second": "sm_clock",
}


def to_base(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0,
             "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0,
             "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0, "GHz": 1e9, "MHz": 1e6}
    return x * scale.get(unit, 1.0)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    launches = []
    for r in body:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in KEEP and r[i] not in ("", "n/a"):
                try:
                    d[KEEP[h]] = to_base(r[i], units[i])
                except ValueError:
                    d[KEEP[h]] = r[i]
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        d["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        launches.append(d)
    json.dump({"source": rep.split("/")[-1], "note": note, "units": "seconds, bytes, Hz, percent", "launches": launches},
              open(out, "w"), indent=1)
    for d in launches:
        print(d["kernel"][:60], f"{d.get('duration', 0)*1e6:.1f} us", f"dram {d.get('dram_bytes', 0)/1e6:.1f} MB",
              f"issue {d.get('issue_active_pct', 0):.1f}%")


if __name__ == "__main__":
    main()
