"""python tools/bench_table.py profiles/r02_bench_*.jsonl: markdown tables of bench.py lines (the per-config
summary and the per-kernel rooflines) for profiles/README.md (development aid)."""
import json
import sys

lines = []
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    d["_file"] = f.split("/")[-1]
    lines.append(d)

print("| file | config | GPUs | ms/step | value acc/s | e2e ms | e2e acc/s | makermt+consel ms | SM MHz (median) | step vs 8(d) floor | spot check |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for d in lines:
    if d.get("impl") == "reference":
        continue
    cfg = d["config"]["workload"].split(":")[0]
    p = d.get("pipeline") or {}
    r = d["roofline"]
    print(f"| `{d['_file']}` | {cfg} | {d['n_gpus']} | {d['ms_per_step']:.2f} | {d['value']:.3e} | {d['e2e']['ms_per_step']:.1f} | {d['e2e']['value']:.3e} | "
          f"{p.get('makermt_consel_wall_ms', float('nan')):.1f} | {(d.get('clocks') or {}).get('sm_mhz')} | {r['step']['frac']:.2f} | "
          f"{'bit-equal' if (d.get('spot_check') or {}).get('bit_equal_to_exact_rounded_sums') else 'FAILED'} |")
print()
for d in lines:
    if d.get("impl") != "reference":
        continue
    cfg = d["config"]["workload"].split(":")[0]
    p = d.get("pipeline") or {}
    print(f"reference arm `{d['_file']}` ({cfg}, {d['cpu_baseline']['cores']} host cores): {d['value']:.3e} acc/s; makermt+consel extrapolated to full size "
          f"{p.get('makermt_consel_wall_ms_extrapolated', float('nan'))/1e3:.0f} s (consel {p.get('consel_wall_s_at_sample', float('nan')):.1f} s at {p.get('replicates_per_scale_in_sample')} replicates per scale)")
print()
for d in lines:
    if d.get("impl") == "reference":
        continue
    cfg = d["config"]["workload"].split(":")[0]
    pk = d["roofline"]["int8_peak_tops_measured_live"]
    if not isinstance(pk, dict):  # lines written before the sustained measurement existed
        pk = {"burst": pk, "sustained_2s": float("nan"), "used": pk}
    print(f"**{cfg}, N={d['n_gpus']}** (`{d['_file']}`; int8 peak burst {pk['burst']:.0f} / sustained {pk['sustained_2s']:.0f} TOP/s, graded against {pk['used']:.0f}; "
          f"HBM {d['roofline']['kernels'][0]['peak']:.0f} GB/s)")
    print()
    print("| kernel | bound | launches | µs/launch | achieved | peak | frac | share |")
    print("|---|---|---:|---:|---:|---:|---:|---:|")
    for k in d["roofline"]["kernels"]:
        bound = k.get("bound_detail", k["bound"])
        print(f"| {k['kernel'].split(' (')[0][:60]}{' (consel)' if '(consel side)' in k['kernel'] else ''} | {bound} | {k['launches_per_step']} | {k['us_per_launch']:.1f} | "
              f"{k['achieved']:.1f} {k['unit'].split(' ')[0]} | {k['peak']:.1f} | {k['frac']:.2f} | {k['share_of_step']:.3f} |")
    print()
