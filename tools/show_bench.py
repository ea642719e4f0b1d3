"""python tools/show_bench.py file.jsonl ...: condensed view of bench.py lines (development aid)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f"== {f}: {d['config']['workload'][:60]}")
    print(f"   value {d['value']:.3e} {d['unit']}  ms/step {d['ms_per_step']:.2f}  n_gpus {d['n_gpus']} scaling {d['scaling']}  launches {d.get('gpu_launches')}")
    if d.get("impl") == "reference":
        print("   ", d["cpu_baseline"]["cores"], "cores", d.get("pipeline")); continue
    e = d["e2e"]; print(f"   e2e {e['value']:.3e} ms {e['ms_per_step']:.2f} h2d {e['h2d_bytes_per_step']/1e6:.0f}MB d2h {e['d2h_bytes_per_step']/1e6:.0f}MB")
    print("   clocks", d["clocks"], "spot", {k: v for k, v in (d.get("spot_check") or {}).items() if k != "what"})
    print("   phases", {k: round(v, 2) for k, v in d["phases_ms_per_step"].items() if v})
    r = d["roofline"]; print(f"   roofline: {r['kernel'][:40]} frac {r['frac']:.3f}; step floor {r['step']['floor_ms']:.2f} ms frac {r['step']['frac']:.3f}; i8 peak {r['int8_peak_tops_measured_live']}")
    for k in r["kernels"]:
        print(f"      {k['kernel'][:48]:48s} {k['bound']:6s} {k['achieved']:9.1f}/{k['peak']:7.1f} {k['unit'][:10]:10s} frac {k['frac']:.3f} n={k['launches_per_step']:4d} {k['us_per_launch']:9.1f}us share {k['share_of_step']:.3f}")
    print("   modes", {k: round(v["ms_per_step"], 2) for k, v in d["modes"].items()})
    print("   pipeline", {k: v for k, v in (d.get("pipeline") or {}).items() if k != "what"})
    print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:60])
