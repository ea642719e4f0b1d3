set -x
for ms in 2 10 50; do
BENCH_SAMPLER_MS=$ms timeout 900 python bench.py --config C3 --steps 3 --warmup 2 > gpurun_out/r02_bench_C3_s$ms.jsonl 2> gpurun_out/r02_bench_C3_s$ms.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_C3_s$ms.jsonl").read().strip().splitlines()[-1])
p=d["phases_ms_per_step"]; print("sampler $ms ms: total", p["total"], "sum", p["rng"]+p["resample"]+p["gemm_i8"], "clocks", d["clocks"])
PY
done
