set -x
for c in C3 C1 C2 C5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 2 > gpurun_out/r02_bench_$c.jsonl 2> gpurun_out/r02_bench_$c.err
  echo "== $c rc=$?"; tail -c 600 gpurun_out/r02_bench_$c.err
done
