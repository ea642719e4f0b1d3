set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_d.log 2>&1; tail -15 gpurun_out/r02_gputests_d.log
timeout 900 python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/r02_bench_C3_b.jsonl 2> gpurun_out/r02_bench_C3_b.err; tail -c 400 gpurun_out/r02_bench_C3_b.err
timeout 900 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r02_bench_C2_b.jsonl 2> gpurun_out/r02_bench_C2_b.err; tail -c 400 gpurun_out/r02_bench_C2_b.err
