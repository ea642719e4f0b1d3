set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_e.log 2>&1; tail -15 gpurun_out/r02_gputests_e.log
I8_ONLY=1 python tools/quick_time.py C3 C5 > gpurun_out/r02_quick_e.log 2>&1; cat gpurun_out/r02_quick_e.log
timeout 900 python bench.py --config C5 --steps 3 --warmup 2 > gpurun_out/r02_bench_C5_b.jsonl 2> gpurun_out/r02_bench_C5_b.err; tail -c 400 gpurun_out/r02_bench_C5_b.err
