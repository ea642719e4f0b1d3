set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_c.log 2>&1; tail -15 gpurun_out/r02_gputests_c.log
I8_ONLY=1 python tools/quick_time.py C2 C3 C5 > gpurun_out/r02_quick_c.log 2>&1; cat gpurun_out/r02_quick_c.log
