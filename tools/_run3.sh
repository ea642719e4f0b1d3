set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_b.log 2>&1; tail -8 gpurun_out/r02_gputests_b.log
CHOST_TIMING=1 PIPE_ITERS=3 python tools/pipeline_time.py C2 C3 > gpurun_out/r02_pipe_b.log 2>&1; grep -v "^   iter 0" gpurun_out/r02_pipe_b.log | tail -12
