set -x
nvidia-smi -L
I8_ONLY=1 python tools/quick_time.py C3 C5 > gpurun_out/r02_base_quick.log 2>&1
CHOST_TIMING=1 PIPE_ITERS=2 python tools/pipeline_time.py C3 > gpurun_out/r02_base_pipe_c3.log 2>&1
CHOST_TIMING=1 PIPE_ITERS=2 timeout 600 python tools/pipeline_time.py C5s > gpurun_out/r02_base_pipe_c5s.log 2>&1
PIPE_ITERS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_base_pipe_c3_launches.csv python tools/pipeline_time.py C3 > gpurun_out/r02_base_ncu.log 2>&1
tail -3 gpurun_out/r02_base_quick.log gpurun_out/r02_base_pipe_c3.log gpurun_out/r02_base_pipe_c5s.log
