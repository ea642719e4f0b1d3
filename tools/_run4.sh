set -x
CGPU_TRACE=1 python tools/trace_c3.py > gpurun_out/r02_trace_c3.log 2>&1
grep -c trace gpurun_out/r02_trace_c3.log
grep "iter" gpurun_out/r02_trace_c3.log
python tools/dist_check.py 2 > gpurun_out/r02_dist2.log 2>&1; tail -12 gpurun_out/r02_dist2.log
