set -x
python - <<'PY' > gpurun_out/r02_peaks.log 2>&1
import ctypes as C, sys
sys.path.insert(0, ".")
import consel_b200 as cg
cg.init()
t = C.c_double()
for i in range(3):
    cg.api._check(cg.load().cgpu_bench_i8_peak(C.byref(t))); print("int8 peak TOP/s", t.value)
cg.api._check(cg.load().cgpu_bench_fp64_peak(C.byref(t))); print("fp64 peak TFLOP/s", t.value)
PY
cat gpurun_out/r02_peaks.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_a.log 2>&1; tail -5 gpurun_out/r02_gputests_a.log
I8_ONLY=1 python tools/quick_time.py C3 C5 > gpurun_out/r02_quick_a.log 2>&1; cat gpurun_out/r02_quick_a.log
CHOST_TIMING=1 PIPE_ITERS=2 python tools/pipeline_time.py C3 > gpurun_out/r02_pipe_c3_a.log 2>&1; tail -4 gpurun_out/r02_pipe_c3_a.log
CHOST_TIMING=1 PIPE_ITERS=2 timeout 600 python tools/pipeline_time.py C5s > gpurun_out/r02_pipe_c5s_a.log 2>&1; tail -4 gpurun_out/r02_pipe_c5s_a.log
