set -x
python tools/exp_gemm2.py C3 C5 > gpurun_out/r02_exp_gemm2.log 2>&1; cat gpurun_out/r02_exp_gemm2.log
