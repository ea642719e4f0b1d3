set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_f.log 2>&1; tail -15 gpurun_out/r02_gputests_f.log
NCU="ncu --set full --clock-control none"
$NCU -k regex:"gemm_i8_2cta|hist_wide|mt_draw_kernel|mt_jump_kernel" --launch-skip 13 -c 7 -f -o /tmp/r02_c3_makermt python tools/prof_r02.py C3 makermt > gpurun_out/r02_ncu_c3.log 2>&1
python tools/ncu_summary.py /tmp/r02_c3_makermt.ncu-rep gpurun_out/r02_c3_makermt_ncu_full.json "C3 shape (K=1000, n=1e5, B=1e4, scales 0.5 and 1.4), MT19937 + int8: ncu --set full --clock-control none -k regex:gemm_i8_2cta|hist_wide|mt_draw_kernel|mt_jump_kernel --launch-skip 13 -c 7 python tools/prof_r02.py C3 makermt"
$NCU -k regex:"gemm_i8_2cta|hist_wide" --launch-skip 4 -c 4 -f -o /tmp/r02_c5_makermt python tools/prof_r02.py C5 makermt > gpurun_out/r02_ncu_c5.log 2>&1
python tools/ncu_summary.py /tmp/r02_c5_makermt.ncu-rep gpurun_out/r02_c5_makermt_ncu_full.json "C5 shape (K=1e4, n=5e4, B=1e4, scales 0.5 and 1.4): ncu --set full --clock-control none -k regex:gemm_i8_2cta|hist_wide --launch-skip 4 -c 4 python tools/prof_r02.py C5 makermt"
$NCU -k regex:"col_top2_kernel|stat_build|rowsum_multi|rowsum_sorted|rowsum_kernel|center_rows|pair_weight|wsh_single|order_stat" -c 24 -f -o /tmp/r02_c5_consel python tools/prof_r02.py C5 consel > gpurun_out/r02_ncu_c5c.log 2>&1
python tools/ncu_summary.py /tmp/r02_c5_consel.ncu-rep gpurun_out/r02_c5_consel_ncu_full.json "C5 shape consel side (K=1e4 items, B=1e4, scales 0.9 and 1.0): ncu --set full --clock-control none -k regex:col_top2_kernel|stat_build|rowsum_*|center_rows|pair_weight|wsh_single|order_stat -c 24 python tools/prof_r02.py C5 consel"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_c3_launches.csv python bench.py --config C3 --steps 2 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r02_ncu_bench.log
python - > gpurun_out/r02_c4_probe.log 2>&1 <<'PY'
import sys, time
import numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
rng = np.random.Generator(np.random.PCG64(1))
K, n, B = 100, 10_000_000, 64
base = -2.3 - 4.07 * rng.exponential(size=n)
L = np.empty((K, n))
for k in range(K):
    L[k] = base + 0.106 * rng.standard_normal(n) - 0.5 * k / n
cg.init()
RR = [0.5, 1.0, 1.4]
_, bn, _ = cg.scale_setup(L, RR)
m = cg.Matrix(L)
for rm in (0, 1):
    for it in range(2):
        t = time.perf_counter()
        r = m.scaleboot(bn, [B] * 3, 7, rng_mode=rm, sum_mode=2)
        print("rng", rm, "wall", time.perf_counter() - t, cg.last_timing(), flush=True)
        r.free()
PY
cat gpurun_out/r02_c4_probe.log
du -sh gpurun_out
