set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"gemm_i8_2cta|hist_wide|mt_draw_kernel|mt_jump_kernel" --launch-skip 13 -c 7 -f -o gpurun_out/r02_c3_makermt python tools/prof_r02.py C3 makermt > gpurun_out/r02_ncu_c3.log 2>&1; tail -3 gpurun_out/r02_ncu_c3.log
$NCU -k regex:"gemm_i8_2cta|hist_wide" --launch-skip 4 -c 4 -f -o gpurun_out/r02_c5_makermt python tools/prof_r02.py C5 makermt > gpurun_out/r02_ncu_c5.log 2>&1; tail -3 gpurun_out/r02_ncu_c5.log
$NCU -k regex:"col_top2_kernel|stat_build|rowsum_multi|rowsum_sorted|rowsum_kernel|center_rows|pair_weight|wsh_single|order_stat" -c 24 -f -o gpurun_out/r02_c5_consel python tools/prof_r02.py C5 consel > gpurun_out/r02_ncu_c5c.log 2>&1; tail -3 gpurun_out/r02_ncu_c5c.log
ls -la gpurun_out/*.ncu-rep
# launch list of the C3 bench command (after it has run once without ncu in earlier calls)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_c3_launches.csv python bench.py --config C3 --steps 2 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r02_ncu_bench.log
