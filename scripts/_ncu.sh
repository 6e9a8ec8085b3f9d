set -x
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
# full captures: tile sweeps + fill + factor (DILU), one solve's kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tw_sweep|tw_fill|dilu_factor|relayout" -s 6 -c 6 -o gpurun_out/r02_ncu_dilu python scripts/ncu_target.py C3 auto dilu 3 update > gpurun_out/r02_ncu_dilu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tw_sweep|spmv_kernel|vec_" -s 30 -c 9 -o gpurun_out/r02_ncu_solve python scripts/ncu_target.py C3 auto dilu 1 solve > gpurun_out/r02_ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ilu0_factor|tw_sweep" -s 3 -c 3 -o gpurun_out/r02_ncu_ilu0 python scripts/ncu_target.py C3 auto ilu0 2 update > gpurun_out/r02_ncu_ilu0.log 2>&1
SCALE=1.0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_kernel|tw_sweep" -s 2 -c 3 -o gpurun_out/r02_ncu_c5 python scripts/ncu_target.py C5slab auto dilu 2 apply > gpurun_out/r02_ncu_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep
