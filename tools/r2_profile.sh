#!/bin/bash
# ncu captures of round 2 (run under gpurun, ONE GPU).  Outputs land in gpurun_out/.
set -x
M=smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-c4c5 > $O/r2_bench_under_ncu.log 2>&1
ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_ -c 70 -f -o $O/r2_c3 python tools/profile_driver.py c3 > $O/r2_prof_c3.log 2>&1
ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_match_grid_coop|k_rs_fused" -s 24 -c 4 -f -o $O/r2_steady python tools/profile_driver.py steady > $O/r2_prof_steady.log 2>&1
ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_bf_nn" -c 2 -f -o $O/r2_brute python tools/profile_driver.py brute > $O/r2_prof_brute.log 2>&1
ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_range_keep|k_match_grid" -c 4 -f -o $O/r2_bunny python tools/profile_driver.py bunny > $O/r2_prof_bunny.log 2>&1
ncu --set full --metrics $M --clock-control none --import-source on -k regex:"_batch" -c 24 -f -o $O/r2_batch python tools/profile_driver.py batch > $O/r2_prof_batch.log 2>&1
ls -la $O/*.ncu-rep
