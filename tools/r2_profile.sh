#!/bin/bash
# ncu captures of round 2 (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ (limit 64 MiB:
# the reports are converted to their raw CSV page on the box; only the steady-state report of the two
# hot kernels travels as .ncu-rep, with source, for the per-instruction view).
set -x
M=smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-c4c5 > $O/r2_bench_under_ncu.log 2>&1
cap() {  # name, kernel regex, extra ncu args, driver mode
  ncu --set full --metrics $M --clock-control none -k "regex:$2" $3 -f -o /tmp/$1 python tools/profile_driver.py $4 > $O/r2_prof_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/r2_$1_raw.csv 2>> $O/r2_prof_$1.log
}
cap c3 "k_" "-c 70" c3
cap brute "k_bf_nn" "-c 2" brute
cap bunny "k_range_keep|k_match_grid" "-c 4" bunny
cap batch "_batch" "-c 24" batch
ncu --set full --metrics $M --clock-control none --import-source on -k "regex:k_match_grid_coop|k_rs_fused" -s 24 -c 2 -f -o $O/r2_steady \
    python tools/profile_driver.py steady > $O/r2_prof_steady.log 2>&1
ncu -i $O/r2_steady.ncu-rep --page raw --csv > $O/r2_steady_raw.csv
du -sh $O
