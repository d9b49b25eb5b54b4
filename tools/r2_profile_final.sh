#!/bin/bash
# Final-commit refresh of the ncu evidence (run under gpurun, ONE GPU): launch list of the bench
# command, full captures of one C3 registration and of the steady state; raw CSV pages only.
set -x
M=smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-c4c5 > $O/r2_bench_under_ncu.log 2>&1
cap() {  # name, kernel regex, extra ncu args, driver mode
  ncu --set full --metrics $M --clock-control none -k "regex:$2" $3 -f -o /tmp/$1 python tools/profile_driver.py $4 > $O/r2_prof_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $O/r2_$1_raw.csv 2>> $O/r2_prof_$1.log
}
cap c3 "k_" "-c 70" c3
cap steady "k_match_grid_coop|k_rs_fused" "-s 24 -c 4" steady
cap batch "_batch" "-c 24" batch
du -sh $O
