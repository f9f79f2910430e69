#!/bin/bash
# round 2, call 25: the three default-bench ncu outputs of final_check.sh with the launch indices of the 256 Mi batches
mkdir -p gpurun_out/final
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
ncu --metrics gpu__time_duration.sum --clock-control none -c 170 --csv --log-file gpurun_out/final/launches_default.csv $CMD > gpurun_out/final/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 37 --launch-count 1 -f -o gpurun_out/final/prof_deep $CMD > gpurun_out/final/ncu_deep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 36 --launch-count 1 -f -o gpurun_out/final/prof_pass0 $CMD > gpurun_out/final/ncu_pass0.log 2>&1
tail -n 2 gpurun_out/final/ncu_deep.log
