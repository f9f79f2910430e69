#!/bin/bash
# round 2, call 3: ncu capture of k_extend_wide (2 and 4 lanes per ray) on config 5
CMD="python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline"
IPT_WIDE_LPR=2 ncu --set full --clock-control none --import-source on -k regex:k_extend_wide -s 4 -c 1 -f -o gpurun_out/r02_prof_wide_v1 $CMD > gpurun_out/r02_ncu4.log 2>&1
IPT_WIDE_LPR=4 ncu --set full --clock-control none --import-source on -k regex:k_extend_wide -s 4 -c 1 -f -o gpurun_out/r02_prof_wide_v1_lpr4 $CMD > gpurun_out/r02_ncu5.log 2>&1
tail -3 gpurun_out/r02_ncu4.log
