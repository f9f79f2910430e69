#!/bin/bash
# round 2, call 39: ncu --set full of one k_extend_grid2 launch (same launch index as profiles/r02_ncu_grid_final.txt)
# (record of a command that was run: the IPT_GRID_V1 / IPT_GRID_MODE switches and build/ab/*.so existed only in the A/B builds of that hour)
SYN="python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
ncu --set full --clock-control none --import-source on -k regex:k_extend_grid -s 4 -c 1 -f -o gpurun_out/prof_grid2 $SYN > gpurun_out/ncu_grid2.log 2>&1; tail -2 gpurun_out/ncu_grid2.log
