#!/bin/bash
# round 2, call 1: baseline evidence for the shipped BVH split pipeline (config 5) before it is rebuilt
CMD="python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/r02_syn_plain.log 2> gpurun_out/r02_syn_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_syn_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend_bvh -s 4 -c 1 -f -o gpurun_out/r02_prof_extend_v7 $CMD > gpurun_out/r02_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 4 -c 1 -f -o gpurun_out/r02_prof_shade_v7 $CMD > gpurun_out/r02_ncu3.log 2>&1
cut -c1-400 gpurun_out/r02_syn_plain.log; tail -3 gpurun_out/r02_ncu2.log
