#!/bin/bash
# round 2, call 61: the list-loop captures once more at the final launch shape (128 threads per CTA)
mkdir -p gpurun_out/final
for w in mirrors maze; do timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 2 --launch-count 1 -f -o gpurun_out/final/prof_$w python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/final/ncu_$w.log 2>&1; tail -1 gpurun_out/final/ncu_$w.log; done
