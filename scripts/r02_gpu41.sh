#!/bin/bash
# round 2, call 41: grid walk modes: 1 = k_extend_grid, 2 = records in cell order, 3 = references a step ahead
# (record of a command that was run: the IPT_GRID_V1 / IPT_GRID_MODE switches and build/ab/*.so existed only in the A/B builds of that hour)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  e2e %.3f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['e2e']['gbounces_per_s'], d['frame_sha256'][:12]))"; }
IPT_GRID_MODE=3 timeout 600 python -m pytest tests -m gpu -q -x -k "grid or large_bvh or synthetic or public_abi" 2>&1 | tail -2
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
for m in 3 2 1 3; do IPT_GRID_MODE=$m timeout 300 $B > gpurun_out/r02_grid_m$m.jsonl 2> /dev/null; echo "mode $m $(g gpurun_out/r02_grid_m$m.jsonl)"; done
SYN="python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
IPT_GRID_MODE=3 ncu --set full --clock-control none --import-source on -k regex:k_extend_grid -s 4 -c 1 -f -o gpurun_out/prof_grid3 $SYN > gpurun_out/ncu_grid3.log 2>&1; tail -1 gpurun_out/ncu_grid3.log
