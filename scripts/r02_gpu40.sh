#!/bin/bash
# round 2, call 40: k_extend_grid2 with a consumer-less load as the prefetch; grid density sweep with the inline records
# (record of a command that was run: the IPT_GRID_V1 / IPT_GRID_MODE switches and build/ab/*.so existed only in the A/B builds of that hour)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  e2e %.3f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['e2e']['gbounces_per_s'], d['frame_sha256'][:12]))"; }
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
timeout 300 $B > gpurun_out/r02_grid_v2b.jsonl 2> gpurun_out/r02_grid_v2b.err; echo "v2 ld-prefetch $(g gpurun_out/r02_grid_v2b.jsonl)"
for dn in 0.5 0.7 1.0; do IPT_GRID_DENSITY=$dn timeout 300 $B > gpurun_out/r02_grid_v2_d$dn.jsonl 2> /dev/null; echo "v2 density $dn $(g gpurun_out/r02_grid_v2_d$dn.jsonl)"; done
SYN="python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
ncu --set full --clock-control none --import-source on -k regex:k_extend_grid -s 4 -c 1 -f -o gpurun_out/prof_grid2b $SYN > gpurun_out/ncu_grid2b.log 2>&1; tail -1 gpurun_out/ncu_grid2b.log
