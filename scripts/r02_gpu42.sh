#!/bin/bash
# round 2, call 42: grid walk with references two steps ahead and the next record requested by a load nobody waits for (-DIPT_GRID_SLOT2)
# (record of a command that was run: the IPT_GRID_V1 / IPT_GRID_MODE switches and build/ab/*.so existed only in the A/B builds of that hour)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
for v in base slot2 base slot2; do cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so; timeout 300 $B > gpurun_out/r02_grid_$v.jsonl 2> /dev/null; echo "$v $(g gpurun_out/r02_grid_$v.jsonl)"; done
