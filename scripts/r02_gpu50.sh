#!/bin/bash
# round 2, call 50: grid walk with bursts of 1 / 2 / 4 (the default) steps between two looks at the idle lanes (-DIPT_GRID_BURST)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
for v in b4 b1 b2 b4 b1; do cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so; timeout 300 $B > gpurun_out/r02_burst_$v.jsonl 2> /dev/null; echo "$v $(g gpurun_out/r02_burst_$v.jsonl)"; done
