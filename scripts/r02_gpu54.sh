#!/bin/bash
# round 2, call 54: tree walk with its thresholds as compile-time constants (A/B: build/ab/base.so = the commit before), config 5 through the tree
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --workload synthetic1m --spp 32 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
for v in base treec base treec; do cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so; IPT_NO_GRID=1 timeout 300 $B > gpurun_out/r02_treec_$v.jsonl 2> /dev/null; echo "$v $(g gpurun_out/r02_treec_$v.jsonl)"; done
