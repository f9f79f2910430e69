#!/bin/bash
# round 2, call 51: grid walk without bursts: thresholds one at a time around refill 12 / cell step 12 / primitive step 16
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
for v in base r8 r16 c8 c16 p8 p24 base; do cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so; timeout 300 $B > gpurun_out/r02_thr_$v.jsonl 2> /dev/null; echo "$v $(g gpurun_out/r02_thr_$v.jsonl)"; done
