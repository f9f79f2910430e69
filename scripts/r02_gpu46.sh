#!/bin/bash
# round 2, call 46: the ray's meta word kept apart across the bounces made in registers (A/B: build/ab/base.so = the commit before)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  bounces %d  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['config']['traced_bounces_per_step'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
for v in base new base new; do
  cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so
  timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_meta_${v}_4k.jsonl 2>/dev/null; echo "$v spheres4k $(g gpurun_out/r02_meta_${v}_4k.jsonl)"
  for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_meta_${v}_$w.jsonl 2>/dev/null; echo "$v $w $(g gpurun_out/r02_meta_${v}_$w.jsonl)"; done
done
