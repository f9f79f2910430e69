#!/bin/bash
# round 2, call 43: MUFU.RSQ without rsqrtf()'s denormal guard, s24/u23 as one FMA (A/B: build/ab/base.so = the commit before)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
for v in base trim base trim; do
  cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so
  timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_trim_${v}_4k.jsonl 2>/dev/null; echo "$v spheres4k $(g gpurun_out/r02_trim_${v}_4k.jsonl)"
  for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_trim_${v}_$w.jsonl 2>/dev/null; echo "$v $w $(g gpurun_out/r02_trim_${v}_$w.jsonl)"; done
done
