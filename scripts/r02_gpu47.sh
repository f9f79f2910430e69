#!/bin/bash
# round 2, call 47: typed-list kernel variants (build/ab/*.so): A = the commit before; B = scenes with unknown materials go to the
# generic kernel (no 'teleport' selects in the loop) + bitwise zero tests; C = B + 'starts on a surface' asked before the scatter;
# D = B + Russian roulette as a template parameter; E = C + D; F = E + hit codes of the list loops by addition
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  bounces %d  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['config']['traced_bounces_per_step'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
for v in A B C D E F A F; do
  cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so
  timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_var_${v}_4k.jsonl 2>/dev/null; echo "$v spheres4k $(g gpurun_out/r02_var_${v}_4k.jsonl)"
  for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_var_${v}_$w.jsonl 2>/dev/null; echo "$v $w $(g gpurun_out/r02_var_${v}_$w.jsonl)"; done
done
