#!/bin/bash
# round 2, call 57: 128 x 7 for the multi-bounce passes (now the default), and for pass 0 as well (-DIPT_FIRST_THREADS=128 -DIPT_FIRST_CTAS=7)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
for v in deep128 both128 deep128 both128; do
  cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so
  timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_occ2_${v}_4k.jsonl 2>/dev/null; echo "$v spheres4k $(g gpurun_out/r02_occ2_${v}_4k.jsonl)"
  for w in spheres mirrors; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_occ2_${v}_$w.jsonl 2>/dev/null; echo "$v $w $(g gpurun_out/r02_occ2_${v}_$w.jsonl)"; done
done
