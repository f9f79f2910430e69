#!/bin/bash
# round 2, call 56: more resident warps for the multi-bounce passes now that they need 71 registers: 128 x 7 and 224 x 4 (28 warps per SM)
# against 256 x 3 (24)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
for v in base t128c7 t224c4 base; do
  cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so
  timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_occ_${v}_4k.jsonl 2>/dev/null; echo "$v spheres4k $(g gpurun_out/r02_occ_${v}_4k.jsonl)"
  for w in mirrors maze; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_occ_${v}_$w.jsonl 2>/dev/null; echo "$v $w $(g gpurun_out/r02_occ_${v}_$w.jsonl)"; done
done
