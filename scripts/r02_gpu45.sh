#!/bin/bash
# round 2, call 45: (a) materials per hit code in shared memory (box rooms), (b) tree walk with the top of its stack in a register
# (A/B: build/ab/base.so = the commit before)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  bounces %d  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['config']['traced_bounces_per_step'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
for v in base new base new; do
  cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so
  timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_mc_${v}_4k.jsonl 2>/dev/null; echo "$v spheres4k $(g gpurun_out/r02_mc_${v}_4k.jsonl)"
  timeout 300 $B --workload spheres --steps 20 --warmup 3 > gpurun_out/r02_mc_${v}_spheres.jsonl 2>/dev/null; echo "$v spheres $(g gpurun_out/r02_mc_${v}_spheres.jsonl)"
  IPT_NO_GRID=1 timeout 300 $B --workload synthetic1m --spp 64 --steps 3 --warmup 2 > gpurun_out/r02_mc_${v}_tree.jsonl 2>/dev/null; echo "$v synthetic1m through the tree $(g gpurun_out/r02_mc_${v}_tree.jsonl)"
done
IPT_NO_GRID=1 timeout 600 python -m pytest tests -m gpu -q -x -k "bvh or tree or grid or synthetic or nearest" 2>&1 | tail -2
