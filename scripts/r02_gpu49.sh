#!/bin/bash
# round 2, call 49: coplanar groups as their own instantiation (SHAPE = -1): the four typed-list workloads, a subset of the GPU tests
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  bounces %d  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['config']['traced_bounces_per_step'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_cand2_4k.jsonl 2>/dev/null; echo "spheres4k $(g gpurun_out/r02_cand2_4k.jsonl)"
for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_cand2_$w.jsonl 2>/dev/null; echo "$w $(g gpurun_out/r02_cand2_$w.jsonl)"; done
IPT_NO_GROUP=1 timeout 300 $B --workload maze --steps 20 --warmup 3 > gpurun_out/r02_cand2_nogroup_maze.jsonl 2>/dev/null; echo "maze without groups $(g gpurun_out/r02_cand2_nogroup_maze.jsonl)"
python -m pytest tests -m gpu -q -x -k "fp32_matches or every_kernel or committed or box_room or bounces_per_pass or russian or degenerate or schedule" 2>&1 | tail -2
