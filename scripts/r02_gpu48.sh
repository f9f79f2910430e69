#!/bin/bash
# round 2, call 48: the candidate (variant D of call 47 + hit codes by addition + coplanar groups) with and without the groups
# (IPT_NO_GROUP=1), then the GPU tests
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.3f  bounces %d  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['config']['traced_bounces_per_step'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_cand_4k.jsonl 2>/dev/null; echo "spheres4k $(g gpurun_out/r02_cand_4k.jsonl)"
for w in spheres mirrors maze; do
  timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_cand_$w.jsonl 2>/dev/null; echo "$w $(g gpurun_out/r02_cand_$w.jsonl)"
  IPT_NO_GROUP=1 timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_cand_nogroup_$w.jsonl 2>/dev/null; echo "$w without groups $(g gpurun_out/r02_cand_nogroup_$w.jsonl)"
done
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
