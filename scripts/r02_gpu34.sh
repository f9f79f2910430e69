#!/bin/bash
# round 2, call 34: refracted direction behind a warp vote (A/B, rebuilt on the box with -DIPT_LOBE_VOTE)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  ms %.3f  hash_ok %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_check']['matches_committed']))"; }
B="python bench.py --no-cpu-baseline --no-per-config"
run() { timeout 300 $B --steps 4 --warmup 2 > gpurun_out/r02_vote_$1_4k.jsonl 2>/dev/null; echo "$1 spheres4k $(g gpurun_out/r02_vote_$1_4k.jsonl)"; for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 --warmup 3 > gpurun_out/r02_vote_$1_$w.jsonl 2>/dev/null; echo "$1 $w $(g gpurun_out/r02_vote_$1_$w.jsonl)"; done; }
run always
touch improved-path-tracer_b200/csrc/ipt_render.cu; make -C improved-path-tracer_b200 EXTRA_NVFLAGS="-DIPT_LOBE_VOTE" > /dev/null 2>&1 || echo "build failed"
run vote
