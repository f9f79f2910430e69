#!/bin/bash
# round 2, call 14: hashes incl. the fp64 one; claim-chunk A/B (rebuilt on the box); where set_scene spends its time on config 5
python scripts/update_frame_hashes.py > gpurun_out/r02_hashes5.log 2>&1; tail -8 gpurun_out/r02_hashes5.log | cut -c1-120
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-per-config"
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  ms %.3f  e2e_ms %.3f' % (d['gbounces_per_s'], d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run() { for w in spheres mirrors maze; do timeout 300 $B --workload $w > gpurun_out/r02_chunk_$1_$w.jsonl 2>/dev/null; echo "$1 $w $(g gpurun_out/r02_chunk_$1_$w.jsonl)"; done; timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config > gpurun_out/r02_chunk_$1_4k.jsonl 2>/dev/null; echo "$1 spheres4k $(g gpurun_out/r02_chunk_$1_4k.jsonl)"; }
run deep2_first8
IPT_STATIC_SLICES=1 run static
for cfg in "1 8" "4 8" "2 4" "2 16"; do set -- $cfg
  touch improved-path-tracer_b200/csrc/ipt_render.cu; make -C improved-path-tracer_b200 EXTRA_NVFLAGS="-DIPT_DEEP_CHUNK=${1}u -DIPT_FIRST_CHUNK=${2}u" > /dev/null 2>&1 || echo "build failed"
  run deep${1}_first${2}
done
touch improved-path-tracer_b200/csrc/ipt_render.cu; make -C improved-path-tracer_b200 > /dev/null 2>&1
IPT_VERBOSE=1 timeout 300 python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-per-config 2>&1 >/dev/null | grep "set_scene\|\[grid\]\|\[ipt\] grid" | tail -8
