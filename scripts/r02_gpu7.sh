#!/bin/bash
# round 2, call 7: k_bounce_fast with claimed slices against static assignment; where the e2e gap comes from
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f  call %.2f  events %.2f  hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['rank0_ms']['render_call'], d['e2e']['rank0_kernel_ms_events'], d['frame_check']['matches_committed']))"; }
for w in spheres4k spheres mirrors maze; do
  IPT_STATIC_SLICES=1 timeout 300 $B --workload $w > gpurun_out/r02_claim_static_$w.jsonl 2>/dev/null; echo "static  $w $(g gpurun_out/r02_claim_static_$w.jsonl)"
  timeout 300 $B --workload $w > gpurun_out/r02_claim_dynamic_$w.jsonl 2>/dev/null; echo "dynamic $w $(g gpurun_out/r02_claim_dynamic_$w.jsonl)"
done
IPT_VERBOSE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r02_verbose_4k.jsonl 2> gpurun_out/r02_verbose_4k.err; grep "\[ipt\]" gpurun_out/r02_verbose_4k.err | tail -12
