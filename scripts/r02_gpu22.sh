#!/bin/bash
# round 2, call 22: new default batch rule: default bench, config 5 at its own 256 spp, small scenes, parity subset
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f launches %d hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['frame_check']['matches_committed']))"; }
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
timeout 300 $B > gpurun_out/r02_d_4k.jsonl 2>/dev/null; echo "4k $(g gpurun_out/r02_d_4k.jsonl)"
timeout 600 $B --workload synthetic1m --steps 2 --warmup 1 > gpurun_out/r02_d_syn256.jsonl 2>/dev/null; echo "synthetic1m s256 $(g gpurun_out/r02_d_syn256.jsonl)"
timeout 600 $B --workload synthetic1m --spp 64 --steps 3 --warmup 1 > gpurun_out/r02_d_syn64.jsonl 2>/dev/null; echo "synthetic1m s64 $(g gpurun_out/r02_d_syn64.jsonl)"
timeout 600 $B --workload spheres4k_x3 --steps 2 --warmup 1 > gpurun_out/r02_d_x3.jsonl 2>/dev/null; echo "x3 s1024 $(g gpurun_out/r02_d_x3.jsonl)"
for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 > gpurun_out/r02_d_$w.jsonl 2>/dev/null; echo "$w $(g gpurun_out/r02_d_$w.jsonl)"; done
timeout 900 python -m pytest tests -m gpu -q -x -k "schedule or committed or tracer_program or dropin or properties" 2>&1 | tail -3
