#!/bin/bash
# round 2, call 21: inside-box test against constants (A/B vs IPT_NO_PAIR), batch sizes, quick parity subset
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f launches %d hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['frame_check']['matches_committed']))"; }
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
IPT_NO_PAIR=1 timeout 300 $B > gpurun_out/r02_c_nopair_4k.jsonl 2>/dev/null; echo "no pair 4k $(g gpurun_out/r02_c_nopair_4k.jsonl)"
timeout 300 $B > gpurun_out/r02_c_pair_4k.jsonl 2>/dev/null; echo "pair    4k $(g gpurun_out/r02_c_pair_4k.jsonl)"
IPT_NO_PAIR=1 timeout 300 $B --workload spheres --steps 20 > gpurun_out/r02_c_nopair_s.jsonl 2>/dev/null; echo "no pair spheres $(g gpurun_out/r02_c_nopair_s.jsonl)"
timeout 300 $B --workload spheres --steps 20 > gpurun_out/r02_c_pair_s.jsonl 2>/dev/null; echo "pair    spheres $(g gpurun_out/r02_c_pair_s.jsonl)"
for b in 33554432 134217728 268435456; do timeout 300 $B --batch $b > gpurun_out/r02_batch_$b.jsonl 2>/dev/null; echo "batch $b $(g gpurun_out/r02_batch_$b.jsonl)"; done
timeout 900 python -m pytest tests -m gpu -q -x -k "box_room or oracle_per_pixel or schedule or bounces_per_pass or committed" 2>&1 | tail -3
