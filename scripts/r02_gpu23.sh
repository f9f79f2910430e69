#!/bin/bash
# round 2, call 23: scatter normal selection + uniform unknown-material flag: benches, hashes, the tests that cover them
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['frame_check']['matches_committed']))"; }
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
timeout 300 $B > gpurun_out/r02_e_4k.jsonl 2>/dev/null; echo "4k $(g gpurun_out/r02_e_4k.jsonl)"
for w in spheres mirrors maze; do timeout 300 $B --workload $w --steps 20 > gpurun_out/r02_e_$w.jsonl 2>/dev/null; echo "$w $(g gpurun_out/r02_e_$w.jsonl)"; done
timeout 900 python -m pytest tests -m gpu -q -x -k "degenerate or box_room or oracle_per_pixel or committed or edge_cases or bounces_per_pass" 2>&1 | tail -3
