#!/bin/bash
# round 2, call 30: axis lists unrolled by length (list-loop kernel): mirrors, maze, hashes, the list tests
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.3f  hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['frame_check']['matches_committed']))"; }
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-per-config"
for w in mirrors maze spheres; do timeout 300 $B --workload $w > gpurun_out/r02_unr_$w.jsonl 2>/dev/null; echo "unrolled $w $(g gpurun_out/r02_unr_$w.jsonl)"; done
IPT_NO_SHAPE=1 timeout 300 $B --workload spheres > gpurun_out/r02_unr_spheres_noshape.jsonl 2>/dev/null; echo "unrolled spheres through the list kernel $(g gpurun_out/r02_unr_spheres_noshape.jsonl)"
timeout 1200 python -m pytest tests -m gpu -q -x -k "box_room or oracle_per_pixel or committed or edge_cases or bounces_per_pass or degenerate or statistical or full_size" 2>&1 | tail -2
