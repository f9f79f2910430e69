#!/bin/bash
# round 2, call 20: batch size of the default bench (64 Mi default) - 32 / 128 / 256 Mi
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f launches %d hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['frame_check']['matches_committed']))"; }
for b in 33554432 67108864 134217728 268435456; do timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config --batch $b > gpurun_out/r02_batch_$b.jsonl 2>/dev/null; echo "batch $b $(g gpurun_out/r02_batch_$b.jsonl)"; done
for b in 67108864 268435456; do timeout 600 python bench.py --workload spheres4k_x3 --spp 256 --steps 2 --warmup 1 --no-cpu-baseline --no-per-config --batch $b > gpurun_out/r02_batch_x3_$b.jsonl 2>/dev/null; echo "x3 batch $b $(g gpurun_out/r02_batch_x3_$b.jsonl)"; done
