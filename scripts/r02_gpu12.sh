#!/bin/bash
# round 2, call 12: micro-tile-first sample ids + chunked claims: parity suite, hashes unchanged?, benches, launch list
python scripts/update_frame_hashes.py > gpurun_out/r02_hashes4.log 2>&1
python - <<'PY'
import json
a=json.load(open('gpurun_out/frame_hashes.json')); b=json.load(open('tests/golden/frame_hashes.json'))
print("hashes unchanged by the new sample order:", a==b)
PY
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_v5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_v5.log
tail -6 gpurun_out/r02_pytest_v5.log
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f launches %d hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['frame_check']['matches_committed']))"; }
for w in spheres4k spheres mirrors maze spheres4k_x3; do timeout 300 $B --workload $w > gpurun_out/r02_v5_$w.jsonl 2>/dev/null; echo "v5 $w $(g gpurun_out/r02_v5_$w.jsonl)"; done
IPT_STATIC_SLICES=1 timeout 300 $B > gpurun_out/r02_v5_static_spheres4k.jsonl 2>/dev/null; echo "v5 static spheres4k $(g gpurun_out/r02_v5_static_spheres4k.jsonl)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_default_v3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu12.log 2>&1
