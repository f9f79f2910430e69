#!/bin/bash
# round 2, call 10: parity suite on the current build (wall pairs, grid defaults), wall-pair A/B, hashes must be unchanged
python scripts/update_frame_hashes.py > gpurun_out/r02_hashes3.log 2>&1
python - <<'PY'
import json
a=json.load(open('gpurun_out/frame_hashes.json')); b=json.load(open('tests/golden/frame_hashes.json'))
print("hashes unchanged by the wall-pair scan:", a==b)
PY
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_v4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_v4.log
tail -15 gpurun_out/r02_pytest_v4.log
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f  hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['frame_check']['matches_committed']))"; }
for w in spheres4k spheres spheres4k_x3; do
  IPT_NO_PAIR=1 timeout 300 $B --workload $w --spp 64 > gpurun_out/r02_pair_off_$w.jsonl 2>/dev/null; echo "6 tests  $w $(g gpurun_out/r02_pair_off_$w.jsonl)"
  timeout 300 $B --workload $w --spp 64 > gpurun_out/r02_pair_on_$w.jsonl 2>/dev/null; echo "3 tests  $w $(g gpurun_out/r02_pair_on_$w.jsonl)"
done
