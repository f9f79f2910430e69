#!/bin/bash
# round 2, call 9: Philox-7 + burst grid kernel + equal batches: hashes, parity suite, small benches, grid knob sweep
python scripts/update_frame_hashes.py > gpurun_out/r02_hashes2.log 2>&1; tail -7 gpurun_out/r02_hashes2.log
cp gpurun_out/frame_hashes.json tests/golden/frame_hashes.json
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_v3.log
tail -12 gpurun_out/r02_pytest_v3.log
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  e2e_ms %.2f  hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['frame_check']['matches_committed']))"; }
for w in spheres4k spheres mirrors maze; do timeout 300 $B --workload $w > gpurun_out/r02_p7_$w.jsonl 2>/dev/null; echo "philox7 $w $(g gpurun_out/r02_p7_$w.jsonl)"; done
S="python bench.py --workload synthetic1m --spp 16 --steps 2 --warmup 1 --no-cpu-baseline --no-per-config"
h() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f ms %.2f' % (d['gbounces_per_s'], d['ms_per_step']))"; }
timeout 300 $S > gpurun_out/r02_grid2_default.jsonl 2>/dev/null; echo "grid burst default $(h gpurun_out/r02_grid2_default.jsonl)"
for lm in 12 16 20 24; do for dm in 8 12 16 20; do IPT_LEAF_MIN=$lm IPT_DESCEND_MIN=$dm timeout 300 $S > gpurun_out/r02_grid2_lm${lm}_dm$dm.jsonl 2>/dev/null; echo "leaf_min=$lm descend_min=$dm $(h gpurun_out/r02_grid2_lm${lm}_dm$dm.jsonl)"; done; done
for rm in 4 12 16; do IPT_REFILL_MIN=$rm timeout 300 $S > gpurun_out/r02_grid2_rm$rm.jsonl 2>/dev/null; echo "refill_min=$rm $(h gpurun_out/r02_grid2_rm$rm.jsonl)"; done
