#!/bin/bash
# round 2, call 18: references of a grid cell sorted by primitive type (A/B), the new tests
S="python bench.py --workload synthetic1m --spp 16 --steps 3 --warmup 1 --no-cpu-baseline --no-per-config"
h() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f ms %.2f hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:10]))"; }
timeout 300 $S > gpurun_out/r02_bytype_off.jsonl 2>/dev/null; echo "by slot $(h gpurun_out/r02_bytype_off.jsonl)"
IPT_GRID_BY_TYPE=1 timeout 300 $S > gpurun_out/r02_bytype_on.jsonl 2>/dev/null; echo "by type $(h gpurun_out/r02_bytype_on.jsonl)"
timeout 300 $S > gpurun_out/r02_bytype_off2.jsonl 2>/dev/null; echo "by slot $(h gpurun_out/r02_bytype_off2.jsonl)"
IPT_GRID_BY_TYPE=1 timeout 300 $S > gpurun_out/r02_bytype_on2.jsonl 2>/dev/null; echo "by type $(h gpurun_out/r02_bytype_on2.jsonl)"
timeout 900 python -m pytest tests -m gpu -q -k "many_lights or progress or committed" 2>&1 | tail -3
