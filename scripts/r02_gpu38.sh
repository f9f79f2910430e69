#!/bin/bash
# round 2, call 38: k_extend_grid2 (cell entries one step ahead, references' records inline) against k_extend_grid (IPT_GRID_V1=1)
# (record of a command that was run: the IPT_GRID_V1 / IPT_GRID_MODE switches and build/ab/*.so existed only in the A/B builds of that hour)
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  e2e %.3f  hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['e2e']['gbounces_per_s'], d['frame_sha256'][:12], d['frame_check'].get('matches_committed')))"; }
timeout 600 python -m pytest tests -m gpu -q -x -k "grid or large_bvh or synthetic or public_abi" 2>&1 | tail -3
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
IPT_GRID_V1=1 timeout 300 $B > gpurun_out/r02_grid_v1.jsonl 2> gpurun_out/r02_grid_v1.err; echo "v1 $(g gpurun_out/r02_grid_v1.jsonl)"
timeout 300 $B > gpurun_out/r02_grid_v2.jsonl 2> gpurun_out/r02_grid_v2.err; echo "v2 $(g gpurun_out/r02_grid_v2.jsonl)"
IPT_GRID_V1=1 timeout 300 $B > gpurun_out/r02_grid_v1b.jsonl 2> gpurun_out/r02_grid_v1b.err; echo "v1 $(g gpurun_out/r02_grid_v1b.jsonl)"
