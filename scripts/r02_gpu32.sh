#!/bin/bash
# round 2, call 32: set_scene of the 1M-primitive scene after the parallel scene box: laps and end-to-end step; culling test
IPT_VERBOSE=1 timeout 300 python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 1 --no-cpu-baseline --no-per-config 2> gpurun_out/r02_setscene.err > gpurun_out/r02_setscene.jsonl
grep "set_scene" gpurun_out/r02_setscene.err | tail -7
python -c "
import json; d=json.loads(open('gpurun_out/r02_setscene.jsonl').read()); print(d['gbounces_per_s'], d['e2e']['gbounces_per_s'], d['e2e']['rank0_ms'], d['frame_check']['matches_committed'])"
timeout 600 python -m pytest tests -m gpu -q -x -k "culling or 4k_frame or committed or x3" 2>&1 | tail -2
