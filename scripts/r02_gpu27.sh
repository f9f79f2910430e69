#!/bin/bash
# round 2, call 27: sanity of the committed tree: GPU tests, smoke, the driver's own bench command
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_driver_like.jsonl 2> gpurun_out/r02_driver_like.err ) 2>&1 | grep real
python -c "
import json; d=json.loads(open('gpurun_out/r02_driver_like.jsonl').read()); print(d['value'], d['gbounces_per_s'], d['ms_per_step'], d['e2e']['value'], d['frame_check']['matches_committed'], len(d['per_config']), d['cpu_baseline']['value'])"
