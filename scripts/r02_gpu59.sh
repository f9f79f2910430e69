#!/bin/bash
# round 2, call 59: the driver's own bench command on the last commit
time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_driver_cmd.jsonl 2> gpurun_out/r02_bench_driver_cmd.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_driver_cmd.jsonl').read()); print(d['value'], d['gbounces_per_s'], d['ms_per_step'], d['e2e']['value'], d['frame_check']['matches_committed'], d['roofline']['frac'], d['clocks'])
for p in d['per_config']: print(p['name'], p['dtype'], round(p['gbounces_per_s'],2), round(p['e2e_gbounces_per_s'],2), p['frame_check'])"
