#!/bin/bash
# round 2, call 31: the committed tree once more: GPU tests, smoke, default bench (driver's command)
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log; tail -3 gpurun_out/final/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_default.jsonl 2> gpurun_out/final/bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/final/bench_default.jsonl').read()); print(d['value'], d['gbounces_per_s'], d['ms_per_step'], d['e2e']['value'], d['frame_check']['matches_committed'])
for p in d['per_config']: print(p['name'], p['dtype'], round(p['gbounces_per_s'],2), round(p['e2e_gbounces_per_s'],2), p['frame_check'])"
