#!/bin/bash
# round 2, call 13 (2 GPUs): the whole parity suite including the two multi-GPU tests, measured values behind the fp32 thresholds,
# the default bench under torchrun on 2 GPUs (frame hash, N=1 re-render check, per_config)
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_2gpu.log
grep -E "MEASURED|passed|failed|skipped|rc=" gpurun_out/r02_pytest_2gpu.log | sort | uniq -c | sort -rn | head -40
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_v5.jsonl 2> gpurun_out/r02_bench_n2_v5.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_v5.jsonl').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['gbounces_per_s'], d['ms_per_step'], d['e2e']['value'], d['frame_check'])
for p in d['per_config']: print(p['name'], p['dtype'], round(p['gbounces_per_s'],2), round(p['e2e_gbounces_per_s'],2), p['frame_check'], p['ms_per_step_per_rank'])
PY
