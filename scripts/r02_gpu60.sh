#!/bin/bash
# round 2, call 60: the last commit (lanes without a ray read no record in k_bounce_fast): GPU tests, default workload
python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r02_pytest_last.log
python bench.py --no-cpu-baseline --no-per-config --steps 4 --warmup 2 > gpurun_out/r02_bench_last.jsonl 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/r02_bench_last.jsonl').read()); print('Gb/s %.3f ms %.3f hash %s %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12], d['frame_check']['matches_committed']))"
