#!/bin/bash
# round 2, call 58: the committed tree with 128 x 7 for the multi-bounce passes: GPU tests, smoke, default bench, launch list and the
# deep-pass capture (the other captures of scripts/final_check.sh are those of the commit before: their kernels did not change)
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log; tail -3 gpurun_out/final/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final/smoke.log
python bench.py > gpurun_out/final/bench_default.jsonl 2> gpurun_out/final/bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/final/bench_default.jsonl
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
ncu --metrics gpu__time_duration.sum --clock-control none -c 170 --csv --log-file gpurun_out/final/launches_default.csv $CMD > gpurun_out/final/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 37 --launch-count 1 -f -o gpurun_out/final/prof_deep $CMD > gpurun_out/final/ncu_deep.log 2>&1
