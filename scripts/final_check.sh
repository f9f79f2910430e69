#!/bin/bash
# End-of-round check on a B200 box: GPU tests, smoke, the default bench, its ncu launch list and one full capture.
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench_default.jsonl 2> gpurun_out/final_bench_default.err; cut -c1-160 gpurun_out/final_bench_default.jsonl
for w in spheres mirrors maze; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null; done > gpurun_out/final_bench_small.jsonl; cut -c1-130 gpurun_out/final_bench_small.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 120 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 272 --launch-count 1 -f -o gpurun_out/prof_final_deep python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 270 --launch-count 1 -f -o gpurun_out/prof_final_pass0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_full0.log 2>&1
tail -1 gpurun_out/final_ncu_full.log | cut -c1-120
