#!/bin/bash
# End-of-round check on ONE B200 (round 2): GPU tests, smoke, the default bench (with per_config), the reference arm, the ncu launch
# lists and --set full captures the numbers in DESIGN.md / profiles/ come from, and tracer's benchmark matrix.
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log; tail -3 gpurun_out/final/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final/smoke.log
python bench.py > gpurun_out/final/bench_default.jsonl 2> gpurun_out/final/bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/final/bench_default.jsonl
python bench.py --impl reference > gpurun_out/final/bench_reference.jsonl 2> gpurun_out/final/bench_reference.err; cut -c1-200 gpurun_out/final/bench_reference.jsonl
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
ncu --metrics gpu__time_duration.sum --clock-control none -c 170 --csv --log-file gpurun_out/final/launches_default.csv $CMD > gpurun_out/final/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 37 --launch-count 1 -f -o gpurun_out/final/prof_deep $CMD > gpurun_out/final/ncu_deep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 36 --launch-count 1 -f -o gpurun_out/final/prof_pass0 $CMD > gpurun_out/final/ncu_pass0.log 2>&1
SYN="python bench.py --workload synthetic1m --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/final/launches_synthetic1m.csv $SYN > gpurun_out/final/ncu_launches_syn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend_grid -s 4 -c 1 -f -o gpurun_out/final/prof_grid $SYN > gpurun_out/final/ncu_grid.log 2>&1
for w in mirrors maze; do ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 2 --launch-count 1 -f -o gpurun_out/final/prof_$w python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/final/ncu_$w.log 2>&1; done
(cd improved-path-tracer_b200 && mkdir -p scenes && cp ../oracle/_ref/scenes/*.json scenes/ && rm -f benchmark.txt && timeout 900 python ../tools/trace_bench.py > ../gpurun_out/final/trace_bench.log 2>&1; cp benchmark.txt ../gpurun_out/final/benchmark_matrix.txt; rm -rf scenes benchmark.txt *.png)
tail -c 300 gpurun_out/final/benchmark_matrix.txt; echo
IPT_VERBOSE=1 improved-path-tracer_b200/tracer -d=10 -s=40 oracle/_ref/scenes/spheres.json 2>&1 | tr '\r' '\n' | grep -v "^Rendering" | tail -12 | tee gpurun_out/final/tracer_verbose.log; rm -f spheresD10S40.png benchmark.txt
