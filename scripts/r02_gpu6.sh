#!/bin/bash
# round 2, call 6: full GPU suite with the round-2 tests, frame hashes, default bench line with per_config, reference arm
python scripts/update_frame_hashes.py > gpurun_out/r02_hashes.log 2>&1; tail -7 gpurun_out/r02_hashes.log
cp gpurun_out/frame_hashes.json tests/golden/frame_hashes.json
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_v2.log
tail -15 gpurun_out/r02_pytest_v2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default_v1.jsonl 2> gpurun_out/r02_bench_default_v1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_default_v1.err
cut -c1-300 gpurun_out/r02_bench_default_v1.jsonl
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_v1.jsonl 2> gpurun_out/r02_bench_reference_v1.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r02_bench_reference_v1.jsonl
