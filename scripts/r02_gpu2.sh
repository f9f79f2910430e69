#!/bin/bash
# round 2, call 2: first contact of k_extend_wide - GPU parity suite, then A/B on config 5 (2-wide vs 8-wide at 1/2/4 lanes per ray)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_wide.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_wide.log
tail -5 gpurun_out/r02_pytest_wide.log
B="python bench.py --workload synthetic1m --spp 16 --steps 2 --warmup 1 --no-cpu-baseline"
IPT_BVH2=1 timeout 300 $B > gpurun_out/r02_ab_bvh2.jsonl 2> gpurun_out/r02_ab_bvh2.err
for l in 1 2 4; do IPT_VERBOSE=1 IPT_WIDE_LPR=$l timeout 300 $B > gpurun_out/r02_ab_wide_lpr$l.jsonl 2> gpurun_out/r02_ab_wide_lpr$l.err; echo "lpr $l rc=$?"; done
for l in 2 4; do for lf in 4 16; do IPT_WIDE_LEAF=$lf IPT_WIDE_LPR=$l timeout 300 $B > gpurun_out/r02_ab_wide_lpr${l}_leaf$lf.jsonl 2> gpurun_out/r02_ab_wide_lpr${l}_leaf$lf.err; done; done
for f in gpurun_out/r02_ab_*.jsonl; do echo $f; cut -c1-140 $f; done
