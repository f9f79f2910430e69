#!/bin/bash
# round 2, call 5: first contact of k_extend_cw (one ray per lane over the 8-wide tree) - parity suite, config 5 A/B, knob sweep, ncu
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_cw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_cw.log
tail -4 gpurun_out/r02_pytest_cw.log
B="python bench.py --workload synthetic1m --spp 16 --steps 2 --warmup 1 --no-cpu-baseline --no-per-config"
g() { grep -o '"gbounces_per_s": [0-9.]*' $1 | head -1; }
IPT_BVH2=1 timeout 300 $B > gpurun_out/r02_cw_bvh2.jsonl 2> gpurun_out/r02_cw_bvh2.err; echo "bvh2 $(g gpurun_out/r02_cw_bvh2.jsonl)"
timeout 300 $B > gpurun_out/r02_cw_default.jsonl 2> gpurun_out/r02_cw_default.err; echo "cw default $(g gpurun_out/r02_cw_default.jsonl)"; tail -3 gpurun_out/r02_cw_default.err
for dm in 4 8 16 20; do IPT_DESCEND_MIN=$dm timeout 300 $B > gpurun_out/r02_cw_dm$dm.jsonl 2>/dev/null; echo "descend_min=$dm $(g gpurun_out/r02_cw_dm$dm.jsonl)"; done
for lm in 1 4 16 24; do IPT_LEAF_MIN=$lm timeout 300 $B > gpurun_out/r02_cw_lm$lm.jsonl 2>/dev/null; echo "leaf_min=$lm $(g gpurun_out/r02_cw_lm$lm.jsonl)"; done
for rm in 2 4 16; do IPT_REFILL_MIN=$rm timeout 300 $B > gpurun_out/r02_cw_rm$rm.jsonl 2>/dev/null; echo "refill_min=$rm $(g gpurun_out/r02_cw_rm$rm.jsonl)"; done
ncu --set full --clock-control none --import-source on -k regex:k_extend_cw -s 4 -c 1 -f -o gpurun_out/r02_prof_cw_v1 $B > gpurun_out/r02_ncu6.log 2>&1
tail -2 gpurun_out/r02_ncu6.log
