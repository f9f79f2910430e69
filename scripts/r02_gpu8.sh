#!/bin/bash
# round 2, call 8: first contact of the uniform grid (k_extend_grid): parity suite, config 5 against the tree, density sweep, ncu
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_grid.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_grid.log
tail -12 gpurun_out/r02_pytest_grid.log
B="python bench.py --workload synthetic1m --spp 16 --steps 2 --warmup 1 --no-cpu-baseline --no-per-config"
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f ms %.2f e2e_ms %.2f set_scene %.1f | %s' % (d['gbounces_per_s'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['rank0_ms']['set_scene'], d['roofline_fp32']['flops_model']))"; }
IPT_NO_GRID=1 timeout 300 $B > gpurun_out/r02_grid_off.jsonl 2>/dev/null; echo "tree   $(g gpurun_out/r02_grid_off.jsonl)"
for dn in 0.1 0.2 0.35 0.6 1.0 2.0; do IPT_GRID_DENSITY=$dn timeout 300 $B > gpurun_out/r02_grid_d$dn.jsonl 2>/dev/null; echo "grid density=$dn $(g gpurun_out/r02_grid_d$dn.jsonl)"; done
for dm in 4 8 20; do IPT_DESCEND_MIN=$dm timeout 300 $B > gpurun_out/r02_grid_dm$dm.jsonl 2>/dev/null; echo "grid descend_min=$dm $(g gpurun_out/r02_grid_dm$dm.jsonl)"; done
for lm in 1 4 16; do IPT_LEAF_MIN=$lm timeout 300 $B > gpurun_out/r02_grid_lm$lm.jsonl 2>/dev/null; echo "grid leaf_min=$lm $(g gpurun_out/r02_grid_lm$lm.jsonl)"; done
ncu --set full --clock-control none --import-source on -k regex:k_extend_grid -s 4 -c 1 -f -o gpurun_out/r02_prof_grid_v1 $B > gpurun_out/r02_ncu7.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_grid_launches.csv $B > gpurun_out/r02_ncu8.log 2>&1
tail -2 gpurun_out/r02_ncu7.log
