#!/bin/bash
# round 2, call 28: grid: spheres filed under the cells they reach (A/B against box filing), big primitives through the primitive steps,
# density re-sweep; parity of the grid tests
S="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 1 --no-cpu-baseline --no-per-config"
h() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f ms %.2f hash_ok %s | %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_check']['matches_committed'], d['roofline_fp32']['flops_model'][23:]))"; }
timeout 300 $S > gpurun_out/r02_g2_default.jsonl 2>/dev/null; echo "sphere filing      $(h gpurun_out/r02_g2_default.jsonl)"
IPT_GRID_BOX_FILING=1 timeout 300 $S > gpurun_out/r02_g2_box.jsonl 2>/dev/null; echo "box filing         $(h gpurun_out/r02_g2_box.jsonl)"
for dn in 0.2 0.25 0.5; do IPT_GRID_DENSITY=$dn timeout 300 $S > gpurun_out/r02_g2_d$dn.jsonl 2>/dev/null; echo "density $dn $(h gpurun_out/r02_g2_d$dn.jsonl)"; done
for lm in 12 20; do IPT_LEAF_MIN=$lm timeout 300 $S > gpurun_out/r02_g2_lm$lm.jsonl 2>/dev/null; echo "leaf_min $lm $(h gpurun_out/r02_g2_lm$lm.jsonl)"; done
timeout 1200 python -m pytest tests -m gpu -q -x -k "grid or config5 or committed or large_bvh or wide" 2>&1 | tail -3
