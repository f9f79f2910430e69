#!/bin/bash
# round 2, call 29: 2-entry mailbox in the grid walk, A/B in one box (rebuilt without it)
S="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 1 --no-cpu-baseline --no-per-config"
h() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f ms %.2f hash_ok %s | %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_check']['matches_committed'], d['roofline_fp32']['flops_model'][23:]))"; }
timeout 300 $S > gpurun_out/r02_mb_on.jsonl 2>/dev/null; echo "mailbox    $(h gpurun_out/r02_mb_on.jsonl)"
touch improved-path-tracer_b200/csrc/ipt_render.cu; make -C improved-path-tracer_b200 EXTRA_NVFLAGS="-DIPT_GRID_MAILBOX=0" > /dev/null 2>&1 || echo "build failed"
timeout 300 $S > gpurun_out/r02_mb_off.jsonl 2>/dev/null; echo "no mailbox $(h gpurun_out/r02_mb_off.jsonl)"
for lm in 12 16; do for rm in 8 12; do IPT_LEAF_MIN=$lm IPT_REFILL_MIN=$rm timeout 300 $S > gpurun_out/r02_mb_off_$lm_$rm.jsonl 2>/dev/null; echo "no mailbox leaf_min $lm refill_min $rm $(h gpurun_out/r02_mb_off_$lm_$rm.jsonl)"; done; done
