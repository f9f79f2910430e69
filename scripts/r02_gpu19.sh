#!/bin/bash
# round 2, call 19: k_extend_grid occupancy and burst length A/B (rebuilt on the box)
S="python bench.py --workload synthetic1m --spp 16 --steps 3 --warmup 1 --no-cpu-baseline --no-per-config"
h() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f ms %.2f' % (d['gbounces_per_s'], d['ms_per_step']))"; }
timeout 300 $S > gpurun_out/r02_gridab_base.jsonl 2>/dev/null; echo "ctas4 burst4 $(h gpurun_out/r02_gridab_base.jsonl)"
for cfg in "5 4" "6 4" "4 2" "4 8" "4 16"; do set -- $cfg
  touch improved-path-tracer_b200/csrc/ipt_render.cu; make -C improved-path-tracer_b200 EXTRA_NVFLAGS="-DIPT_GRID_CTAS=$1 -DIPT_GRID_BURST=$2" > /dev/null 2>&1 || echo "build failed"
  timeout 300 $S > gpurun_out/r02_gridab_c$1_b$2.jsonl 2>/dev/null; echo "ctas$1 burst$2 $(h gpurun_out/r02_gridab_c$1_b$2.jsonl)"
done
