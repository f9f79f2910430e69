#!/bin/bash
# round 2, call 4: leaf-size sweep of the wide traversal on config 5 (smaller leaves = more box culling, fewer primitive tests)
B="python bench.py --workload synthetic1m --spp 16 --steps 2 --warmup 1 --no-cpu-baseline"
for l in 1 2; do for lf in 1 2; do for l2 in 1 2 4; do
  IPT_WIDE_LEAF=$lf IPT_WIDE_LPR=$l timeout 300 $B --leaf $l2 > gpurun_out/r02_sweep_lpr${l}_wl${lf}_l${l2}.jsonl 2>/dev/null
  echo "lpr=$l wide_leaf=$lf leaf2=$l2 $(cut -c1-120 gpurun_out/r02_sweep_lpr${l}_wl${lf}_l${l2}.jsonl | grep -o 'gbounces_per_s": [0-9.]*')"
done; done; done
for l2 in 1 2 8; do IPT_BVH2=1 timeout 300 $B --leaf $l2 > gpurun_out/r02_sweep_bvh2_l${l2}.jsonl 2>/dev/null; echo "bvh2 leaf2=$l2 $(grep -o 'gbounces_per_s": [0-9.]*' gpurun_out/r02_sweep_bvh2_l${l2}.jsonl)"; done
