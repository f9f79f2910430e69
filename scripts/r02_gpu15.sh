#!/bin/bash
# round 2, call 15: where do the sporadic host-side milliseconds of a resident render go?  (e2e loop of small frames)
for i in 1 2 3 4; do
  IPT_VERBOSE=1 timeout 300 python bench.py --workload mirrors --steps 20 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r02_spike_$i.jsonl 2> gpurun_out/r02_spike_$i.err
  python - <<PY
import json,re
d=json.loads(open('gpurun_out/r02_spike_$i.jsonl').read())
print("run $i: device %.2f e2e %.2f call %.2f events %.2f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['rank0_ms']['render_call'], d['e2e']['rank0_kernel_ms_events']))
n=0
for l in open('gpurun_out/r02_spike_$i.err'):
    m=re.search(r'lists ([0-9.]+) ms, enqueue \+ wait ([0-9.]+) ms, kernels ([0-9.]+) ms', l)
    if m:
        a,b,c=map(float,m.groups()); n+=1
        if a+b-c>2: print("   render %d: lists %.2f enqueue+wait %.2f kernels %.2f" % (n,a,b,c))
PY
done
