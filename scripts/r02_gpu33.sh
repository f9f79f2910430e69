#!/bin/bash
# round 2, call 33: warp instructions per bounce of the typed-list kernel per scene (ncu instruction count over 3 renders), rgb8 e2e leg
for w in spheres mirrors maze; do
  ncu --metrics smsp__inst_executed.sum --clock-control none -k regex:k_bounce_fast --csv --log-file gpurun_out/r02_inst_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r02_inst_$w.jsonl 2>/dev/null
  python - <<PY
import csv,json
rows=list(csv.reader(open('gpurun_out/r02_inst_$w.csv')))
hdr=None; tot=0; n=0
for r in rows:
    if r and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        try: tot+=float(r[hdr.index("Metric Value")].replace(',','')); n+=1
        except: pass
d=json.loads([l for l in open('gpurun_out/r02_inst_$w.jsonl') if l.startswith('{')][-1])
b=d['config']['traced_bounces_per_step']
print("$w: %d launches, %.3f G warp instructions over 3 renders, %.0f bounces per render -> %.2f warp instructions per bounce" % (n, tot/1e9, b, tot/3/b))
PY
done
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-per-config --e2e-rgb8 > gpurun_out/r02_e2e_rgb8.jsonl 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_e2e_rgb8.jsonl').read()); print(d['value'], d['e2e'])"
