#!/bin/bash
# round 2, call 16 (8 GPUs): the default bench (with per_config) at N = 1, 2, 4, 8 back to back, as the driver runs it
python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r02_scale_n1.jsonl 2> gpurun_out/r02_scale_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/r02_scale_n$n.jsonl 2> gpurun_out/r02_scale_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try: d=json.loads(open('gpurun_out/r02_scale_n%d.jsonl'%n).read().strip().splitlines()[-1])
    except Exception as e: print(n,'failed',e); continue
    if n==1: base=d
    print("N=%d  %.0f Msamples/s  %.1f Gbounces/s  %.2f ms  eff %.3f | e2e %.0f eff %.3f | hash %s committed %s rerender %s" % (n, d['value'], d['gbounces_per_s'], d['ms_per_step'], d['value']/(n*base['value']), d['e2e']['value'], d['e2e']['value']/(n*base['e2e']['value']), d['frame_sha256'][:12], d['frame_check']['matches_committed'], d['frame_check'].get('n1_rerender_identical')))
    for p in d['per_config']: print("      %-13s %s  %8.2f Gb/s  e2e %8.2f Gb/s  %9.3f ms  roof %s %.3f  check %s" % (p['name'], p['dtype'], p['gbounces_per_s'], p['e2e_gbounces_per_s'], p['ms_per_step'], p['roofline']['bound'], p['roofline']['frac'], p['frame_check']))
PY
