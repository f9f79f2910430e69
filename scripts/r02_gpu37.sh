#!/bin/bash
# round 2, call 37 (8 GPUs): BASELINE configs 5 (synthetic 1M, -d=10 -s=256) and 4 in its evenly loaded reading (x3, -d=32 -s=1024) at their
# own sample counts, N = 1, 2, 4, 8
for w in synthetic1m spheres4k_x3; do
  python bench.py --gpus 1 --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_full_${w}_n1.jsonl 2> gpurun_out/r02_full_${w}_n1.err; echo "$w n1 rc=$?"
  for n in 2 4 8; do
    if [ $w = spheres4k_x3 ] && [ $n != 8 ]; then continue; fi
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_full_${w}_n$n.jsonl 2> gpurun_out/r02_full_${w}_n$n.err; echo "$w n$n rc=$?"
  done
done
python - <<'PY'
import json
for w in ('synthetic1m','spheres4k_x3'):
    base=None
    for n in (1,2,4,8):
        try: d=json.loads([l for l in open('gpurun_out/r02_full_%s_n%d.jsonl'%(w,n)) if l.startswith('{')][-1])
        except Exception as e: print(w,n,'failed',e); continue
        if n==1: base=d
        print("%s N=%d  %.1f Msamples/s  %.2f Gbounces/s  %.2f ms  eff %.3f | e2e %.2f Gb/s eff %.3f | rerender %s  roofline %s %.3f" % (w, n, d['value'], d['gbounces_per_s'], d['ms_per_step'], d['value']/(n*base['value']), d['e2e']['gbounces_per_s'], d['e2e']['value']/(n*base['e2e']['value']), d['frame_check'].get('n1_rerender_identical'), d['roofline']['bound'], d['roofline']['frac']))
PY
