#!/bin/bash
# round 2, call 11: ncu captures of the current k_bounce_fast (one 8-bounce pass and one pass 0 from the middle of the default frame), launch list
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config"
IPT_VERBOSE=1 timeout 300 python bench.py --workload spheres --steps 1 --warmup 1 --no-cpu-baseline --no-per-config 2>&1 >/dev/null | grep "typed lists" | head -1
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.2f  Ms/s %.0f  ms %.2f  hash_ok %s' % (d['gbounces_per_s'], d['value'], d['ms_per_step'], d['frame_check']['matches_committed']))"; }
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-per-config"
for w in spheres4k spheres; do
  IPT_NO_PAIR=1 timeout 300 $B --workload $w > gpurun_out/r02_pair2_off_$w.jsonl 2>/dev/null; echo "6 tests  $w $(g gpurun_out/r02_pair2_off_$w.jsonl)"
  timeout 300 $B --workload $w > gpurun_out/r02_pair2_on_$w.jsonl 2>/dev/null; echo "3 tests  $w $(g gpurun_out/r02_pair2_on_$w.jsonl)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r02_launches_default_v2.csv $CMD > gpurun_out/r02_ncu9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 136 --launch-count 1 -f -o gpurun_out/r02_prof_deep_v2 $CMD > gpurun_out/r02_ncu10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce_fast --launch-skip 135 --launch-count 1 -f -o gpurun_out/r02_prof_pass0_v2 $CMD > gpurun_out/r02_ncu11.log 2>&1
tail -n 2 gpurun_out/r02_ncu10.log; tail -n 2 gpurun_out/r02_ncu11.log
