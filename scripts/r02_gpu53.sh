#!/bin/bash
# round 2, call 53: grid walk: walls among the big primitives as pairs (A/B: build/ab/base.so = the commit before), grid tests
g() { python -c "import json,sys; d=json.loads(open('$1').read()); print('Gb/s %.3f  ms %.2f  hash %s' % (d['gbounces_per_s'], d['ms_per_step'], d['frame_sha256'][:12]))"; }
B="python bench.py --workload synthetic1m --spp 64 --steps 3 --warmup 2 --no-cpu-baseline --no-per-config"
for v in base pairs base pairs; do cp build/ab/$v.so improved-path-tracer_b200/libipt_b200.so; timeout 300 $B > gpurun_out/r02_bigpairs_$v.jsonl 2> /dev/null; echo "$v $(g gpurun_out/r02_bigpairs_$v.jsonl)"; done
IPT_VERBOSE=1 timeout 300 $B 2>&1 >/dev/null | grep "grid" | head -2
timeout 600 python -m pytest tests -m gpu -q -x -k "grid or config5 or public_abi or acceleration" 2>&1 | tail -2
