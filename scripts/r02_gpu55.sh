#!/bin/bash
# round 2, call 55 (2 GPUs): the two multi-GPU tests on the final kernels, and the default bench on 2 GPUs under torchrun
python -m pytest tests -m gpu -q -s -k "multi_gpu_gather or torchrun_two_process" 2>&1 | tail -6 | tee gpurun_out/r02_pytest_2gpus_final.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 6 --warmup 3 --no-per-config --no-cpu-baseline > gpurun_out/r02_final_n2.jsonl 2> gpurun_out/r02_final_n2.err; echo "n2 rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_final_n2.jsonl') if l.startswith('{')][-1]); print('N=2 %.0f Msamples/s %.1f Gb/s %.2f ms e2e %.0f hash %s committed %s rerender %s' % (d['value'], d['gbounces_per_s'], d['ms_per_step'], d['e2e']['value'], d['frame_sha256'][:12], d['frame_check']['matches_committed'], d['frame_check'].get('n1_rerender_identical')))"
