python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in spheres4k spheres mirrors maze; do
  echo "== $w adaptive"; IPT_PASS_TIMES=1 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline 2>gpurun_out/pt_$w.err | cut -c1-140; tail -1 gpurun_out/pt_$w.err | cut -c1-300
done
echo "== x3"; python bench.py --workload spheres4k_x3 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | cut -c1-140
