mkdir -p gpurun_out
for b in 3 4 6; do
  BVH_B200_USE_WIDE=1 BVH_B200_INNER_BUDGET=$b timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench19_wide_$b.log 2>&1
  echo "wide budget=$b rc=$? $(grep '^{' gpurun_out/bench19_wide_$b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'checksum', d['config']['hits_checksum'])")"
done
BVH_B200_USE_WIDE=1 BVH_B200_INNER_BUDGET=4 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --mesh grid > gpurun_out/bench19_wide_grid.log 2>&1
echo "wide grid rc=$? $(grep '^{' gpurun_out/bench19_wide_grid.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3))")"
BVH_B200_USE_WIDE=1 BVH_B200_INNER_BUDGET=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_wide -s 3 -c 1 -o gpurun_out/prof_wide2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_wide2.log 2>&1; echo "ncu_wide=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or wide" > gpurun_out/pytest19.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest19.log)"
