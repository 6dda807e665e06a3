mkdir -p gpurun_out
for i2f in 0 1; do for b in 3 4; do
  BVH_B200_WIDE_I2F=$i2f BVH_B200_USE_WIDE=1 BVH_B200_INNER_BUDGET=$b timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench20_$i2f_$b.log 2>&1
  echo "wide i2f=$i2f budget=$b rc=$? $(grep '^{' gpurun_out/bench20_$i2f_$b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'build', round(d['build']['ms'],3))")"
done; done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench20_default.log 2>&1
echo "default rc=$? $(grep '^{' gpurun_out/bench20_default.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'build', round(d['build']['ms'],3))")"
