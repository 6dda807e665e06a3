mkdir -p gpurun_out
for cfg in "1 8" "0 8" "1 12" "1 16" "1 6"; do
  set -- $cfg
  BVH_B200_E2E_TAPER=$1 BVH_B200_E2E_CHUNKS=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench18_$1_$2.log 2>&1
  echo "taper=$1 chunks=$2 rc=$? $(grep '^{' gpurun_out/bench18_$1_$2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'build_ms', round(d['build']['ms'],4))")"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest18.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest18.log)"
grep -n "^FAILED\|^ERROR\|Error" gpurun_out/pytest18.log | head
