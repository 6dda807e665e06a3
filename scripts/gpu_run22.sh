mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather or ragged or tiny or nan" > gpurun_out/pytest22.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest22.log)"; grep -n "^FAILED\|^ERROR\|Error" gpurun_out/pytest22.log | head -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench22.log 2>&1
echo "bench rc=$? $(grep '^{' gpurun_out/bench22.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'build', round(d['build']['ms'],3))")"
