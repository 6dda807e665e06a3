mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest14.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest14.log)"
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench14.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/bench14.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], {k:v for k,v in d['config'].items() if 'build' in k})"
timeout 300 python scripts/gpu_probe3.py > gpurun_out/probe3.log 2>&1
echo "probe3 rc=$?"; cat gpurun_out/probe3.log | tail -12
