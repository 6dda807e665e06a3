mkdir -p gpurun_out
timeout 120 python scripts/gpu_sanitize.py > gpurun_out/small16.log 2>&1; echo "small rc=$?"; tail -5 gpurun_out/small16.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest16.log)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench16.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/bench16.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['build'])"
timeout 400 compute-sanitizer --tool racecheck --kernel-name regex:hierarchy_kernel python scripts/gpu_sanitize.py > gpurun_out/racecheck16.log 2>&1; echo "racecheck rc=$?"; grep -i "RACECHECK SUMMARY\|hazard" gpurun_out/racecheck16.log | head -5
timeout 400 compute-sanitizer --tool memcheck --kernel-name regex:hierarchy_kernel python scripts/gpu_sanitize.py > gpurun_out/memcheck16.log 2>&1; echo "memcheck rc=$?"; grep -i "ERROR SUMMARY" gpurun_out/memcheck16.log | head -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hierarchy_kernel -s 2 -c 1 -o gpurun_out/prof_hier5 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hier5.log 2>&1; echo "ncu_hier=$?"
