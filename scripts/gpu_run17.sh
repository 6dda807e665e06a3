mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe4.py > gpurun_out/probe4.log 2>&1; echo "probe4 rc=$?"; tail -9 gpurun_out/probe4.log
timeout 400 compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py > gpurun_out/racecheck17.log 2>&1; echo "racecheck rc=$?"; grep -i "RACECHECK SUMMARY\|hazard" gpurun_out/racecheck17.log | head -5
BVH_B200_HIERARCHY=rounds128 timeout 400 compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py > gpurun_out/racecheck17r.log 2>&1; echo "racecheck rounds rc=$?"; grep -i "RACECHECK SUMMARY\|hazard" gpurun_out/racecheck17r.log | head -5
timeout 400 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/memcheck17.log 2>&1; echo "memcheck rc=$?"; grep -i "ERROR SUMMARY" gpurun_out/memcheck17.log | head -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest17.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest17.log)"
