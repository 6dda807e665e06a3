mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe3.py > gpurun_out/probe3.log 2>&1
echo "probe3 rc=$?"; tail -8 gpurun_out/probe3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hierarchy_kernel -s 2 -c 1 -o gpurun_out/prof_hier4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hier4.log 2>&1; echo "ncu_hier=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest15.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest15.log)"
grep -n "^FAILED\|^ERROR" gpurun_out/pytest15.log | head
