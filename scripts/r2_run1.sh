# round 2, GPU run 1: first hardware execution of the SAH treelet pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
BVH_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k treelets > gpurun_out/r2_pytest_treelets.log 2>&1
echo "pytest treelets rc=$? $(tail -1 gpurun_out/r2_pytest_treelets.log)"
BVH_B200_SAH_TREELETS=1 timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/r2_memcheck_treelets.log 2>&1
echo "memcheck rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/r2_memcheck_treelets.log) $(grep 'ERROR SUMMARY' gpurun_out/r2_memcheck_treelets.log | tail -1)"
BVH_B200_SAH_TREELETS=1 timeout 900 compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py > gpurun_out/r2_racecheck_treelets.log 2>&1
echo "racecheck rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/r2_racecheck_treelets.log | tail -1)"
for mesh in soup grid; do
  for t in plain treelets; do
    flag=""; [ $t = treelets ] && flag="--sah-treelets"
    timeout 400 python bench.py --mesh $mesh $flag --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_${mesh}_$t.log 2>&1
    echo "$mesh $t rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2_bench_${mesh}_$t.log | head -2 | tr '\n' ' ') $(grep -o '"inner_steps_per_ray": [0-9.]*' gpurun_out/r2_bench_${mesh}_$t.log) $(grep -o '"ms": [0-9.]*' gpurun_out/r2_bench_${mesh}_$t.log)"
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_treelets.csv python bench.py --sah-treelets --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launches_treelets.log 2>&1
echo "ncu rc=$?"
