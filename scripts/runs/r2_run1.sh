# round 2, GPU run 1: first hardware execution of the SAH treelet pass (Quality Medium/High) + whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest.log)"
timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r2_memcheck.log | tail -1)"
timeout 600 compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/r2_racecheck.log | tail -1)"
for mesh in soup grid; do
  for q in low high; do
    timeout 400 python bench.py --mesh $mesh --quality $q --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_${mesh}_$q.log 2>&1
    echo "$mesh $q rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2_bench_${mesh}_$q.log | head -2 | tr '\n' ' ') $(grep -o '"inner_steps_per_ray": [0-9.]*' gpurun_out/r2_bench_${mesh}_$q.log) $(grep -o '"ms": [0-9.]*' gpurun_out/r2_bench_${mesh}_$q.log)"
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_high.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_full.log 2>&1
echo "full rc=$? $(tail -c 600 gpurun_out/r2_bench_full.log)"
