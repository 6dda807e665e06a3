mkdir -p gpurun_out
export BVH_B200_WATCHDOG=4000000
timeout 600 python -m pytest tests -m gpu -q --timeout=200 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-220
for wide in 1 0; do for mesh in soup grid; do
  BVH_B200_USE_WIDE=$wide timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mesh $mesh > gpurun_out/bench_w${wide}_$mesh.log 2>&1
  echo "wide=$wide $mesh rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_w${wide}_$mesh.log | head -3 | tr '\n' ' ')"
done; done
for b in 4 8 24; do
  BVH_B200_USE_WIDE=1 BVH_B200_INNER_BUDGET=$b timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_w1_b$b.log 2>&1
  echo "wide budget=$b $(grep -o '"value": [0-9.]*' gpurun_out/bench_w1_b$b.log | head -1)"
done
BVH_B200_USE_WIDE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_wide -s 3 -c 1 -o gpurun_out/prof_wide -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_wide.log 2>&1; echo "ncu_wide=$?"
BVH_B200_USE_WIDE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_wide.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_wide.log 2>&1; echo "ncu_launches=$?"
