mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu.log
for variant in 2 0 1; do for budget in 4 8 12 0; do
  BVH_B200_VARIANT=$variant BVH_B200_INNER_BUDGET=$budget timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_v${variant}_b${budget}.log 2>&1
  echo "variant=$variant budget=$budget rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/sweep_v${variant}_b${budget}.log | head -2 | tr '\n' ' ')"
done; done
for variant in 2 0; do
BVH_B200_VARIANT=$variant BVH_B200_INNER_BUDGET=8 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mesh grid > gpurun_out/sweep_grid_v${variant}.log 2>&1; echo "grid v$variant $(grep -o '"value": [0-9.]*' gpurun_out/sweep_grid_v${variant}.log | head -2 | tr '\n' ' ')"
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_pair -s 3 -c 1 -o gpurun_out/prof_pair -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pair.log 2>&1; echo "ncu_pair=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches2.log 2>&1; echo "ncu_launches=$?"
