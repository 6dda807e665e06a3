mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu.log
python scripts/pcie_bw.py > gpurun_out/pcie.log 2>&1; cat gpurun_out/pcie.log
for tma in 1 0; do for budget in 0 4 8 16; do
  BVH_B200_TMA=$tma BVH_B200_INNER_BUDGET=$budget timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_t${tma}_b${budget}.log 2>&1
  echo "tma=$tma budget=$budget rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/sweep_t${tma}_b${budget}.log | head -1)"
done; done
BVH_B200_TMA=1 BVH_B200_INNER_BUDGET=8 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mesh grid > gpurun_out/sweep_grid_b8.log 2>&1; echo "grid b8 $(grep -o '"value": [0-9.]*' gpurun_out/sweep_grid_b8.log | head -1)"
BVH_B200_TMA=1 BVH_B200_INNER_BUDGET=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mesh grid > gpurun_out/sweep_grid_b0.log 2>&1; echo "grid b0 $(grep -o '"value": [0-9.]*' gpurun_out/sweep_grid_b0.log | head -1)"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 3 -c 1 -o gpurun_out/prof_trace2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_trace2.log 2>&1; echo "ncu_trace=$?"
