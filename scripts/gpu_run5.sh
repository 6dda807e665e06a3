mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu.log
for budget in 4 8 16 0; do
  BVH_B200_VARIANT=2 BVH_B200_INNER_BUDGET=$budget timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_v2_b${budget}.log 2>&1
  echo "pair budget=$budget rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/sweep_v2_b${budget}.log | head -2 | tr '\n' ' ')"
done
BVH_B200_VARIANT=2 BVH_B200_INNER_BUDGET=8 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mesh grid > gpurun_out/sweep_grid_v2.log 2>&1; echo "grid pair $(grep -o '"value": [0-9.]*' gpurun_out/sweep_grid_v2.log | head -2 | tr '\n' ' ')"
for chunks in 1 2 4 8 16; do
  BVH_B200_VARIANT=1 BVH_B200_INNER_BUDGET=12 BVH_B200_E2E_CHUNKS=$chunks timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_c${chunks}.log 2>&1
  echo "e2e chunks=$chunks $(grep -o '"e2e": {"value": [0-9.]*, "unit": "Mrays/s", "ms_per_step": [0-9.]*' gpurun_out/e2e_c${chunks}.log)"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_pair -s 3 -c 1 -o gpurun_out/prof_pair2 -f env BVH_B200_VARIANT=2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pair2.log 2>&1; echo "ncu_pair=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hierarchy_kernel|rs_scatter_kernel|rs_scan_bins|rs_tile_hist" -s 16 -c 5 -o gpurun_out/prof_build2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_build2.log 2>&1; echo "ncu_build=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches3.log 2>&1; echo "ncu_launches=$?"
