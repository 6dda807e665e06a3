mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench=$?"; tail -1 gpurun_out/bench.log | cut -c1-250
timeout 200 python bench.py --steps 20 --warmup 3 --mesh grid --no-cpu-baseline > gpurun_out/bench_grid.log 2>&1; echo "bench_grid=$?"; tail -1 gpurun_out/bench_grid.log | cut -c1-250
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench_ref=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-400
for chunks in 2 4 8 16; do
  BVH_B200_E2E_CHUNKS=$chunks timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_c${chunks}.log 2>&1
  echo "e2e chunks=$chunks $(grep -o '"e2e": {"value": [0-9.]*, "unit": "Mrays/s", "ms_per_step": [0-9.]*' gpurun_out/e2e_c${chunks}.log)"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 3 -c 1 -o gpurun_out/prof_trace3 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_trace3.log 2>&1; echo "ncu_trace=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hierarchy_kernel|rs_scatter_kernel|rs_scan_bins|rs_tile_hist|morton_kernel|centre_bounds" -s 15 -c 6 -o gpurun_out/prof_build3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_build3.log 2>&1; echo "ncu_build=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches3.log 2>&1; echo "ncu_launches=$?"
