mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench=$?"; tail -2 gpurun_out/bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --kernel simple --no-cpu-baseline > gpurun_out/bench_simple.log 2>&1; echo "bench_simple=$?"; tail -1 gpurun_out/bench_simple.log
timeout 300 python bench.py --steps 10 --warmup 3 --mesh grid --no-cpu-baseline > gpurun_out/bench_grid.log 2>&1; echo "bench_grid=$?"; tail -1 gpurun_out/bench_grid.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu_launches=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 3 -c 1 -o gpurun_out/prof_trace -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_trace.log 2>&1; echo "ncu_trace=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hierarchy_kernel|rs_scatter_kernel|rs_scan_kernel|morton_kernel" -s 16 -c 6 -o gpurun_out/prof_build -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_build.log 2>&1; echo "ncu_build=$?"
ls -la gpurun_out
