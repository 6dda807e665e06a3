mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench=$?"; tail -4 gpurun_out/bench.log
timeout 300 python bench.py --steps 5 --warmup 3 --kernel simple --no-cpu-baseline > gpurun_out/bench_simple.log 2>&1; echo "bench_simple=$?"; tail -2 gpurun_out/bench_simple.log
timeout 300 python bench.py --steps 5 --warmup 3 --mesh grid --no-cpu-baseline > gpurun_out/bench_grid.log 2>&1; echo "bench_grid=$?"; tail -2 gpurun_out/bench_grid.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "tiny or ragged or boxes" > gpurun_out/memcheck.log 2>&1; echo "memcheck=$?"; tail -8 gpurun_out/memcheck.log
