# round 2, GPU run 7: defaults refill_min 10 / budget 8, 64-byte triangles, stager state in shared memory; gather probe on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2g_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2g_pytest.log | cut -c1-150 | head -20
timeout 300 python scripts/gather_probe.py > gpurun_out/r2g_gather_probe.log 2>&1; echo "gather probe rc=$?"; cat gpurun_out/r2g_gather_probe.log | tail -8
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1) build $(grep -o '"ms": [0-9.]*' $2 | head -1)"; }
timeout 300 python bench.py --mesh soup $B > gpurun_out/r2g_soup.log 2>&1; line "soup default" gpurun_out/r2g_soup.log
timeout 300 python bench.py --mesh grid $B > gpurun_out/r2g_grid.log 2>&1; line "grid default" gpurun_out/r2g_grid.log
timeout 300 python bench.py --mesh soup --quality low $B > gpurun_out/r2g_soup_low.log 2>&1; line "soup low" gpurun_out/r2g_soup_low.log
timeout 300 python bench.py --mesh soup --kernel wide $B > gpurun_out/r2g_soup_wide.log 2>&1; line "soup wide" gpurun_out/r2g_soup_wide.log
BVH_B200_REFILL_MIN=1 timeout 300 python bench.py --mesh soup --kernel wide $B > gpurun_out/r2g_soup_wide_r1.log 2>&1; line "soup wide refill 1" gpurun_out/r2g_soup_wide_r1.log
timeout 300 python bench.py --mesh grid --kernel wide $B > gpurun_out/r2g_grid_wide.log 2>&1; line "grid wide" gpurun_out/r2g_grid_wide.log
timeout 400 python bench.py --config c3 $B > gpurun_out/r2g_c3.log 2>&1; line "c3" gpurun_out/r2g_c3.log
timeout 400 python bench.py --config c3 --kernel wide $B > gpurun_out/r2g_c3_wide.log 2>&1; line "c3 wide" gpurun_out/r2g_c3_wide.log
timeout 400 python bench.py --config c5 $B > gpurun_out/r2g_c5.log 2>&1; line "c5" gpurun_out/r2g_c5.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2g_full.log 2>&1; line "full default" gpurun_out/r2g_full.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2g_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2g_persistent python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2g_ncu_persistent.log 2>&1
echo "ncu persistent rc=$?"
