# round 2, GPU run 14 (short): the new switch-invariance test; a few one-line experiments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "traversal_switches or gather_entry_point or full_size" > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2r_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2r_pytest.log | cut -c1-200 | head
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) build $(grep -o '"ms": [0-9.]*' $2 | head -1)"; }
BVH_B200_TREELET_BLOCKS=4 timeout 300 python bench.py $B > gpurun_out/r2r_soup_tb4.log 2>&1; line "soup treelet blocks 4" gpurun_out/r2r_soup_tb4.log
BVH_B200_REFILL_MIN=6 timeout 300 python bench.py --mesh grid $B > gpurun_out/r2r_grid_r6.log 2>&1; line "grid refill 6" gpurun_out/r2r_grid_r6.log
BVH_B200_REFILL_MIN=12 timeout 300 python bench.py --mesh grid $B > gpurun_out/r2r_grid_r12.log 2>&1; line "grid refill 12" gpurun_out/r2r_grid_r12.log
BVH_B200_REFILL_MIN=20 timeout 300 python bench.py --mesh grid $B > gpurun_out/r2r_grid_r20.log 2>&1; line "grid refill 20" gpurun_out/r2r_grid_r20.log
BVH_B200_REFILL_MIN=6 timeout 300 python bench.py --config c3 $B > gpurun_out/r2r_c3_r6.log 2>&1; line "c3 refill 6" gpurun_out/r2r_c3_r6.log
BVH_B200_REFILL_MIN=12 timeout 300 python bench.py --config c3 $B > gpurun_out/r2r_c3_r12.log 2>&1; line "c3 refill 12" gpurun_out/r2r_c3_r12.log
