# round 2, GPU run 5: duo mode for real (run 4 never selected it: the C API dropped the flag), finer refill threshold sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2e_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2e_pytest.log | cut -c1-150 | head -20
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1)"; }
for mesh in soup grid; do
  timeout 300 python bench.py --mesh $mesh --kernel duo $B > gpurun_out/r2e_${mesh}_duo.log 2>&1; line "$mesh duo" gpurun_out/r2e_${mesh}_duo.log
done
for b in 6 8 16; do
  BVH_B200_INNER_BUDGET=$b timeout 300 python bench.py --kernel duo $B > gpurun_out/r2e_soup_duo_b$b.log 2>&1; line "soup duo budget $b" gpurun_out/r2e_soup_duo_b$b.log
done
for m in 6 8 10 12; do
  BVH_B200_REFILL_MIN=$m timeout 300 python bench.py --kernel persistent $B > gpurun_out/r2e_soup_solo_r$m.log 2>&1; line "soup solo refill_min $m" gpurun_out/r2e_soup_solo_r$m.log
  BVH_B200_REFILL_MIN=$m timeout 300 python bench.py --kernel duo $B > gpurun_out/r2e_soup_duo_r$m.log 2>&1; line "soup duo refill_min $m" gpurun_out/r2e_soup_duo_r$m.log
done
for m in 8 12 20; do
  BVH_B200_REFILL_MIN=$m timeout 300 python bench.py --mesh grid --kernel persistent $B > gpurun_out/r2e_grid_solo_r$m.log 2>&1; line "grid solo refill_min $m" gpurun_out/r2e_grid_solo_r$m.log
done
for m in 8 16; do
  BVH_B200_REFILL_MIN=$m BVH_B200_INNER_BUDGET=8 timeout 300 python bench.py --kernel persistent $B > gpurun_out/r2e_soup_solo_r${m}_b8.log 2>&1; line "soup solo refill_min $m budget 8" gpurun_out/r2e_soup_solo_r${m}_b8.log
  BVH_B200_REFILL_MIN=$m BVH_B200_INNER_BUDGET=16 timeout 300 python bench.py --kernel persistent $B > gpurun_out/r2e_soup_solo_r${m}_b16.log 2>&1; line "soup solo refill_min $m budget 16" gpurun_out/r2e_soup_solo_r${m}_b16.log
done
BVH_B200_REFILL_MIN=8 timeout 400 python bench.py --config c3 --kernel persistent $B > gpurun_out/r2e_c3_r8.log 2>&1; line "c3 solo refill_min 8" gpurun_out/r2e_c3_r8.log
timeout 400 python bench.py --config c3 --kernel duo $B > gpurun_out/r2e_c3_duo.log 2>&1; line "c3 duo" gpurun_out/r2e_c3_duo.log
BVH_B200_REFILL_MIN=8 timeout 400 python bench.py --config c5 $B > gpurun_out/r2e_c5_r8.log 2>&1; line "c5 refill_min 8" gpurun_out/r2e_c5_r8.log
BVH_B200_REFILL_MIN=8 timeout 400 python bench.py --kernel wide $B > gpurun_out/r2e_soup_wide_r8.log 2>&1; line "soup wide (refill_min has no effect there yet)" gpurun_out/r2e_soup_wide_r8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2e_duo python bench.py --kernel duo --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2e_ncu_duo.log 2>&1
echo "ncu duo rc=$?"
BVH_B200_REFILL_MIN=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2e_solo_r8 python bench.py --kernel persistent --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2e_ncu_solo_r8.log 2>&1
echo "ncu solo r8 rc=$?"
