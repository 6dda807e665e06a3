# round 2, GPU run 4: duo mode of the persistent kernel (lane pairs share node fetches), refill threshold, budgets
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2d_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2d_pytest.log | cut -c1-150 | head -20
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1)"; }
for mesh in soup grid; do
  for k in persistent duo; do
    timeout 300 python bench.py --mesh $mesh --kernel $k $B > gpurun_out/r2d_${mesh}_$k.log 2>&1; line "$mesh $k" gpurun_out/r2d_${mesh}_$k.log
  done
done
for b in 6 8 16 24; do
  BVH_B200_INNER_BUDGET=$b timeout 300 python bench.py --kernel duo $B > gpurun_out/r2d_soup_duo_b$b.log 2>&1; line "soup duo budget $b" gpurun_out/r2d_soup_duo_b$b.log
done
for m in 4 8 16 24 32; do
  BVH_B200_REFILL_MIN=$m timeout 300 python bench.py --kernel persistent $B > gpurun_out/r2d_soup_solo_r$m.log 2>&1; line "soup solo refill_min $m" gpurun_out/r2d_soup_solo_r$m.log
  BVH_B200_REFILL_MIN=$m timeout 300 python bench.py --kernel duo $B > gpurun_out/r2d_soup_duo_r$m.log 2>&1; line "soup duo refill_min $m" gpurun_out/r2d_soup_duo_r$m.log
done
BVH_B200_REFILL_MIN=16 timeout 300 python bench.py --mesh grid --kernel persistent $B > gpurun_out/r2d_grid_solo_r16.log 2>&1; line "grid solo refill_min 16" gpurun_out/r2d_grid_solo_r16.log
timeout 400 python bench.py --config c3 --kernel duo $B > gpurun_out/r2d_c3_duo.log 2>&1; line "c3 duo" gpurun_out/r2d_c3_duo.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2d_duo python bench.py --kernel duo --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2d_ncu_duo.log 2>&1
echo "ncu duo rc=$?"
