# round 2, GPU run 6: select-form inner step + address-based stack with sentinel; refill threshold x inner budget matrix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2f_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2f_pytest.log | cut -c1-150 | head -20
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1)"; }
run() { # name mesh/config-args refill budget
  BVH_B200_REFILL_MIN=$3 BVH_B200_INNER_BUDGET=$4 timeout 300 python bench.py $2 $B > gpurun_out/r2f_$1_r$3_b$4.log 2>&1; line "$1 refill_min $3 budget $4" gpurun_out/r2f_$1_r$3_b$4.log
}
run soup "--mesh soup" 1 12
run grid "--mesh grid" 1 12
for r in 6 8 10; do for b in 6 8 10; do run soup "--mesh soup" $r $b; done; done
for r in 8 12 16 20 24; do for b in 8 12; do run grid "--mesh grid" $r $b; done; done
for r in 1 8 16; do run c3 "--config c3" $r 8; done
run c3 "--config c3" 8 12
for r in 1 8; do run c5 "--config c5" $r 8; done
for r in 1 8 16; do BVH_B200_REFILL_MIN=$r timeout 300 python bench.py --kernel wide $B > gpurun_out/r2f_soup_wide_r$r.log 2>&1; line "soup wide refill_min $r" gpurun_out/r2f_soup_wide_r$r.log; done
BVH_B200_REFILL_MIN=8 timeout 300 python bench.py --config c3 --kernel wide $B > gpurun_out/r2f_c3_wide_r8.log 2>&1; line "c3 wide refill_min 8" gpurun_out/r2f_c3_wide_r8.log
BVH_B200_REFILL_MIN=16 timeout 300 python bench.py --mesh grid --kernel wide $B > gpurun_out/r2f_grid_wide_r16.log 2>&1; line "grid wide refill_min 16" gpurun_out/r2f_grid_wide_r16.log
