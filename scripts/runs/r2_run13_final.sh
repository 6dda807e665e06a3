# round 2, GPU run 13: final state — full GPU test suite, smoke, bench lines of every config, reference arm, ncu launch list and capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2q_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2q_pytest.log | cut -c1-150 | head -20
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2q_smoke.log)"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1) build $(grep -o '"ms": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2)"; }
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2q_full.log 2>&1; line "full default" gpurun_out/r2q_full.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2q_reference.log 2>&1; line "reference arm" gpurun_out/r2q_reference.log
timeout 400 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/r2q_c3.log 2>&1; line "c3" gpurun_out/r2q_c3.log
timeout 400 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r2q_c5.log 2>&1; line "c5" gpurun_out/r2q_c5.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 python bench.py --mesh grid $B > gpurun_out/r2q_grid.log 2>&1; line "grid" gpurun_out/r2q_grid.log
timeout 300 python bench.py --quality low $B > gpurun_out/r2q_soup_low.log 2>&1; line "soup low" gpurun_out/r2q_soup_low.log
timeout 300 python bench.py --kernel wide $B > gpurun_out/r2q_soup_wide.log 2>&1; line "soup wide" gpurun_out/r2q_soup_wide.log
timeout 300 python bench.py --config c3 --kernel wide $B > gpurun_out/r2q_c3_wide.log 2>&1; line "c3 wide" gpurun_out/r2q_c3_wide.log
timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_c4_n1.log 2>&1; line "c4 on one GPU" gpurun_out/r2q_c4_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2q_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2q_persistent python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2q_ncu_persistent.log 2>&1
echo "ncu persistent rc=$?"
