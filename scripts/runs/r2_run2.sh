# round 2, GPU run 2: compaction + warp treelets + one-sweep sort + staged gather + optimizer on the GPU suite; wide kernel
# on the treelet tree; c3 / c5 lines; launch list; ncu of the treelet pass and of the default traversal kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2b_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2b_pytest.log | head -20
if grep -qE "^(FAILED|ERROR)" gpurun_out/r2b_pytest.log; then
  # localise: the same parity tests with one new component switched off at a time
  K="golden or valid_reference_bvh or treelets or tiny or switches"
  BVH_B200_SORT_ONESWEEP=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K" > gpurun_out/r2b_pytest_sort3.log 2>&1; echo "3-kernel sort: $(tail -1 gpurun_out/r2b_pytest_sort3.log)"
  BVH_B200_SAH_TREELETS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or tiny or ragged or refit or gather" > gpurun_out/r2b_pytest_notreelets.log 2>&1; echo "no treelets: $(tail -1 gpurun_out/r2b_pytest_notreelets.log)"
  BVH_B200_GATHER_STAGING=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gather" > gpurun_out/r2b_pytest_nostaging.log 2>&1; echo "direct gather stores: $(tail -1 gpurun_out/r2b_pytest_nostaging.log)"
  timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck: $(grep 'ERROR SUMMARY' gpurun_out/r2b_memcheck.log | tail -1)"
fi
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -2 | tr '\n' ' ') $(grep -o '"ms": [0-9.]*' $2) $(grep -o '"inner_steps_per_ray": [0-9.]*' $2)"; }
for mesh in soup grid; do
  for k in auto wide persistent; do
    timeout 300 python bench.py --mesh $mesh --kernel $k $B > gpurun_out/r2b_bench_${mesh}_$k.log 2>&1; line "$mesh high $k" gpurun_out/r2b_bench_${mesh}_$k.log
  done
done
for wb in 2 3 6; do
  BVH_B200_WIDE_BUDGET=$wb timeout 300 python bench.py --kernel wide $B > gpurun_out/r2b_bench_soup_wide_b$wb.log 2>&1; line "soup wide budget $wb" gpurun_out/r2b_bench_soup_wide_b$wb.log
done
timeout 300 python bench.py --quality low --kernel wide $B > gpurun_out/r2b_bench_soup_low_wide.log 2>&1; line "soup low wide" gpurun_out/r2b_bench_soup_low_wide.log
timeout 300 python bench.py --quality low --kernel persistent $B > gpurun_out/r2b_bench_soup_low_bin.log 2>&1; line "soup low binary" gpurun_out/r2b_bench_soup_low_bin.log
BVH_B200_SPECULATE=1 timeout 300 python bench.py --kernel wide $B > gpurun_out/r2b_bench_soup_wide_spec.log 2>&1; line "soup wide speculate" gpurun_out/r2b_bench_soup_wide_spec.log
BVH_B200_SPECULATE=1 timeout 300 python bench.py --mesh grid --kernel wide $B > gpurun_out/r2b_bench_grid_wide_spec.log 2>&1; line "grid wide speculate" gpurun_out/r2b_bench_grid_wide_spec.log
BVH_B200_SORT_ONESWEEP=0 timeout 300 python bench.py --quality low --kernel persistent $B > gpurun_out/r2b_bench_soup_low_sort3.log 2>&1; line "soup low 3-kernel sort" gpurun_out/r2b_bench_soup_low_sort3.log
for c in c3 c5; do
  for k in auto wide; do
    [ $c = c5 ] && [ $k = wide ] && continue
    timeout 400 python bench.py --config $c --kernel $k --steps 10 --warmup 3 --no-e2e > gpurun_out/r2b_bench_${c}_$k.log 2>&1; line "$c $k" gpurun_out/r2b_bench_${c}_$k.log
  done
done
for k in auto wide; do
  timeout 400 python bench.py --config c3 --kernel $k --sort-rays --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_c3_${k}_sorted.log 2>&1; line "c3 $k sorted rays" gpurun_out/r2b_bench_c3_${k}_sorted.log
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:treelet_kernel -s 2 -c 1 -o gpurun_out/r2b_treelet python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_treelet.log 2>&1
echo "ncu treelet rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_wide_kernel -s 1 -c 1 -o gpurun_out/r2b_wide python bench.py --kernel wide --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_wide.log 2>&1
echo "ncu wide rc=$?"
