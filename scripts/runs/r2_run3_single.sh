# round 2, GPU run 3: binary default again, wide opt-in with I2F dequantisation, windowed one-sweep look-back, treelet occupancy
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2c_pytest.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2c_pytest.log | cut -c1-150 | head -20
timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck: $(grep 'ERROR SUMMARY' gpurun_out/r2c_memcheck.log | tail -1)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gather_entry_point or nan_and_degenerate or sorted_ray" > gpurun_out/r2c_memcheck_gather.log 2>&1; echo "memcheck gather/degenerate/sorted: $(grep 'ERROR SUMMARY' gpurun_out/r2c_memcheck_gather.log | tail -1) $(tail -1 gpurun_out/r2c_memcheck_gather.log | cut -c1-80)"
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -2 | tr '\n' ' ') $(grep -o '"ms": [0-9.]*' $2 | head -1)"; }
for mesh in soup grid; do
  for k in auto wide; do
    timeout 300 python bench.py --mesh $mesh --kernel $k $B > gpurun_out/r2c_bench_${mesh}_$k.log 2>&1; line "$mesh high $k" gpurun_out/r2c_bench_${mesh}_$k.log
  done
done
timeout 300 python bench.py --quality low --kernel wide $B > gpurun_out/r2c_bench_soup_low_wide.log 2>&1; line "soup low wide" gpurun_out/r2c_bench_soup_low_wide.log
timeout 300 python bench.py --quality low $B > gpurun_out/r2c_bench_soup_low.log 2>&1; line "soup low" gpurun_out/r2c_bench_soup_low.log
BVH_B200_SORT_ONESWEEP=0 timeout 300 python bench.py --quality low $B > gpurun_out/r2c_bench_soup_low_sort3.log 2>&1; line "soup low 3-kernel sort" gpurun_out/r2c_bench_soup_low_sort3.log
for tb in 2 4; do
  BVH_B200_TREELET_BLOCKS=$tb timeout 300 python bench.py $B > gpurun_out/r2c_bench_soup_tb$tb.log 2>&1; line "soup high treelet blocks $tb" gpurun_out/r2c_bench_soup_tb$tb.log
done
for wb in 3 5; do
  BVH_B200_WIDE_BUDGET=$wb timeout 300 python bench.py --kernel wide $B > gpurun_out/r2c_bench_soup_wide_b$wb.log 2>&1; line "soup wide budget $wb" gpurun_out/r2c_bench_soup_wide_b$wb.log
done
for c in c3 c5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2c_bench_$c.log 2>&1; line "$c" gpurun_out/r2c_bench_$c.log
done
timeout 400 python bench.py --config c3 --kernel wide --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_c3_wide.log 2>&1; line "c3 wide" gpurun_out/r2c_bench_c3_wide.log
timeout 400 python bench.py --config c3 --sort-rays --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_c3_sorted.log 2>&1; line "c3 sorted" gpurun_out/r2c_bench_c3_sorted.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2c_bench_full.log 2>&1; line "full default" gpurun_out/r2c_bench_full.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2c_bench_reference.log 2>&1; line "reference arm" gpurun_out/r2c_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2c_persistent python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_persistent.log 2>&1
echo "ncu persistent rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_wide_kernel -s 1 -c 1 -o gpurun_out/r2c_wide python bench.py --kernel wide --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_wide.log 2>&1
echo "ncu wide rc=$?"
