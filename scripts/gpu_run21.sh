mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest21.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest21.log)"
timeout 600 python bench.py > gpurun_out/bench21.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/bench21.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['build']['ms'], d['roofline']['frac'], d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench21_ref.log 2>&1
echo "ref rc=$?"; grep '^{' gpurun_out/bench21_ref.log | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 3 -c 1 -o gpurun_out/prof_trace_final -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_trace_final.log 2>&1; echo "ncu_trace=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hierarchy|rs_scatter_kernel|rs_scan_bins|rs_tile_hist|morton_kernel|centre_bounds" -s 15 -c 15 -o gpurun_out/prof_build_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_build_final.log 2>&1; echo "ncu_build=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_final.log 2>&1; echo "ncu_launches=$?"
