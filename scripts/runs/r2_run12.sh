# round 2, GPU run 12: streaming ray loads + 64-ray runs + L1::no_allocate node loads as the defaults; A/B: triangle loads no_allocate too,
# node loads allocating (the old form); refill threshold around the new defaults
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1)"; }
for lib in default trina nona; do
  if [ $lib = default ]; then export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c.so; else export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c_$lib.so; fi
  for cfg in "--mesh soup" "--mesh grid" "--config c3" "--config c5" "--mesh soup --kernel wide" "--config c3 --kernel wide"; do
    name=$(echo $cfg | tr -d ' -' )
    timeout 300 python bench.py $cfg $B > gpurun_out/r2l_${lib}_$name.log 2>&1; line "$lib $cfg" gpurun_out/r2l_${lib}_$name.log
  done
done
export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c.so
for r in 6 8 12 14; do
  BVH_B200_REFILL_MIN=$r timeout 300 python bench.py --mesh soup $B > gpurun_out/r2l_soup_r$r.log 2>&1; line "default soup refill $r" gpurun_out/r2l_soup_r$r.log
done
BVH_B200_REFILL_MIN=16 timeout 300 python bench.py --mesh grid $B > gpurun_out/r2l_grid_r16.log 2>&1; line "default grid refill 16" gpurun_out/r2l_grid_r16.log
BVH_B200_INNER_BUDGET=6 timeout 300 python bench.py --mesh soup $B > gpurun_out/r2l_soup_b6.log 2>&1; line "default soup budget 6" gpurun_out/r2l_soup_b6.log
BVH_B200_INNER_BUDGET=10 timeout 300 python bench.py --mesh soup $B > gpurun_out/r2l_soup_b10.log 2>&1; line "default soup budget 10" gpurun_out/r2l_soup_b10.log
BVH_B200_VARIANT=1 timeout 300 python bench.py --mesh soup $B > gpurun_out/r2l_soup_tma.log 2>&1; line "default soup TMA staging" gpurun_out/r2l_soup_tma.log
BVH_B200_CHUNK_RAYS=32 timeout 300 python bench.py --mesh soup $B > gpurun_out/r2l_soup_ch32.log 2>&1; line "default soup run 32" gpurun_out/r2l_soup_ch32.log
BVH_B200_CHUNK_RAYS=128 timeout 300 python bench.py --mesh soup $B > gpurun_out/r2l_soup_ch128.log 2>&1; line "default soup run 128" gpurun_out/r2l_soup_ch128.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2l_pytest.log)"
