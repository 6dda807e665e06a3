# round 2, GPU run 15 (last): resident blocks per SM of the persistent kernel — 8 (64 registers), 9 (56), 10 (48, 34 bytes of spills)
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1)"; }
for lib in default b9 b10; do
  if [ $lib = default ]; then export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c.so; else export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c_$lib.so; fi
  for cfg in "--mesh soup" "--mesh grid" "--config c3"; do
    name=$(echo $cfg | tr -d ' -' )
    timeout 200 python bench.py $cfg $B > gpurun_out/r2s_${lib}_$name.log 2>&1; line "$lib $cfg" gpurun_out/r2s_${lib}_$name.log
  done
done
for lib in b9 b10; do
  BVH_B200_LIB=$PWD/bvh_b200/libbvh_c_$lib.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2s_pytest_$lib.log 2>&1
  echo "pytest $lib rc=$? $(tail -1 gpurun_out/r2s_pytest_$lib.log)"
done
