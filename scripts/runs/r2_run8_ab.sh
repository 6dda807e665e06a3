# round 2, GPU run 8: A/B of the 64-byte triangle records and of the "memory" clobbers of the stack asm (three builds)
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1)"; }
for lib in default tri48 clobber; do
  if [ $lib = default ]; then export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c.so; else export BVH_B200_LIB=$PWD/bvh_b200/libbvh_c_$lib.so; fi
  for cfg in "--mesh soup" "--mesh grid" "--config c3" "--config c5" "--mesh soup --kernel wide" "--config c3 --kernel wide"; do
    name=$(echo $cfg | tr -d ' -' )
    timeout 300 python bench.py $cfg $B > gpurun_out/r2h_${lib}_$name.log 2>&1; line "$lib $cfg" gpurun_out/r2h_${lib}_$name.log
  done
done
