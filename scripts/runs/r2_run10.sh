# round 2, GPU run 10: how far does "more L1" go?  streaming ray loads (no TMA staging buffers: 15 KB of shared memory per block)
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1)"; }
for combo in "0 -1" "0 100" "0 60" "0 45" "1 -1"; do
  set -- $combo
  export BVH_B200_VARIANT=$1 BVH_B200_SMEM_CARVEOUT=$2
  for cfg in "--mesh soup" "--mesh grid" "--config c3"; do
    name=$(echo $cfg | tr -d ' -' )
    timeout 300 python bench.py $cfg $B > gpurun_out/r2j_v$1_c$2_$name.log 2>&1; line "variant $1 carveout $2 $cfg" gpurun_out/r2j_v$1_c$2_$name.log
  done
done
export BVH_B200_VARIANT=1 BVH_B200_SMEM_CARVEOUT=-1
for r in 16 24; do
BVH_B200_REFILL_MIN=$r timeout 300 python bench.py --config c3 $B > gpurun_out/r2j_c3_r$r.log 2>&1; line "c3 refill $r" gpurun_out/r2j_c3_r$r.log
done
timeout 300 python bench.py --config c3 --sort-rays $B > gpurun_out/r2j_c3_sorted.log 2>&1; line "c3 sorted" gpurun_out/r2j_c3_sorted.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent_kernel -s 1 -c 1 -o gpurun_out/r2j_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2j_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
