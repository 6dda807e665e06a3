# round 2, GPU run 9: why 3157 -> 3057 Mrays/s between runs 6 and 7?  stack rounding (smem per block) x shared-memory carve-out
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1)"; }
for combo in "2 -1" "8 -1" "2 100" "2 86" "4 -1"; do
  set -- $combo
  export BVH_B200_STACK_ROUND=$1 BVH_B200_SMEM_CARVEOUT=$2
  for cfg in "--mesh soup" "--mesh grid" "--config c3" "--mesh soup --kernel wide"; do
    name=$(echo $cfg | tr -d ' -' )
    timeout 300 python bench.py $cfg $B > gpurun_out/r2i_r$1_c$2_$name.log 2>&1; line "round $1 carveout $2 $cfg" gpurun_out/r2i_r$1_c$2_$name.log
  done
done
unset BVH_B200_STACK_ROUND BVH_B200_SMEM_CARVEOUT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2i_pytest.log)"
timeout 300 python scripts/gather_probe.py > gpurun_out/r2i_gather_probe.log 2>&1; echo "gather probe rc=$?"; tail -5 gpurun_out/r2i_gather_probe.log
