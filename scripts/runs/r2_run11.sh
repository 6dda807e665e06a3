# round 2, GPU run 11: a warp's private run of consecutive rays (64 / 128 / 256) with and without TMA staging, L1::no_allocate node loads
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
line() { echo "$1: rc=$? $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"median": [0-9.]*' $2 | head -1) $(grep -o '"kernel": "[^"]*"' $2 | head -1)"; }
for v in 1 0; do for ch in 32 64 128 256; do
  export BVH_B200_VARIANT=$v BVH_B200_CHUNK_RAYS=$ch
  for cfg in "--mesh soup" "--mesh grid" "--config c3"; do
    name=$(echo $cfg | tr -d ' -' )
    timeout 300 python bench.py $cfg $B > gpurun_out/r2k_v${v}_ch${ch}_$name.log 2>&1; line "variant $v chunk $ch $cfg" gpurun_out/r2k_v${v}_ch${ch}_$name.log
  done
done; done
export BVH_B200_VARIANT=0 BVH_B200_CHUNK_RAYS=128 BVH_B200_NODE_NA=1
for cfg in "--mesh soup" "--mesh grid" "--config c3"; do
  name=$(echo $cfg | tr -d ' -' )
  timeout 300 python bench.py $cfg $B > gpurun_out/r2k_na_$name.log 2>&1; line "variant 0 chunk 128 no_allocate $cfg" gpurun_out/r2k_na_$name.log
done
export BVH_B200_SMEM_CARVEOUT=100
timeout 300 python bench.py --mesh soup $B > gpurun_out/r2k_na_c100_soup.log 2>&1; line "variant 0 chunk 128 no_allocate carveout 100 soup" gpurun_out/r2k_na_c100_soup.log
unset BVH_B200_NODE_NA
for r in 6 8 12; do
BVH_B200_REFILL_MIN=$r timeout 300 python bench.py --mesh soup $B > gpurun_out/r2k_c100_soup_r$r.log 2>&1; line "variant 0 carveout 100 soup refill $r" gpurun_out/r2k_c100_soup_r$r.log
done
unset BVH_B200_SMEM_CARVEOUT BVH_B200_VARIANT BVH_B200_CHUNK_RAYS
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r2k_pytest.log)"
