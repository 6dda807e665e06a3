# round 2, GPU run 17 (the last minute of the budget): the default bench line of the final build
mkdir -p gpurun_out
timeout 100 python bench.py --steps 20 --warmup 3 > gpurun_out/r2u_full.log 2>&1; echo "rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2u_full.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2u_full.log)"
