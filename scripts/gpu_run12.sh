mkdir -p gpurun_out
port=29520
for mode in multicast peer nccl; do
  port=$((port+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 20 --warmup 3 --gather $mode --no-e2e > gpurun_out/bench_n2_$mode.log 2>&1
  echo "n2 $mode rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_n2_$mode.log | head -1) $(grep -o '"gather": "[^"]*"' gpurun_out/bench_n2_$mode.log | cut -c1-120)"
  tail -4 gpurun_out/bench_n2_$mode.log | grep -v '^{' | cut -c1-300
done
