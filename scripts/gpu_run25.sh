mkdir -p gpurun_out
N=${1:-4}
port=29800
for mode in peer nccl; do
  port=$((port+1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --gather $mode > gpurun_out/bench25_n${N}_$mode.log 2>&1
  echo "n$N $mode rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench25_n${N}_$mode.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench25_n${N}_$mode.log | head -1)"
done
