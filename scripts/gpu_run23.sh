mkdir -p gpurun_out
N=${1:-4}
port=29600
for fwd in 1 0; do
  port=$((port+1))
  BVH_B200_GATHER_FORWARD=$fwd timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/bench23_n${N}_fwd$fwd.log 2>&1
  echo "n$N forward=$fwd rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench23_n${N}_fwd$fwd.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench23_n${N}_fwd$fwd.log | head -1) $(grep -o '"hits_checksum": [0-9]*' gpurun_out/bench23_n${N}_fwd$fwd.log | head -1)"
  grep -v '^{' gpurun_out/bench23_n${N}_fwd$fwd.log | grep -iv "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -3 | cut -c1-300
done
