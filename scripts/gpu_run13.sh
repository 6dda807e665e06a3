mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n$N.log 2>&1
echo "n$N auto rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_n$N.log | head -1) $(grep -o '"gather": "[^"]*"' gpurun_out/bench_n$N.log | cut -c1-160)"
grep -v '^{' gpurun_out/bench_n$N.log | grep -iv "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -5 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --gather nccl > gpurun_out/bench_n${N}_nccl.log 2>&1
echo "n$N nccl rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_n${N}_nccl.log | head -1)"
