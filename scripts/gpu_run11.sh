mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench_n2=$?"; tail -3 gpurun_out/bench_n2.log | cut -c1-600
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.log 2>&1; echo "bench_n2_ref=$?"; tail -1 gpurun_out/bench_n2_ref.log | cut -c1-300
