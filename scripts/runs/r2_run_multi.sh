# round 2, multi-GPU run (N = $1, default 2): multi-rank gather test, default bench line (twice), gather modes, c4 sharded, c3
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2n_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -q > gpurun_out/r2n_pytest_n$N.log 2>&1
echo "pytest multi-gpu rc=$? $(tail -1 gpurun_out/r2n_pytest_n$N.log)"
port=29900
run() {  # name, extra args...
  name=$1; shift; port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/r2n_n${N}_$name.log 2>&1
  echo "n$N $name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2n_n${N}_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2n_n${N}_$name.log | head -1) $(grep -o '"step_ms": {[^}]*}' gpurun_out/r2n_n${N}_$name.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2n_n${N}_$name.log)"
}
run full --steps 20 --warmup 3
run full_again --steps 20 --warmup 3 --no-e2e
for mode in direct peer multicast; do
  run gather_$mode --steps 20 --warmup 3 --no-e2e --gather $mode
done
run c4 --config c4 --steps 5 --warmup 3 --no-e2e
run c3 --config c3 --steps 10 --warmup 3 --no-e2e
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_n1_on_n$N.log 2>&1
echo "n1 rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2n_n1_on_n$N.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2n_n1_on_n$N.log)"
