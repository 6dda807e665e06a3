# round 2, 8-GPU run: default bench line (e2e included), gather modes, c4 sharded 8-way, 4-GPU line on the same box
N=8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2o_topo_n8.txt 2>&1
port=29900
run() {  # ranks, name, extra args...
  n=$1; name=$2; shift; shift; port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/r2o_n${n}_$name.log 2>&1
  echo "n$n $name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2o_n${n}_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2o_n${n}_$name.log | head -1) $(grep -o '"step_ms": {[^}]*}' gpurun_out/r2o_n${n}_$name.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2o_n${n}_$name.log)"
}
run 8 full --steps 20 --warmup 3
run 8 gather_direct --steps 20 --warmup 3 --no-e2e --gather direct
run 8 gather_multicast --steps 20 --warmup 3 --no-e2e --gather multicast
run 8 c4 --config c4 --steps 5 --warmup 3 --no-e2e
run 4 full --steps 20 --warmup 3
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2o_n1_on_n8.log 2>&1
echo "n1 rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2o_n1_on_n8.log | head -1)"
