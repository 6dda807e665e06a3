mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench24_n2.log 2>&1
echo "n2 full rc=$? $(grep '^{' gpurun_out/bench24_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'], 'clocks', d['clocks'])")"
grep -v '^{' gpurun_out/bench24_n2.log | grep -iv "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -4 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench24_n2_ref.log 2>&1
echo "n2 ref rc=$? $(grep -c '^{' gpurun_out/bench24_n2_ref.log) line(s)"
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest24.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest24.log)"
