# round 2, second 8-GPU run (short): warp-staged records sent with ONE bulk copy to the NVSwitch multicast address
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --gather multicast_staged > gpurun_out/r2p_n8_multicast_staged.log 2>&1
echo "n8 multicast_staged rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2p_n8_multicast_staged.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2p_n8_multicast_staged.log | head -1) $(grep -o '"step_ms": {[^}]*}' gpurun_out/r2p_n8_multicast_staged.log | head -1)"
tail -3 gpurun_out/r2p_n8_multicast_staged.log | cut -c1-300
