mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/bench2_final.log 2>&1; echo rc=$?
grep '^{' gpurun_out/bench2_final.log | tail -1 | cut -c1-1500
