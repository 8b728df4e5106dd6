python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_8gpu_r1.json
cut -c1-330 gpurun_out/bench_8gpu_r1.json
