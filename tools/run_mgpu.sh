set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_2gpu_r1.json
cat gpurun_out/bench_2gpu_r1.json | cut -c1-400
python bench.py --impl reference --steps 1 --warmup 1 | cut -c1-300
