#!/bin/bash
mkdir -p gpurun_out
for arch in efficientnet_b0 resnet50; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 30 --warmup 5 --arch $arch --no-cpu 2>&1 | tail -1 > gpurun_out/bench_${arch}_8gpu.json
cut -c1-300 gpurun_out/bench_${arch}_8gpu.json
done
