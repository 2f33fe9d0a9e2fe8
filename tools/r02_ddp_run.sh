#!/bin/bash
# 2-GPU: the NCCL DDP parity test, then the weak-scaling bench lines (B0 and resnet50)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_boundary_gpu.py -q -m gpu -k two_ranks 2>&1 | tail -5
for arch in efficientnet_b0 resnet50; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --arch $arch 2>&1 | tail -1 > gpurun_out/bench_${arch}_2gpu.json
cut -c1-400 gpurun_out/bench_${arch}_2gpu.json
done
