#!/bin/bash
# 2-GPU lines of the final build (+ the DDP parity test)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_boundary_gpu.py -q -m gpu -k two_ranks 2>&1 | tail -2
for arch in efficientnet_b0 resnet50; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --arch $arch 2>&1 | tail -1 > gpurun_out/final_${arch}_2gpu.json
cut -c1-220 gpurun_out/final_${arch}_2gpu.json
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_dense or conv_implicit" 2>&1 | tail -2
