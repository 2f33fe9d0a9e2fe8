#!/bin/bash
for mb in 4 16 1000; do
  echo "== bucket_mb $mb"
  DFD_DDP_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-250
done
DFD_DDP_BUCKET_MB=1000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --arch resnet50 2>&1 | tail -1 | cut -c1-250
