#!/bin/bash
cd deepfake_detection_b200
cp libdfd_b200.so /tmp/lib_u3.so
cd ..
echo "== U3"; python bench.py --arch resnet50 --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-200
cp deepfake_detection_b200/libdfd_b200_u4.so deepfake_detection_b200/libdfd_b200.so
echo "== U4"; python bench.py --arch resnet50 --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-200
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bn_chain or act_bwd" 2>&1 | tail -2
