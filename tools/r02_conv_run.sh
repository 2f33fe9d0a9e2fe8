#!/bin/bash
# implicit-GEMM conv: kernel tests, layer timings, resnet engine tests, resnet50 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_implicit or conv_dense" 2>&1 | tail -15
timeout 300 python tools/conv_time.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -k resnet 2>&1 | tail -8
timeout 600 python bench.py --arch resnet50 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -2
DFD_NO_IMPLICIT_CONV=1 timeout 600 python bench.py --arch resnet50 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -2
