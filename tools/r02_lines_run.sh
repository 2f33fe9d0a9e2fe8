#!/bin/bash
# the remaining bench lines of round 2: B4 fp16 b128 (configs[4]), the stock-PyTorch library arm, the CPU reference arm,
# whole-step parity at the B4 size
mkdir -p gpurun_out
timeout 900 python bench.py --arch efficientnet_b4 --dtype fp16 --batch 128 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_b4.json; cut -c1-300 gpurun_out/bench_b4.json
timeout 600 python bench.py --impl library --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_lib_b0.json; cut -c1-300 gpurun_out/bench_lib_b0.json
timeout 600 python bench.py --impl library --arch resnet50 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_lib_r50.json; cut -c1-300 gpurun_out/bench_lib_r50.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_b0.json; cut -c1-400 gpurun_out/bench_ref_b0.json
timeout 1500 python tools/parity_full.py efficientnet_b4 128 380 fp16 2>&1 | tail -3
