#!/bin/bash
# full GPU suite + default bench with the per-op profile
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/full_pytest.log
cat gpurun_out/full_pytest.log
DFD_PROFILE_OUT=gpurun_out/per_op_b0.txt timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_b0.json
cut -c1-600 gpurun_out/bench_b0.json
DFD_PROFILE_OUT=gpurun_out/per_op_r50.txt timeout 600 python bench.py --arch resnet50 --steps 20 2>&1 | tail -1 > gpurun_out/bench_r50.json
cut -c1-300 gpurun_out/bench_r50.json
