#!/bin/bash
# last refresh after the strided / reduction-store ResNet kernels: full suite, ResNet bench lines + per-op + ncu launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/final3_pytest.log; cat gpurun_out/final3_pytest.log
DFD_PROFILE_OUT=gpurun_out/final3_per_op_r50.txt timeout 600 python bench.py --arch resnet50 --steps 20 2>&1 | tail -1 > gpurun_out/final3_r50.json; cut -c1-200 gpurun_out/final3_r50.json
timeout 600 python bench.py --arch resnet18 --steps 20 --no-cpu 2>&1 | tail -1 > gpurun_out/final3_r18.json; cut -c1-200 gpurun_out/final3_r18.json
timeout 600 python bench.py --steps 30 --no-cpu 2>&1 | tail -1 | cut -c1-200
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
NCU_STEPS=2 timeout 900 ncu --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/final3_launches_r50.csv python tools/ncu_target.py 256 resnet50 > gpurun_out/ncu_r50.log 2>&1; tail -1 gpurun_out/ncu_r50.log
