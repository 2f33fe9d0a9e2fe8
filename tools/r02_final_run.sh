#!/bin/bash
# final evidence of round 2 on ONE GPU: full test suite, bench lines with per-op profiles, ncu launch lists of the final build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
DFD_PROFILE_OUT=gpurun_out/final_per_op_b0.txt timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/final_b0.json; cut -c1-200 gpurun_out/final_b0.json
DFD_PROFILE_OUT=gpurun_out/final_per_op_r50.txt timeout 600 python bench.py --arch resnet50 --steps 20 2>&1 | tail -1 > gpurun_out/final_r50.json; cut -c1-200 gpurun_out/final_r50.json
DFD_PROFILE_OUT=gpurun_out/final_per_op_b4.txt timeout 900 python bench.py --arch efficientnet_b4 --dtype fp16 --batch 128 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/final_b4.json; cut -c1-200 gpurun_out/final_b4.json
timeout 600 python bench.py --arch resnet18 --steps 20 --no-cpu 2>&1 | tail -1 > gpurun_out/final_r18.json; cut -c1-200 gpurun_out/final_r18.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --cpu-world 2 2>&1 | tail -1 > gpurun_out/final_ref_b0.json; cut -c1-200 gpurun_out/final_ref_b0.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
NCU_STEPS=2 timeout 900 ncu --metrics $M --clock-control none -c 1600 --csv --log-file gpurun_out/final_launches_b0.csv python tools/ncu_target.py 256 > gpurun_out/ncu_b0.log 2>&1; tail -1 gpurun_out/ncu_b0.log
NCU_STEPS=2 timeout 900 ncu --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/final_launches_r50.csv python tools/ncu_target.py 256 resnet50 > gpurun_out/ncu_r50.log 2>&1; tail -1 gpurun_out/ncu_r50.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
