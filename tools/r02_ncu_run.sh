#!/bin/bash
# ncu evidence of round 2 (one GPU): launch lists with DRAM bytes (B0, resnet50), section captures of the depthwise and
# tcgen05 kernels, a full-set capture with source of the top depthwise backward launches. Reports are exported to CSV on the
# box (the .ncu-rep files exceed the 64 MiB copy-back limit) and removed.
mkdir -p gpurun_out /tmp/ncu
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
NCU_STEPS=2 timeout 900 ncu --metrics $M --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_b0.csv python tools/ncu_target.py 256 > gpurun_out/ncu_b0.log 2>&1
tail -1 gpurun_out/ncu_b0.log
NCU_STEPS=2 timeout 900 ncu --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_r50.csv python tools/ncu_target.py 256 resnet50 > gpurun_out/ncu_r50.log 2>&1
tail -1 gpurun_out/ncu_r50.log
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --section ComputeWorkloadAnalysis"
timeout 900 ncu $SEC --clock-control none -k regex:dwconv -c 32 -o /tmp/ncu/r02_dw python tools/ncu_target.py 256 > gpurun_out/ncu_dw.log 2>&1
ncu -i /tmp/ncu/r02_dw.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_dw_raw.csv.gz
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dwconv_bwd -s 10 -c 2 -o /tmp/ncu/r02_dwbwd_full python tools/ncu_target.py 256 > gpurun_out/ncu_dwfull.log 2>&1
ncu -i /tmp/ncu/r02_dwbwd_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_dwbwd_full_raw.csv.gz
ncu -i /tmp/ncu/r02_dwbwd_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_dwbwd_full_source.csv.gz
ncu -i /tmp/ncu/r02_dwbwd_full.ncu-rep --page details 2>/dev/null | gzip > gpurun_out/r02_dwbwd_full_details.txt.gz
timeout 900 ncu $SEC --clock-control none -k regex:"gemm_tc_kernel|wgrad_tc_kernel" -c 70 -o /tmp/ncu/r02_r50_tc python tools/ncu_target.py 256 resnet50 > gpurun_out/ncu_r50tc.log 2>&1
ncu -i /tmp/ncu/r02_r50_tc.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_r50_tc_raw.csv.gz
ls -la gpurun_out/ /tmp/ncu
du -sh gpurun_out
