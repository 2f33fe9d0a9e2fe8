set -x
DFD_PROFILE_OUT=gpurun_out/per_op_r02a.txt python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-400
python tools/dwbwd_time.py 2>&1 | grep " atm" | grep "k5"
DFD_DW_WSM5=1 python tools/dwbwd_time.py 2>&1 | grep " atm" | grep "k5"
DFD_DW_WSM5=1 DFD_DW_PB5=4 python tools/dwbwd_time.py 2>&1 | grep " atm" | grep "k5"
DFD_DW_WSM5=1 DFD_DW_PB5=4 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-330
DFD_DW_WSM5=1 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-330
