# usage (on the GPU box): bash tools/perf_sweep.sh  -- depthwise lane-mapping variants and the whole step
python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=line -k "dwconv" 2>&1 | tail -8
for v in 32 0; do
  echo "== DFD_DW_CPW=$v (0 = heuristic)"
  DFD_DW_CPW=$v python tools/dwbwd_time.py 2>&1 | grep " det" | head -3
  DFD_DW_CPW=$v python tools/dwfwd_time.py 2>&1 | head -4
done
echo "== C=144 with cpw 16"; DFD_DW_CPW=16 python tools/dwbwd_time.py 2>&1 | grep " det" | sed -n 2,3p
DFD_DW_CPW=16 python tools/dwfwd_time.py 2>&1 | sed -n 3,4p
python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-1900
DFD_DW_CPW=32 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-330
