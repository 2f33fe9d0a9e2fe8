# usage (on the GPU box): bash tools/perf_sweep.sh  -- depthwise lane-mapping variants, layer by layer
for v in 32 16; do
  echo "== DFD_DW_CPW=$v"
  DFD_DW_CPW=$v python tools/dwbwd_time.py 2>&1 | grep " det"
  DFD_DW_CPW=$v python tools/dwfwd_time.py 2>&1
done
python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-400
