#!/bin/bash
# row-streaming kernels: act_bwd occupancy variants and CTA-size knob, then correctness of the touched kernels and the step time
echo "== default"; timeout 300 python tools/rowops_time.py
echo "== OCC3"; DFD_ACTBWD_OCC3=1 timeout 300 python tools/rowops_time.py
echo "== MAXT 512"; DFD_ROW_MAXT=512 timeout 300 python tools/rowops_time.py
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "act_bwd or bn or pool or se_ or fused" 2>&1 | tail -4
for v in "" "DFD_ACTBWD_OCC3=1" "DFD_ROW_MAXT=512" "DFD_ACTBWD_OCC3=1 DFD_ROW_MAXT=512"; do
  echo "== bench $v"; env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-260
done
