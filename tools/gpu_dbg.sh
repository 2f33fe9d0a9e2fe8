#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== A eager nonblocking small"; timeout 300 python bench.py --no-graph --no-cpu --steps 3 --batch 32 --res 64 --gemm mma > gpurun_out/dbgA.log 2>&1; echo "exit $?"; tail -2 gpurun_out/dbgA.log | cut -c1-300
echo "== B graph small"; timeout 300 python bench.py --no-cpu --steps 3 --batch 32 --res 64 --gemm mma > gpurun_out/dbgB.log 2>&1; echo "exit $?"; tail -2 gpurun_out/dbgB.log | cut -c1-300
echo "== C eager nonblocking full"; timeout 300 python bench.py --no-graph --no-cpu --steps 3 --gemm mma > gpurun_out/dbgC.log 2>&1; echo "exit $?"; tail -2 gpurun_out/dbgC.log | cut -c1-300
echo "== D sanitizer graph small"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --no-cpu --steps 2 --batch 8 --res 64 --gemm mma > gpurun_out/dbgD.log 2>&1; echo "exit $?"; grep -A12 "Invalid\|Error" gpurun_out/dbgD.log | head -60 | cut -c1-300
