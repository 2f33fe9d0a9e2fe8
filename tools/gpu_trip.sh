#!/bin/bash
# one GPU round trip: every check group in its own process (a poisoned context must not hide later groups)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
for g in "$@"; do
  case "$g" in
    blocking_*)
      impl="${g#blocking_}"
      CUDA_LAUNCH_BLOCKING=1 timeout 900 python bench.py --steps 2 --warmup 1 --gemm "$impl" --no-graph --no-cpu > gpurun_out/$g.log 2>&1; echo "$g exit $?" ;;
    bench_*)
      impl="${g#bench_}"
      timeout 900 python bench.py --steps 10 --warmup 3 --gemm "$impl" > gpurun_out/$g.log 2>&1; echo "$g exit $?" ;;
    engine_*)
      impl="${g#engine_}"
      timeout 900 python tools/gpu_diag.py --only engine --gemm "$impl" --out gpurun_out/diag_$g > gpurun_out/diag_$g.log 2>&1; echo "$g exit $?" ;;
    *)
      timeout 600 python tools/gpu_diag.py --only "$g" --out gpurun_out/diag_$g > gpurun_out/diag_$g.log 2>&1; echo "$g exit $?" ;;
  esac
  tail -3 gpurun_out/*$g.log | cut -c1-400
done
