for d in 0 1 4 8 12 5; do DFD_DBG=$d GT_ONE=1 python tools/gemm_time.py 2>&1 | grep "stats=False"; done
