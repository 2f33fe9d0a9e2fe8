"""Diagnostics: dump CTA 0's pipeline timestamps of the tcgen05 GEMM for one shape (DFD_TS=1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DFD_TS"] = "1"
from deepfake_detection_b200 import _lib
M, N, K = [int(x) for x in sys.argv[1:4]]
A = torch.randn(M, K, device='cuda').bfloat16(); B = torch.randn(N, K, device='cuda').bfloat16()
C = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("dfd_gemm_tn", A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, 0, None, None, None, st)
