"""A/B timing + correctness of the depthwise forward kernels (CUDA-core vs tcgen05) on the B0 stride-1 layer shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from deepfake_detection_b200 import _lib
import gpu_checks as GC

for (N, H, W, C, k, aff) in [(2, 16, 16, 32, 3, True), (2, 14, 14, 144, 5, True), (3, 7, 7, 1152, 5, True), (2, 40, 40, 32, 3, False),
                             (1, 33, 33, 24, 3, False), (2, 28, 28, 240, 5, True)]:
    try:
        r = GC.check_dwconv(N, H, W, C, k, 1, affine=aff, fwd_impl="dfd_dwconv_fwd_tc")
        print("CHECK", (N, H, W, C, k, aff), {kk: round(vv, 5) for kk, vv in r.items() if kk in ("fwd_max", "fwd_rel", "nan", "sum_rel", "sq_rel")}, flush=True)
    except Exception as e:
        print("CHECK FAIL", (N, H, W, C, k, aff), repr(e)[:300], flush=True)
        torch.cuda.synchronize()

def t(impl, N, H, W, C, k, reps=10):
    x = torch.randn(N, H, W, C, device="cuda").bfloat16(); w = torch.randn(C, 1, k, k, device="cuda") * 0.2
    sc = torch.rand(C, device="cuda") + 0.5; sh = torch.randn(C, device="cuda") * 0.1
    out = torch.empty(N, H, W, C, device="cuda", dtype=torch.bfloat16)
    s1 = torch.zeros(8, C, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.call(impl, x.data_ptr(), sc.data_ptr(), sh.data_ptr(), w.data_ptr(), out.data_ptr(), N, H, W, C, k, 1, 1, 0, s1.data_ptr(), s2.data_ptr(), st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("TIME %-18s N=%d %dx%dx%d k%d  ms=%.3f GB/s=%.0f" % (impl, N, H, W, C, k, ms, 4 * N * H * W * C / ms / 1e6), flush=True)

for shp in [(256, 112, 112, 32, 3), (256, 56, 56, 144, 3), (256, 28, 28, 240, 5), (256, 14, 14, 480, 3), (256, 14, 14, 672, 5), (256, 7, 7, 1152, 5)]:
    for impl in ("dfd_dwconv_fwd", "dfd_dwconv_fwd_tc"):
        try:
            t(impl, *shp)
        except Exception as e:
            print("TIME FAIL", impl, shp, repr(e)[:200]); torch.cuda.synchronize()
