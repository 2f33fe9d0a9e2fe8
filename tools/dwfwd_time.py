"""Timing of the depthwise forward (dfd_dwconv_fwd) on the EfficientNet-B0 layer shapes, batch 256."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_detection_b200 import _lib

def t(N, H, W, C, k, s, reps=10):
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = torch.randn(N, H, W, C, device="cuda").bfloat16(); w = torch.randn(C, 1, k, k, device="cuda") * 0.2
    sc = torch.rand(C, device="cuda") + 0.5; sh = torch.randn(C, device="cuda") * 0.1
    out = torch.empty(N, Ho, Wo, C, device="cuda", dtype=torch.bfloat16)
    s1 = torch.zeros(8, C, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.call("dfd_dwconv_fwd", x.data_ptr(), sc.data_ptr(), sh.data_ptr(), w.data_ptr(), out.data_ptr(), N, H, W, C, k, s, 1, 0,
                          s1.data_ptr(), s2.data_ptr(), None, st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("FWD %dx%dx%d k%d s%d ms=%.3f GB/s=%.0f" % (H, W, C, k, s, ms, 2 * N * C * (Ho * Wo + H * W) / ms / 1e6), flush=True)
    return ms

tot = 0
for shp in [(256, 112, 112, 32, 3, 1), (256, 112, 112, 96, 3, 2), (256, 56, 56, 144, 3, 1), (256, 56, 56, 144, 5, 2), (256, 28, 28, 240, 5, 1),
            (256, 28, 28, 240, 3, 2), (256, 14, 14, 480, 3, 1), (256, 14, 14, 480, 5, 1), (256, 14, 14, 672, 5, 1), (256, 14, 14, 672, 5, 2),
            (256, 7, 7, 1152, 5, 1), (256, 7, 7, 1152, 3, 1)]:
    tot += t(*shp)
print("SUM %.3f" % tot)
