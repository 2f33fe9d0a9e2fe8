"""Timing of the fused depthwise backward (dfd_dwconv_bwd) on the EfficientNet-B0 layer shapes, batch 256."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_detection_b200 import _lib

def t(N, H, W, C, k, s, reps=10, det=True):
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = torch.randn(N, H, W, C, device="cuda").bfloat16(); w = torch.randn(C, 1, k, k, device="cuda") * 0.2
    gy = torch.randn(N, Ho, Wo, C, device="cuda").bfloat16(); y = torch.randn(N, Ho, Wo, C, device="cuda").bfloat16()
    gx = torch.empty_like(x); dW = torch.zeros_like(w)
    v = [torch.rand(C, device="cuda") + 0.5 for _ in range(7)]
    s1 = torch.zeros(8, C, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    st = torch.cuda.current_stream().cuda_stream
    import struct
    parts = _lib.lib().cdll.dfd_dwconv_bwd_parts(N, H, W, C, k, s)
    cw = _lib.lib().cdll.dfd_dwconv_block_channels(C)
    cbs = (C + cw - 1) // cw
    ws = torch.empty(cbs * parts * cw * k * k, device="cuda") if det else None
    if det:
        raw = b"".join(struct.pack("<QQqqii", ws.data_ptr() + cb * parts * cw * k * k * 4, dW.data_ptr() + cb * cw * k * k * 4,
                                   min(cw, C - cw * cb) * k * k, cw * k * k, parts, 0) for cb in range(cbs))
        table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    f0 = lambda: _lib.call("dfd_dwconv_bwd", gy.data_ptr(), y.data_ptr(), v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr(), w.data_ptr(),
                          x.data_ptr(), v[3].data_ptr(), v[4].data_ptr(), v[5].data_ptr(), v[6].data_ptr(), None, gx.data_ptr(), dW.data_ptr(),
                          N, H, W, C, k, s, 0, s1.data_ptr(), s2.data_ptr(), ws.data_ptr() if det else None, ws.numel() * 4 if det else 0, None, st)
    def f():
        f0()
        if det:
            _lib.call("dfd_ordered_reduce", table.data_ptr(), cbs, dW.data_ptr(), (cw * k * k // 4 + 7) // 8 if parts > 64 else 1, st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("BWD%s N=%d %dx%dx%d k%d s%d  ms=%.3f GB/s=%.0f" % (" det" if det else " atm", N, H, W, C, k, s, ms, 2 * N * C * (2 * Ho * Wo + 2 * H * W) / ms / 1e6), flush=True)

tot = 0
for shp in [(256, 112, 112, 96, 3, 2), (256, 56, 56, 144, 3, 1), (256, 56, 56, 144, 5, 2), (256, 28, 28, 240, 5, 1), (256, 28, 28, 240, 3, 2),
            (256, 14, 14, 480, 3, 1), (256, 14, 14, 480, 5, 1), (256, 14, 14, 672, 5, 1), (256, 14, 14, 672, 5, 2), (256, 7, 7, 1152, 5, 1),
            (256, 7, 7, 1152, 3, 1)]:
    t(*shp, det=False)
    t(*shp, det=True)
