"""dfd_ordered_reduce on the table of one 7x7 EfficientNet-B0 block (18 depthwise channel-block entries with 64 partials,
two 1x1 weight gradients with 9 partials): which entries cost the time."""
import os, struct, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_detection_b200 import _lib
st = lambda: torch.cuda.current_stream().cuda_stream


def run(entries, label, reps=20):
    ws = torch.randn(sum(n_stride * parts for (n, n_stride, parts) in entries), device="cuda")
    dst = torch.zeros(sum(n for (n, _, _) in entries), device="cuda")
    raw, wo, do = b"", 0, 0
    for n, stride, parts in entries:
        raw += struct.pack("<QQqqii", ws.data_ptr() + wo * 4, dst.data_ptr() + do * 4, n, stride, parts, 0)
        wo += stride * parts; do += n
    table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    bx = max((n // 4 + 7) // 8 if parts > 64 else (n // 4 + 255) // 256 for (n, _, parts) in entries)
    bx = max(1, min(bx, 1024))
    f = lambda: _lib.call("dfd_ordered_reduce", table.data_ptr(), len(entries), dst.data_ptr(), bx, st())
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print("%-40s entries %2d blocks_x %4d  %.1f us  (%.0f MB -> %.0f GB/s)" % (label, len(entries), bx, us, ws.numel() * 4 / 1e6, ws.numel() * 4 / us / 1e3))


dw = [(1600, 1600, 64)] * 18
wg = [(221184, 221184, 9)] * 2
run(dw + wg, "7x7 block: dw + 1x1 wgrads")
run(dw, "depthwise entries only")
run(wg, "1x1 weight gradients only")
run([(1600, 1600, 128)] * 8 + [(53760, 53760, 13)] * 2, "14x14 block (dw parts 128)")
run([(221184, 221184, 9)], "one 1x1 weight gradient")
