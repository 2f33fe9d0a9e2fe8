"""Squeeze-excite FC kernels on the EfficientNet-B0 layer shapes (batch 256): dfd_se_fc_fwd and dfd_se_fc_bwd (+ wgrad)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_detection_b200 import _lib
N = 256
st = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()


def timeit(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tot_f = tot_b = 0.0
for C, Cse, cnt in [(32, 8, 1), (96, 4, 1), (144, 6, 2), (240, 10, 2), (480, 20, 3), (672, 28, 3), (1152, 48, 4)]:
    r = lambda *s: torch.randn(*s, device="cuda")
    pooled, Wr, br, We, be = r(N, C), r(Cse, C) / C ** 0.5, r(Cse), r(C, Cse) / Cse ** 0.5, r(C)
    gate, draw = torch.zeros(N, C, device="cuda"), r(N, C)
    d_e, rr, d_rpre, dpool = torch.zeros(N, C, device="cuda"), torch.zeros(N, Cse, device="cuda"), torch.zeros(N, Cse, device="cuda"), torch.zeros(N, C, device="cuda")
    dWr, dbr, dWe, dbe = torch.zeros_like(Wr), torch.zeros_like(br), torch.zeros_like(We), torch.zeros_like(be)
    f = timeit(lambda: _lib.call("dfd_se_fc_fwd", P(pooled), P(Wr), P(br), P(We), P(be), P(gate), N, C, Cse, st()))
    b = timeit(lambda: _lib.call("dfd_se_fc_bwd", P(draw), P(pooled), P(Wr), P(br), P(We), P(be), P(d_e), P(rr), P(d_rpre), P(dpool),
                                 P(dWr), P(dbr), P(dWe), P(dbe), N, C, Cse, st()))
    tot_f += f * cnt; tot_b += b * cnt
    print("C=%4d Cse=%2d x%d  fwd %6.1f us   bwd+wgrad %6.1f us" % (C, Cse, cnt, f, b), flush=True)
print("B0 total (16 layers): fwd %.0f us  bwd %.0f us" % (tot_f, tot_b))
