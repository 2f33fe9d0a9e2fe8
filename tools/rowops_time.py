"""Row-streaming elementwise / reduction kernels on the EfficientNet-B0 layer shapes (batch 256, bf16): time and achieved
bandwidth of dfd_act_bwd (swish, SE gate + pooled gradient), dfd_se_bwd_reduce, dfd_pool, dfd_bn_act (gated)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_detection_b200 import _lib

N = 256
st = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=8):
    fn(); fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                      # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


only = os.environ.get("RT_ONLY", "")
for hw, C in [(12544, 32), (3136, 96), (3136, 144), (784, 144), (784, 240), (196, 240), (196, 480), (196, 672), (49, 672), (49, 1152)]:
    y = torch.randn(N, hw, C, device="cuda").bfloat16()
    da = torch.randn(N, hw, C, device="cuda").bfloat16()
    out = torch.empty_like(y)
    f = lambda: torch.rand(C, device="cuda") + 0.5
    scale, shift, mean, rstd = f(), f() - 1, f() - 1, f()
    gate = torch.rand(N, C, device="cuda"); dpool = torch.randn(N, C, device="cuda")
    slots = _lib.lib().stat_slots
    s1 = torch.zeros(slots, C, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    draw = torch.zeros(N, C, device="cuda"); pooled = torch.zeros(N, C, device="cuda")
    part = torch.zeros(8 * N * C, device="cuda")
    nb = y.numel() * 2
    r = {}
    r["act_bwd"] = (timeit(lambda: _lib.call("dfd_act_bwd", P(da), P(y), P(scale), P(shift), P(mean), P(rstd), P(gate), P(dpool), P(out),
                                              N, hw, C, 1, 0, P(s1), P(s2), None, st())), 3 * nb)
    r["se_bwd_reduce"] = (timeit(lambda: (draw.zero_(), _lib.call("dfd_se_bwd_reduce", P(da), P(y), P(scale), P(shift), P(draw), N, hw, C, 0, st()))), 2 * nb)
    r["pool"] = (timeit(lambda: _lib.call("dfd_pool", P(y), P(scale), P(shift), P(pooled), N, hw, C, 1, 0, P(part), 8, st())), nb)
    r["bn_act_gated"] = (timeit(lambda: _lib.call("dfd_bn_act", P(y), P(scale), P(shift), P(gate), None, P(out), N, hw, C, 1, 0, 0, st())), 2 * nb)
    print("hw=%5d C=%4d  " % (hw, C) + "  ".join("%s %6.1f us %5.0f GB/s" % (k, v[0], v[1] / v[0] / 1e3) for k, v in r.items()), flush=True)
