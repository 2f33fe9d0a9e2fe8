"""Kernel plan of the ResNet family (resnet18 / resnet50) for `Engine` — host side only.

Reference graph: dfd/timm/models/resnet.py:450-468 (stem 7x7 s2 -> BN -> ReLU -> maxpool 3x3 s2 -> 4 stages -> GAP
-> fc), BasicBlock :150-175, Bottleneck :215-246 (stride on the 3x3, :195-197), downsample 1x1 conv + BN :249-260.

Dense 3x3 convolutions with stride 1 (13 of the 16 in resnet50, all but 3 in resnet18) run as IMPLICIT GEMMs on tcgen05
(`dfd_conv_tc`, csrc/gemm_tc.cu conv mode: the TMA producer fetches the input box shifted by the tap through a 4-D tensor
map, no im2col matrix in memory) in the forward pass and for the input gradient (same kernel on dY with the tap-flipped
[Cin][kh'][kw'][Cout] weights). The strided 3x3 convolutions keep the round-1 formulation (csrc/conv_dense.cu):
materialised im2col -> tcgen05 GEMM (forward), GEMM -> col2im (input gradient). The weight gradient of every 3x3 is the
MN-major tcgen05 wgrad GEMM, implicit too (`dfd_conv_wgrad_tc`: one pipeline stage = one patch of <= 64 output pixels of dY
and the input box shifted by the tap). Stride-2 convolutions (the three strided 3x3 and the strided 1x1 downsample inputs) use
the same kernels with TMA element strides {1, 2, 2, 1} in the forward pass and the weight gradient; the input gradient of a
strided 3x3 is four parity-class implicit GEMMs (`dfd_conv_dgrad_s2_tc`) storing through strided views of dx. Only the strided
1x1 downsample input gradient keeps GEMM + col2im (a scatter fused with the main-path add). 1x1 convolutions are plain GEMMs on the NHWC tensors. BN + ReLU outputs are materialised (`dfd_bn_act`) because three consumers read them.
"""
import os
import struct

import torch

from . import _lib
from .engine import ACT_NONE, ACT_RELU, _ptr


def conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def build_resnet(e):
    spec, N, dev, dt = e.spec, e.N, e.device, e.dt
    e._keep = []
    e.acts = {}
    fwd, bwd = [], []
    mom, eps = e.bn_momentum, e.bn_eps

    # ---- packed (kh, kw, ci) copies of the k x k weights -------------------------------------------------
    pk_off, off = {}, 0
    for n in e.param_names:
        o, s, k = e.p_off[n]
        if len(s) == 4 and s[2] > 1 and not n.startswith("conv1."):
            pk_off[n] = (off, s[0], s[1], s[2])
            off += (k + 7) // 8 * 8
    ar = e.arena        # the packed layouts depend on the parameter shapes only: owned and refreshed by the arena engine
    if getattr(ar, "wpack16", None) is None:
        ar.wpack16 = torch.zeros(max(off, 8), dtype=e.tdtype, device=dev)
        ar.wpackT16 = torch.zeros(max(off, 8), dtype=e.tdtype, device=dev)
        ar.wpackD16 = torch.zeros(max(off, 8), dtype=e.tdtype, device=dev)
        raw = b"".join(struct.pack("<QQQQiiii", _ptr(e.params16, e.p_off[n][0]), _ptr(ar.wpack16, o), _ptr(ar.wpackT16, o),
                                   _ptr(ar.wpackD16, o), O, I, k, 0)
                       for n, (o, O, I, k) in pk_off.items())
        ar._rtable = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        ar._rtable_count = len(pk_off)
        ar._derived_dirty = True
    e.wpack16, e.wpackT16, e.wpackD16 = ar.wpack16, ar.wpackT16, ar.wpackD16
    # two regions: a BasicBlock has two 3x3 convolutions whose packed gradients wait for the block's single ordered reduce
    gperm_max = max([O * I * k * k for (_, O, I, k) in pk_off.values()] + [8])
    e.gperm = torch.zeros(2 * gperm_max, dtype=torch.float32, device=dev)

    P32 = lambda n: _ptr(e.params32, e.p_off[n][0])
    G32 = lambda n: _ptr(e.grads32, e.p_off[n][0])
    P16 = lambda n: _ptr(e.params16, e.p_off[n][0])
    T16 = lambda n: _ptr(e.paramsT16, e.t_off[n][0])
    PK = lambda n: _ptr(e.wpack16, pk_off[n][0])
    PKT = lambda n: _ptr(e.wpackT16, pk_off[n][0])
    PKD = lambda n: _ptr(e.wpackD16, pk_off[n][0])

    # ---- shapes ------------------------------------------------------------------------------------------
    H1, W1 = conv_out(e.H, 7, 2, 3), conv_out(e.W, 7, 2, 3)
    H2, W2 = conv_out(H1, 3, 2, 1), conv_out(W1, 3, 2, 1)
    shapes = []
    h, w = H2, W2
    for b in spec.blocks:
        ho, wo = conv_out(h, 3, b.stride, 1), conv_out(w, 3, b.stride, 1)
        shapes.append((b, h, w, ho, wo))
        h, w = ho, wo
    Hf, Wf = h, w

    # ---- BN arenas ---------------------------------------------------------------------------------------
    bn_specs = [("bn1", 64)]
    for b in spec.blocks:
        if b.kind == "basic":
            bn_specs += [(b.name + ".bn1", b.planes), (b.name + ".bn2", b.cout)]
        else:
            bn_specs += [(b.name + ".bn1", b.planes), (b.name + ".bn2", b.planes), (b.name + ".bn3", b.cout)]
        if b.downsample:
            bn_specs.append((b.name + ".downsample.1", b.cout))
    e._alloc_bn(bn_specs)
    bns = e.bns

    fused_fin = os.environ.get("DFD_FUSED_FINALIZE", "") not in ("", "0", "gemm")   # measured slower than the standalone launches, see engine.py

    def gemm(A, B, C, M, Nn, K, bn=None):
        fs, fq = (bn.fsum, bn.fsq) if bn is not None else (None, None)
        if e.gemm_impl == "tc":
            if bn is not None:
                bn.fused = fused_fin
            return ("dfd_gemm_tn", (A, B, C, M, Nn, K, dt, fs, fq, bn.fin if (bn is not None and fused_fin) else None))
        return ("dfd_gemm_tn_mma", (A, B, C, None, M, Nn, K, dt, fs, fq))

    implicit = e.gemm_impl == "tc" and not os.environ.get("DFD_NO_IMPLICIT_CONV")
    implicit_wgrad = not os.environ.get("DFD_NO_IMPLICIT_WGRAD")
    implicit_s2 = not os.environ.get("DFD_NO_IMPLICIT_S2")          # stride-2 convolutions through TMA element strides
    e.n_implicit = 0

    def conv3x3(xin, name, y, h, w, cin, cout, stride, bn):
        """3x3 / padding 1 forward into y (+ BatchNorm statistics of y)"""
        if implicit and (stride == 1 or implicit_s2) and cin % 64 == 0 and cout % 64 == 0:
            bn.fused = fused_fin
            e.n_implicit += 1
            return [("dfd_conv_tc", (xin, PK(name), y, N, h, w, cin, cout, 3, stride, dt, bn.fsum, bn.fsq,
                                     bn.fin if fused_fin else None))]
        ho, wo = conv_out(h, 3, stride, 1), conv_out(w, 3, stride, 1)
        return [("dfd_im2col", (xin, COLS, N, h, w, cin, 3, stride, 1, dt)),
                gemm(COLS, PK(name), y, N * ho * wo, cout, 9 * cin, bn)]

    def finalize(bn, count):
        # training: finalised by the last CTA of the producing GEMM (bn.fin); this op runs in eval mode only (see Engine._run)
        bn.count = count
        return ("dfd_bn_finalize" + ("_evalonly" if bn.fused else ""),
                [bn.fsum, bn.fsq, float(count), bn.gamma, bn.beta, bn.rm, bn.rv, bn.nbt, mom, eps,
                 "TRAINING", bn.C, bn.scale, bn.shift, bn.mean, bn.rstd])

    def bwd_finalize(bn, count):
        bn.count = count
        if fused_fin:
            return None             # the last CTA of dfd_act_bwd / dfd_bn_bwd_reduce does it (bn.bfin)
        return ("dfd_bn_bwd_finalize", (bn.bs1, bn.bs2, float(count), bn.gamma, bn.mean, bn.rstd, bn.dgamma, bn.dbeta,
                                        bn.cA, bn.cB, bn.cC, bn.C))

    BF = (lambda bn: bn.bfin) if fused_fin else (lambda bn: None)

    def bn_relu(y, bn, out, hw, C):
        return ("dfd_bn_act", (_ptr(y), bn.scale, bn.shift, None, None, _ptr(out), N, hw, C, ACT_RELU, 0, dt))

    # ---- scratch -----------------------------------------------------------------------------------------
    max_act = max([N * H1 * W1 * 64] + [N * hh * ww * max(b.cin, b.planes) for b, hh, ww, ho, wo in shapes] +
                  [N * ho * wo * b.cout for b, hh, ww, ho, wo in shapes])
    max_cols = max([N * ho * wo * 9 * (b.cin if b.kind == "basic" else b.planes) for b, hh, ww, ho, wo in shapes] +
                   [N * ho * wo * 9 * b.planes for b, hh, ww, ho, wo in shapes])
    e.gbuf = [e._alloc16(max_act) for _ in range(6)]
    e.cols = e._alloc16(max_cols)
    COLS = _ptr(e.cols)
    gA, gB, gC, gD, gE, gF = [_ptr(t) for t in e.gbuf]

    # ---- forward -----------------------------------------------------------------------------------------
    e.x_in = torch.zeros(N, spec.in_chans, e.H, e.W, dtype=e.tdtype, device=dev)
    y0 = e._alloc16(N, H1, W1, 64)
    a0 = e._alloc16(N, H1, W1, 64)
    x0 = e._alloc16(N, H2, W2, 64)
    e.pool_idx = torch.zeros(N * H2 * W2 * 64, dtype=torch.uint8, device=dev)
    bn0 = bns["bn1"]
    if e.stem_impl == "gemm":
        taps, Kp = e._stem_gemm_setup("conv1.weight", 64, 7, N * H1 * W1)
        fwd.append(("dfd_stem_im2col", (_ptr(e.x_in), _ptr(e.stem_cols), N, spec.in_chans, e.H, e.W, 7, 2, 3, Kp, dt)))
        fwd.append(gemm(_ptr(e.stem_cols), _ptr(e.stem_wpad), _ptr(y0), N * H1 * W1, 64, Kp, bn0))
    else:
        fwd.append(("dfd_stem_fwd", (_ptr(e.x_in), P32("conv1.weight"), _ptr(y0), N, spec.in_chans, e.H, e.W, 64, 7, 2, 3, dt,
                                     bn0.fsum, bn0.fsq)))
    fwd.append(finalize(bn0, N * H1 * W1))
    fwd.append(bn_relu(y0, bn0, a0, H1 * W1, 64))
    fwd.append(("dfd_maxpool_fwd", (_ptr(a0), _ptr(x0), _ptr(e.pool_idx), N, H1, W1, 64, dt)))
    e.acts["stem.out"] = x0
    x = x0
    recs = []
    for b, h, w, ho, wo in shapes:
        p = b.name
        M1, M2 = N * h * w, N * ho * wo
        rec = dict(b=b, h=h, w=w, ho=ho, wo=wo, x=x)
        if b.kind == "basic":
            bn1, bn2 = bns[p + ".bn1"], bns[p + ".bn2"]
            y1 = e._alloc16(N, ho, wo, b.planes)
            a1 = e._alloc16(N, ho, wo, b.planes)
            y2 = e._alloc16(N, ho, wo, b.cout)
            fwd += conv3x3(_ptr(x), p + ".conv1.weight", _ptr(y1), h, w, b.cin, b.planes, b.stride, bn1)
            fwd.append(finalize(bn1, M2))
            fwd.append(bn_relu(y1, bn1, a1, ho * wo, b.planes))
            fwd += conv3x3(_ptr(a1), p + ".conv2.weight", _ptr(y2), ho, wo, b.planes, b.cout, 1, bn2)
            fwd.append(finalize(bn2, M2))
            rec.update(y1=y1, a1=a1, ylast=y2, bnlast=bn2)
        else:
            bn1, bn2, bn3 = bns[p + ".bn1"], bns[p + ".bn2"], bns[p + ".bn3"]
            y1 = e._alloc16(N, h, w, b.planes)
            a1 = e._alloc16(N, h, w, b.planes)
            y2 = e._alloc16(N, ho, wo, b.planes)
            a2 = e._alloc16(N, ho, wo, b.planes)
            y3 = e._alloc16(N, ho, wo, b.cout)
            fwd.append(gemm(_ptr(x), P16(p + ".conv1.weight"), _ptr(y1), M1, b.planes, b.cin, bn1))
            fwd.append(finalize(bn1, M1))
            fwd.append(bn_relu(y1, bn1, a1, h * w, b.planes))
            fwd += conv3x3(_ptr(a1), p + ".conv2.weight", _ptr(y2), h, w, b.planes, b.planes, b.stride, bn2)
            fwd.append(finalize(bn2, M2))
            fwd.append(bn_relu(y2, bn2, a2, ho * wo, b.planes))
            fwd.append(gemm(_ptr(a2), P16(p + ".conv3.weight"), _ptr(y3), M2, b.cout, b.planes, bn3))
            fwd.append(finalize(bn3, M2))
            rec.update(y1=y1, a1=a1, y2=y2, a2=a2, ylast=y3, bnlast=bn3)
        res = x
        if b.downsample:
            bnd = bns[p + ".downsample.1"]
            yd = e._alloc16(N, ho, wo, b.cout)
            r = e._alloc16(N, ho, wo, b.cout)
            ds_implicit = (implicit and implicit_s2 and implicit_wgrad and b.stride == 2 and b.cin % 64 == 0 and b.cout % 64 == 0 and
                           e._wgrad_name == "dfd_gemm_wgrad")
            if b.stride == 1:
                xs = x
                fwd.append(gemm(_ptr(xs), P16(p + ".downsample.0.weight"), _ptr(yd), M2, b.cout, b.cin, bnd))
            elif ds_implicit:
                # strided 1x1 convolution straight from the block input (k = 1, stride 2 implicit GEMM): no gathered copy
                xs = None
                bnd.fused = fused_fin
                fwd.append(("dfd_conv_tc", (_ptr(x), P16(p + ".downsample.0.weight"), _ptr(yd), N, h, w, b.cin, b.cout, 1, b.stride, dt,
                                            bnd.fsum, bnd.fsq, bnd.fin if fused_fin else None)))
            else:
                xs = e._alloc16(N, ho, wo, b.cin)
                fwd.append(("dfd_im2col", (_ptr(x), _ptr(xs), N, h, w, b.cin, 1, b.stride, 0, dt)))
                fwd.append(gemm(_ptr(xs), P16(p + ".downsample.0.weight"), _ptr(yd), M2, b.cout, b.cin, bnd))
            fwd.append(finalize(bnd, M2))
            fwd.append(("dfd_bn_act", (_ptr(yd), bnd.scale, bnd.shift, None, None, _ptr(r), N, ho * wo, b.cout, ACT_NONE, 0, dt)))
            rec.update(yd=yd, xs=xs, bnd=bnd)
            res = r
        out = e._alloc16(N, ho, wo, b.cout)
        bl = rec["bnlast"]
        fwd.append(("dfd_bn_act", (_ptr(rec["ylast"]), bl.scale, bl.shift, None, _ptr(res), _ptr(out), N, ho * wo, b.cout,
                                   ACT_NONE, 2, dt)))
        e.acts[p + ".out"] = out
        rec["out"] = out
        recs.append(rec)
        x = out
    F, K = spec.num_features, spec.num_classes
    e.pooled = torch.zeros(N, F, dtype=torch.float32, device=dev)
    e.pool_partial = torch.zeros(8 * N * F, dtype=torch.float32, device=dev)
    fwd.append(("dfd_pool", (_ptr(x), None, None, _ptr(e.pooled), N, Hf * Wf, F, ACT_NONE, dt, _ptr(e.pool_partial), 8)))
    e.logits = torch.zeros(N, K, dtype=torch.float32, device=dev)
    e.dlogits = torch.zeros(N, K, dtype=torch.float32, device=dev)
    e.dpooled = torch.zeros(N, F, dtype=torch.float32, device=dev)
    e.target_i = torch.zeros(N, dtype=torch.int64, device=dev)
    e.target_f = torch.zeros(N, K, dtype=torch.float32, device=dev)

    # ---- backward ----------------------------------------------------------------------------------------
    pending_unpack = []         # (gperm region pointer, name, Cout, Cin): unpacked after the block's ordered reduce

    def zero_gperm(ptr, numel):
        return ("dfd_memset_async", (ptr, 0, numel * 4))

    def flush_block(ops):
        """ONE ordered reduce per block (every weight gradient of the block), then the packed 3x3 gradients -> OIHW arena"""
        e._flush_reduce(ops)
        for gp, name, Cout, Cin in pending_unpack:
            ops.append(("dfd_unpack_grad", (gp, G32(name), Cout, Cin, 3)))
        del pending_unpack[:]

    def conv3x3_bwd(name, dy, M_out, Cin, Cout, xin_t, n_h, n_w, stride, dx_out, dx_add=None):
        """dy [M_out, Cout] -> dx_out [N, n_h, n_w, Cin] (+dx_add) and the weight gradient of `name`"""
        if implicit and stride == 1 and dx_add is None and Cin % 64 == 0 and Cout % 64 == 0:
            # input gradient = the same implicit GEMM on dY with the tap-flipped [Cin][kh'][kw'][Cout] weights
            ops = [("dfd_conv_tc", (dy, PKD(name), dx_out, N, n_h, n_w, Cout, Cin, 3, 1, dt, None, None, None))]
        elif implicit and implicit_s2 and stride == 2 and dx_add is None and Cin % 64 == 0 and Cout % 64 == 0 and \
                not os.environ.get("DFD_NO_IMPLICIT_S2_DGRAD"):
            # strided input gradient: four parity-class implicit GEMMs storing through strided views of dx
            ops = [("dfd_conv_dgrad_s2_tc", (dy, PKD(name), dx_out, N, n_h, n_w, Cin, Cout, dt))]
        else:
            ops = [gemm(dy, PKT(name), COLS, M_out, 9 * Cin, Cout),
                   ("dfd_col2im", (COLS, dx_add, dx_out, N, n_h, n_w, Cin, 3, stride, 1, dt))]
        gp = _ptr(e.gperm, len(pending_unpack) * gperm_max)         # this block's next free region
        assert len(pending_unpack) < 2
        if implicit and implicit_wgrad and (stride == 1 or implicit_s2) and Cin % 64 == 0 and e._wgrad_name == "dfd_gemm_wgrad":
            ops += [zero_gperm(gp, Cout * 9 * Cin),
                    e._wgrad_conv(dy, _ptr(xin_t), gp, N, n_h, n_w, Cin, Cout, 3, stride)]
        else:
            ops += [("dfd_im2col", (_ptr(xin_t), COLS, N, n_h, n_w, Cin, 3, stride, 1, dt)),
                    zero_gperm(gp, Cout * 9 * Cin),
                    e._wgrad(dy, COLS, gp, M_out, Cout, 9 * Cin)]
        if os.environ.get("DFD_NONDET"):
            ops.append(("dfd_unpack_grad", (gp, G32(name), Cout, Cin, 3)))       # atomics: complete when the kernel is
        else:
            pending_unpack.append((gp, name, Cout, Cin))     # complete after the block's ordered reduce (flush_block)
        return ops

    bwd.append(("dfd_head_bwd", (_ptr(e.dlogits), _ptr(e.pooled), P32("fc.weight"), G32("fc.weight"), G32("fc.bias"),
                                 _ptr(e.dpooled), N, F, K)))
    bwd.append(("dfd_pool_bwd", (_ptr(e.dpooled), gA, N, Hf * Wf, F, dt)))
    # The gradient entering a block is kept as up to TWO tensors (main-path dx + identity-path gm of the block above): the
    # fused ReLU / BN-backward reduction adds them on the fly (dfd_relu_bn_bwd_reduce), which removes the materialised
    # residual add of every block without a downsample branch. DFD_NO_RELU_FUSE=1 restores the three separate passes.
    relu_fuse = not (fused_fin or os.environ.get("DFD_NO_RELU_FUSE"))
    bufs = [gA, gB, gC, gD, gE, gF]
    dout, dout2 = gA, None
    for rec in reversed(recs):
        b, h, w, ho, wo, xin = rec["b"], rec["h"], rec["w"], rec["ho"], rec["wo"], rec["x"]
        p = b.name
        M1, M2 = N * h * w, N * ho * wo
        gm, t1, t2, t3 = [g for g in bufs if g not in (dout, dout2)][:4]
        bl = rec["bnlast"]
        if not relu_fuse:
            bwd.append(("dfd_relu_bwd", (dout, _ptr(rec["out"]), gm, M2 * b.cout, dt)))
            bwd.append(("dfd_bn_bwd_reduce", (gm, _ptr(rec["ylast"]), None, bl.mean, bl.rstd, N, ho * wo, b.cout, dt, bl.bs1, bl.bs2, BF(bl))))
        else:
            # gm = (dout + dout2) * (out > 0) is produced by the reduction itself (one pass over the block output instead of
            # the residual add, the ReLU backward and the reduction)
            bwd.append(("dfd_relu_bn_bwd_reduce", (dout, dout2, _ptr(rec["ylast"]), _ptr(rec["out"]), gm, bl.mean, bl.rstd, N, ho * wo,
                                                   b.cout, dt, bl.bs1, bl.bs2)))
        bwd.append(bwd_finalize(bl, M2))
        bwd.append(("dfd_bn_bwd_apply", (gm, _ptr(rec["ylast"]), None, bl.cA, bl.cB, bl.cC, t1, N, ho * wo, b.cout, dt)))
        if b.kind == "basic":
            bn1 = bns[p + ".bn1"]
            # conv2 (3x3 s1): dy2 = t1 -> da1 = t2
            bwd += conv3x3_bwd(p + ".conv2.weight", t1, M2, b.planes, b.cout, rec["a1"], ho, wo, 1, t2)
            bwd.append(("dfd_act_bwd", (t2, _ptr(rec["y1"]), bn1.scale, bn1.shift, bn1.mean, bn1.rstd, None, None, t1, N,
                                        ho * wo, b.planes, ACT_RELU, dt, bn1.bs1, bn1.bs2, BF(bn1))))
            bwd.append(bwd_finalize(bn1, M2))
            bwd.append(("dfd_bn_bwd_apply", (t1, _ptr(rec["y1"]), None, bn1.cA, bn1.cB, bn1.cC, t2, N, ho * wo, b.planes, dt)))
            # conv1 (3x3 stride s): dy1 = t2 -> dx = t3 [M1, cin]
            bwd += conv3x3_bwd(p + ".conv1.weight", t2, M2, b.cin, b.planes, xin, h, w, b.stride, t3)
        else:
            bn1, bn2 = bns[p + ".bn1"], bns[p + ".bn2"]
            # conv3 (1x1): dy3 = t1 -> da2 = t2
            bwd.append(gemm(t1, T16(p + ".conv3.weight"), t2, M2, b.planes, b.cout))
            bwd.append(e._wgrad(t1, _ptr(rec["a2"]), G32(p + ".conv3.weight"), M2, b.cout, b.planes))
            bwd.append(("dfd_act_bwd", (t2, _ptr(rec["y2"]), bn2.scale, bn2.shift, bn2.mean, bn2.rstd, None, None, t1, N,
                                        ho * wo, b.planes, ACT_RELU, dt, bn2.bs1, bn2.bs2, BF(bn2))))
            bwd.append(bwd_finalize(bn2, M2))
            bwd.append(("dfd_bn_bwd_apply", (t1, _ptr(rec["y2"]), None, bn2.cA, bn2.cB, bn2.cC, t2, N, ho * wo, b.planes, dt)))
            # conv2 (3x3 stride s): dy2 = t2 -> da1 = t1 [M1, planes]
            bwd += conv3x3_bwd(p + ".conv2.weight", t2, M2, b.planes, b.planes, rec["a1"], h, w, b.stride, t1)
            bwd.append(("dfd_act_bwd", (t1, _ptr(rec["y1"]), bn1.scale, bn1.shift, bn1.mean, bn1.rstd, None, None, t2, N,
                                        h * w, b.planes, ACT_RELU, dt, bn1.bs1, bn1.bs2, BF(bn1))))
            bwd.append(bwd_finalize(bn1, M1))
            bwd.append(("dfd_bn_bwd_apply", (t2, _ptr(rec["y1"]), None, bn1.cA, bn1.cB, bn1.cC, t1, N, h * w, b.planes, dt)))
            # conv1 (1x1): dy1 = t1 -> dx = t3 [M1, cin]
            bwd.append(gemm(t1, T16(p + ".conv1.weight"), t3, M1, b.cin, b.planes))
            bwd.append(e._wgrad(t1, _ptr(xin), G32(p + ".conv1.weight"), M1, b.planes, b.cin))
        # identity / downsample path: gradient gm flows to the block input too
        if b.downsample:
            bnd = rec["bnd"]
            bwd.append(("dfd_bn_bwd_reduce", (gm, _ptr(rec["yd"]), None, bnd.mean, bnd.rstd, N, ho * wo, b.cout, dt, bnd.bs1, bnd.bs2, BF(bnd))))
            bwd.append(bwd_finalize(bnd, M2))
            bwd.append(("dfd_bn_bwd_apply", (gm, _ptr(rec["yd"]), None, bnd.cA, bnd.cB, bnd.cC, t1, N, ho * wo, b.cout, dt)))
            ds_add = (implicit and implicit_s2 and b.cin % 64 == 0 and b.cout % 64 == 0 and
                      not os.environ.get("DFD_NO_DGRAD_ADD"))
            if ds_add:
                # the downsample input gradient is ADDED into t3 (main-path gradient) by the GEMM's own epilogue: a TMA reduction
                # store through the stride-s pixel view of t3 - no scratch tensor, no col2im scatter / add pass
                bwd.append(("dfd_conv1x1_dgrad_add", (t1, T16(p + ".downsample.0.weight"), t3, N, h, w, b.cin, b.cout, b.stride, dt)))
            else:
                bwd.append(gemm(t1, T16(p + ".downsample.0.weight"), t2, M2, b.cin, b.cout))
            if rec["xs"] is None:      # strided 1x1: implicit weight gradient on the block input itself
                bwd.append(e._wgrad_conv(t1, _ptr(xin), G32(p + ".downsample.0.weight"), N, h, w, b.cin, b.cout, 1, b.stride))
            else:
                bwd.append(e._wgrad(t1, _ptr(rec["xs"]), G32(p + ".downsample.0.weight"), M2, b.cout, b.cin))
            if ds_add:
                new_dout = (t3, None)
            elif b.stride == 1:
                bwd.append(("dfd_add_inplace", (t3, t2, M1 * b.cin, dt)))
                new_dout = (t3, None)
            else:
                # scatter the strided gradient back onto the input grid and add the main-path gradient (the old dout buffer
                # has been consumed by the reduction above: reused as the destination)
                bwd.append(("dfd_col2im", (t2, t3, dout, N, h, w, b.cin, 1, b.stride, 0, dt)))
                new_dout = (dout, None)
        elif relu_fuse:
            new_dout = (t3, gm)             # the block below adds them while it masks and reduces
        else:
            bwd.append(("dfd_add_inplace", (t3, gm, M1 * b.cin, dt)))
            new_dout = (t3, None)
        flush_block(bwd)
        dout, dout2 = new_dout
    # stem: maxpool -> relu/bn1 -> conv1 wgrad
    if dout2 is not None:
        bwd.append(("dfd_add_inplace", (dout, dout2, N * H2 * W2 * 64, dt)))
    t1, t2 = [g for g in bufs if g != dout][:2]
    bwd.append(("dfd_maxpool_bwd", (dout, _ptr(e.pool_idx), t1, N, H1, W1, 64, dt)))
    bwd.append(("dfd_act_bwd", (t1, _ptr(y0), bn0.scale, bn0.shift, bn0.mean, bn0.rstd, None, None, t2, N, H1 * W1, 64,
                                ACT_RELU, dt, bn0.bs1, bn0.bs2, BF(bn0))))
    bwd.append(bwd_finalize(bn0, N * H1 * W1))
    if e.stem_impl == "gemm":
        bwd.append(("dfd_bn_bwd_apply", (t2, _ptr(y0), None, bn0.cA, bn0.cB, bn0.cC, t1, N, H1 * W1, 64, dt)))
        bwd.append(("dfd_memset_async", (_ptr(e.stem_gpad), 0, 64 * Kp * 4)))
        bwd.append(e._wgrad(t1, _ptr(e.stem_cols), _ptr(e.stem_gpad), N * H1 * W1, 64, Kp))
        e._flush_reduce(bwd)
        bwd.append(("dfd_unpad_grad", (_ptr(e.stem_gpad), G32("conv1.weight"), 64, taps, Kp)))
    else:
        bwd.append(("dfd_stem_wgrad", (_ptr(e.x_in), t2, _ptr(y0), bn0.cA, bn0.cB, bn0.cC, G32("conv1.weight"), N, spec.in_chans,
                                       e.H, e.W, 64, 7, 2, 3, dt)))

    bwd = e._patch_workspace([op for op in bwd if op is not None])
    e._upload_fin_descs()

    def base_name(n):
        return n[:-9] if n.endswith("_evalonly") else n

    for n, a in fwd + bwd:
        codes = _lib.SIGNATURES[base_name(n)]
        if len(a) != len(codes) - 1:
            raise AssertionError("%s: %d args for signature %r" % (n, len(a), codes))
    L = e.L
    e.fwd_ops = [(getattr(L, base_name(n)), n, a) for n, a in fwd]
    e.bwd_ops = [(getattr(L, n), n, tuple(a)) for n, a in bwd]
    e.n_launch["fwd"], e.n_launch["bwd"] = len(fwd), len(bwd)
