"""Architecture descriptions (host-side only) for the models on the hot path.

A *spec* is a plain description of the layer graph — channel counts, kernel sizes, strides and the
reference's parameter names — from which (a) the native engine lays out its HBM arenas and kernel
plan and (b) the CPU oracle (`oracle/model.py`) builds the same network out of torch fp32 ops.

This file restates the reference's arch-string decoder and stage builder for the variants that
BASELINE.json names; it builds no modules and owns no tensors.
  * EfficientNet arch strings / depth scaling: dfd/timm/models/efficientnet_builder.py:20-191
  * stage builder (stride only on the first block of a stage): efficientnet_builder.py:276-362
  * channel rounding: dfd/timm/models/efficientnet_blocks.py:55-69
  * SE width `make_divisible(block_in_chs * 0.25, 1)`: efficientnet_blocks.py:46-47,98
  * B0/B4 generator (stem 32, head 1280, 7 stages): dfd/timm/models/efficientnet.py:760-803,1078,1132
  * deepfake_v4 (stem 128, head 128, x2.0 / x3.1): efficientnet.py:806-851,1186-1192
  * ResNet-18/50 layout: dfd/timm/models/resnet.py:115-260,280-468,472,523
"""
import math
import re
from dataclasses import dataclass, field
from typing import List, Optional, Tuple


def make_divisible(v, divisor=8, min_value=None):
    # efficientnet_blocks.py:55-61
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def round_channels(channels, multiplier=1.0, divisor=8, channel_min=None):
    # efficientnet_blocks.py:64-69
    if not multiplier:
        return channels
    channels *= multiplier
    return make_divisible(channels, divisor, channel_min)


def conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


# ----------------------------------------------------------------------------------------------
# EfficientNet
# ----------------------------------------------------------------------------------------------

@dataclass
class MBBlock:
    """One DepthwiseSeparableConv ('ds') or InvertedResidual ('ir') block."""
    name: str            # 'blocks.<stage>.<idx>'
    kind: str            # 'ds' | 'ir'
    cin: int
    cmid: int            # == cin for 'ds'
    cout: int
    k: int               # depthwise kernel size
    stride: int
    cse: int             # squeeze width (0 = no SE)
    has_residual: bool

    @property
    def pad(self):
        # symmetric PyTorch padding, layers/padding.py:12-14 (pad_type='' never selects Conv2dSame)
        return (self.k - 1) // 2


@dataclass
class EfficientNetSpec:
    arch: str
    in_chans: int
    stem: int
    blocks: List[MBBlock]
    head_in: int
    num_features: int
    num_classes: int
    input_size: Tuple[int, int, int]
    family: str = "efficientnet"


_EFFNET_ARCH_DEF = [
    "ds_r1_k3_s1_e1_c16_se0.25",
    "ir_r2_k3_s2_e6_c24_se0.25",
    "ir_r2_k5_s2_e6_c40_se0.25",
    "ir_r3_k3_s2_e6_c80_se0.25",
    "ir_r3_k5_s1_e6_c112_se0.25",
    "ir_r4_k5_s2_e6_c192_se0.25",
    "ir_r1_k3_s1_e6_c320_se0.25",
]


def _decode_block_str(s):
    ops = s.split("_")
    kind = ops[0]
    opt = {}
    for op in ops[1:]:
        m = re.split(r"(\d.*)", op)
        if len(m) >= 2:
            opt[m[0]] = m[1]
    return dict(kind=kind, repeat=int(opt["r"]), k=int(opt["k"]), stride=int(opt["s"]),
                exp=float(opt.get("e", 1)), c=int(opt["c"]), se=float(opt["se"]) if "se" in opt else 0.0)


def _efficientnet_spec(arch, channel_multiplier, depth_multiplier, stem_size, num_features,
                       in_chans, num_classes, input_size):
    stem = round_channels(stem_size, channel_multiplier, 8, None)
    blocks = []
    cin = stem
    for si, bs in enumerate(_EFFNET_ARCH_DEF):
        d = _decode_block_str(bs)
        # one block string per stage -> _scale_stage_depth reduces to ceil(r * depth_multiplier)
        repeat = int(math.ceil(d["repeat"] * depth_multiplier))
        for bi in range(repeat):
            stride = d["stride"] if bi == 0 else 1
            cout = round_channels(d["c"], channel_multiplier, 8, None)
            if d["kind"] == "ds":
                cmid = cin
            else:
                cmid = make_divisible(cin * d["exp"])
            cse = make_divisible(cin * d["se"], 1) if d["se"] > 0 else 0
            blocks.append(MBBlock(name="blocks.%d.%d" % (si, bi), kind=d["kind"], cin=cin, cmid=cmid,
                                  cout=cout, k=d["k"], stride=stride, cse=cse,
                                  has_residual=(cin == cout and stride == 1)))
            cin = cout
    return EfficientNetSpec(arch=arch, in_chans=in_chans, stem=stem, blocks=blocks, head_in=cin,
                            num_features=num_features, num_classes=num_classes, input_size=input_size)


# ----------------------------------------------------------------------------------------------
# ResNet
# ----------------------------------------------------------------------------------------------

@dataclass
class ResBlock:
    name: str            # 'layer<l>.<idx>'
    kind: str            # 'basic' | 'bottleneck'
    cin: int
    planes: int
    cout: int
    stride: int
    downsample: bool     # 1x1 conv (stride) + BN on the identity path


@dataclass
class ResNetSpec:
    arch: str
    in_chans: int
    stem: int
    blocks: List[ResBlock]
    num_features: int
    num_classes: int
    input_size: Tuple[int, int, int]
    family: str = "resnet"


def _resnet_spec(arch, kind, layers, in_chans, num_classes, input_size):
    exp = 4 if kind == "bottleneck" else 1
    blocks = []
    cin = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
        for bi in range(n):
            stride = 2 if (bi == 0 and li > 0) else 1
            cout = planes * exp
            blocks.append(ResBlock(name="layer%d.%d" % (li + 1, bi), kind=kind, cin=cin, planes=planes,
                                   cout=cout, stride=stride,
                                   downsample=(bi == 0 and (stride != 1 or cin != cout))))
            cin = cout
    return ResNetSpec(arch=arch, in_chans=in_chans, stem=64, blocks=blocks, num_features=cin,
                      num_classes=num_classes, input_size=input_size)


# ----------------------------------------------------------------------------------------------
# registry of the variants on the hot path
# ----------------------------------------------------------------------------------------------

def get_spec(arch, num_classes=2, in_chans=3):
    if arch == "efficientnet_b0":
        return _efficientnet_spec(arch, 1.0, 1.0, 32, 1280, in_chans, num_classes, (3, 224, 224))
    if arch == "efficientnet_b4":
        return _efficientnet_spec(arch, 1.4, 1.8, 32, round_channels(1280, 1.4, 8, None), in_chans,
                                  num_classes, (3, 380, 380))
    if arch == "efficientnet_deepfake_v4":
        # efficientnet.py:806-851: stem_size=128, num_features=round_channels(128, 2.0)
        return _efficientnet_spec(arch, 2.0, 3.1, 128, round_channels(128, 2.0, 8, None), in_chans,
                                  num_classes, (in_chans, 600, 600))
    if arch == "resnet18":
        return _resnet_spec(arch, "basic", (2, 2, 2, 2), in_chans, num_classes, (3, 224, 224))
    if arch == "resnet50":
        return _resnet_spec(arch, "bottleneck", (3, 4, 6, 3), in_chans, num_classes, (3, 224, 224))
    raise ValueError("arch %r is not on the B200 hot path (see SURVEY.md section 8)" % (arch,))


SUPPORTED_ARCHS = ("efficientnet_b0", "efficientnet_b4", "efficientnet_deepfake_v4", "resnet18", "resnet50")


# ----------------------------------------------------------------------------------------------
# parameter / buffer naming in the reference's state_dict order
# ----------------------------------------------------------------------------------------------

def _bn_entries(prefix, c):
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
            (prefix + ".running_mean", (c,), "bn_rm"), (prefix + ".running_var", (c,), "bn_rv"),
            (prefix + ".num_batches_tracked", (), "bn_nbt")]


def state_entries(spec):
    """[(name, shape, role)] in the reference's `state_dict()` order (params and buffers interleaved).

    role in {'conv_w', 'dw_w', 'bn_w', 'bn_b', 'bn_rm', 'bn_rv', 'bn_nbt', 'se_w', 'se_b', 'fc_w', 'fc_b'}.
    """
    out = []
    if spec.family == "efficientnet":
        out.append(("conv_stem.weight", (spec.stem, spec.in_chans, 3, 3), "conv_w"))
        out += _bn_entries("bn1", spec.stem)
        for b in spec.blocks:
            p = b.name
            if b.kind == "ir":
                out.append((p + ".conv_pw.weight", (b.cmid, b.cin, 1, 1), "conv_w"))
                out += _bn_entries(p + ".bn1", b.cmid)
                out.append((p + ".conv_dw.weight", (b.cmid, 1, b.k, b.k), "dw_w"))
                out += _bn_entries(p + ".bn2", b.cmid)
            else:
                out.append((p + ".conv_dw.weight", (b.cmid, 1, b.k, b.k), "dw_w"))
                out += _bn_entries(p + ".bn1", b.cmid)
            if b.cse:
                out.append((p + ".se.conv_reduce.weight", (b.cse, b.cmid, 1, 1), "se_w"))
                out.append((p + ".se.conv_reduce.bias", (b.cse,), "se_b"))
                out.append((p + ".se.conv_expand.weight", (b.cmid, b.cse, 1, 1), "se_w"))
                out.append((p + ".se.conv_expand.bias", (b.cmid,), "se_b"))
            if b.kind == "ir":
                out.append((p + ".conv_pwl.weight", (b.cout, b.cmid, 1, 1), "conv_w"))
                out += _bn_entries(p + ".bn3", b.cout)
            else:
                out.append((p + ".conv_pw.weight", (b.cout, b.cmid, 1, 1), "conv_w"))
                out += _bn_entries(p + ".bn2", b.cout)
        out.append(("conv_head.weight", (spec.num_features, spec.head_in, 1, 1), "conv_w"))
        out += _bn_entries("bn2", spec.num_features)
        out.append(("classifier.weight", (spec.num_classes, spec.num_features), "fc_w"))
        out.append(("classifier.bias", (spec.num_classes,), "fc_b"))
    else:
        out.append(("conv1.weight", (64, spec.in_chans, 7, 7), "conv_w"))
        out += _bn_entries("bn1", 64)
        for b in spec.blocks:
            p = b.name
            if b.kind == "basic":
                out.append((p + ".conv1.weight", (b.planes, b.cin, 3, 3), "conv_w"))
                out += _bn_entries(p + ".bn1", b.planes)
                out.append((p + ".conv2.weight", (b.cout, b.planes, 3, 3), "conv_w"))
                out += _bn_entries(p + ".bn2", b.cout)
            else:
                out.append((p + ".conv1.weight", (b.planes, b.cin, 1, 1), "conv_w"))
                out += _bn_entries(p + ".bn1", b.planes)
                out.append((p + ".conv2.weight", (b.planes, b.planes, 3, 3), "conv_w"))
                out += _bn_entries(p + ".bn2", b.planes)
                out.append((p + ".conv3.weight", (b.cout, b.planes, 1, 1), "conv_w"))
                out += _bn_entries(p + ".bn3", b.cout)
            if b.downsample:
                out.append((p + ".downsample.0.weight", (b.cout, b.cin, 1, 1), "conv_w"))
                out += _bn_entries(p + ".downsample.1", b.cout)
        out.append(("fc.weight", (spec.num_classes, spec.num_features), "fc_w"))
        out.append(("fc.bias", (spec.num_classes,), "fc_b"))
    return out


def param_entries(spec):
    """[(name, shape, role)] of the learnable tensors, in `named_parameters()` order."""
    return [e for e in state_entries(spec) if e[2] not in ("bn_rm", "bn_rv", "bn_nbt")]


def is_no_decay(name, shape):
    # optim_factory.py:17: 1-D tensors and anything named *.bias get weight_decay = 0
    return len(shape) == 1 or name.endswith(".bias")
