"""LIBRARY BASELINE — measurement infrastructure, not product code (only `bench.py --impl library` imports it).

The same layer graphs as the native path (built from deepfake_detection_b200.arch specs, i.e. the reference's
dfd/timm EfficientNet / ResNet module trees: efficientnet.py:320-348, efficientnet_blocks.py:104-110,177-194,314-348,
resnet.py:150-175,215-246,450-468) as STOCK torch.nn modules, to be run in PyTorch eager mode with bf16/fp16 autocast,
channels_last, cuDNN / cuBLAS kernels and torch DistributedDataParallel over NCCL: the "library" bar SURVEY.md 8(d) names.
nn.SiLU stands in for the reference's Swish (favourable to the library arm: one fused kernel instead of mul + sigmoid with a
custom autograd function, layers/activations.py:19-48).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SE(nn.Module):
    def __init__(self, c, cse):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, cse, 1, bias=True)
        self.conv_expand = nn.Conv2d(cse, c, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(F.silu(self.conv_reduce(s)))
        return x * torch.sigmoid(s)


class MB(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.b = b
        pad = (b.k - 1) // 2
        if b.kind == "ir":
            self.conv_pw = nn.Conv2d(b.cin, b.cmid, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(b.cmid)
            self.conv_dw = nn.Conv2d(b.cmid, b.cmid, b.k, b.stride, pad, groups=b.cmid, bias=False)
            self.bn2 = nn.BatchNorm2d(b.cmid)
            self.se = SE(b.cmid, b.cse) if b.cse else None
            self.conv_pwl = nn.Conv2d(b.cmid, b.cout, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(b.cout)
        else:
            self.conv_dw = nn.Conv2d(b.cmid, b.cmid, b.k, b.stride, pad, groups=b.cmid, bias=False)
            self.bn1 = nn.BatchNorm2d(b.cmid)
            self.se = SE(b.cmid, b.cse) if b.cse else None
            self.conv_pw = nn.Conv2d(b.cmid, b.cout, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(b.cout)

    def forward(self, x):
        r = x
        if self.b.kind == "ir":
            x = F.silu(self.bn1(self.conv_pw(x)))
            x = F.silu(self.bn2(self.conv_dw(x)))
            if self.se is not None:
                x = self.se(x)
            x = self.bn3(self.conv_pwl(x))
        else:
            x = F.silu(self.bn1(self.conv_dw(x)))
            if self.se is not None:
                x = self.se(x)
            x = self.bn2(self.conv_pw(x))
        return x + r if self.b.has_residual else x


class EfficientNet(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.conv_stem = nn.Conv2d(spec.in_chans, spec.stem, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(spec.stem)
        self.blocks = nn.Sequential(*[MB(b) for b in spec.blocks])
        self.conv_head = nn.Conv2d(spec.head_in, spec.num_features, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(spec.num_features)
        self.classifier = nn.Linear(spec.num_features, spec.num_classes)

    def forward(self, x):
        x = F.silu(self.bn1(self.conv_stem(x)))
        x = self.blocks(x)
        x = F.silu(self.bn2(self.conv_head(x)))
        return self.classifier(x.mean((2, 3)))


class Res(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.b = b
        if b.kind == "basic":
            self.conv1 = nn.Conv2d(b.cin, b.planes, 3, b.stride, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(b.planes)
            self.conv2 = nn.Conv2d(b.planes, b.cout, 3, 1, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(b.cout)
        else:
            self.conv1 = nn.Conv2d(b.cin, b.planes, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(b.planes)
            self.conv2 = nn.Conv2d(b.planes, b.planes, 3, b.stride, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(b.planes)
            self.conv3 = nn.Conv2d(b.planes, b.cout, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(b.cout)
        self.downsample = nn.Sequential(nn.Conv2d(b.cin, b.cout, 1, b.stride, bias=False), nn.BatchNorm2d(b.cout)) if b.downsample else None

    def forward(self, x):
        r = x if self.downsample is None else self.downsample(x)
        if self.b.kind == "basic":
            x = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        else:
            x = F.relu(self.bn1(self.conv1(x)))
            x = F.relu(self.bn2(self.conv2(x)))
            x = self.bn3(self.conv3(x))
        return F.relu(x + r)


class ResNet(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.conv1 = nn.Conv2d(spec.in_chans, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layers = nn.Sequential(*[Res(b) for b in spec.blocks])
        self.fc = nn.Linear(spec.num_features, spec.num_classes)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        x = self.layers(x)
        return self.fc(x.mean((2, 3)))


def build(spec):
    return EfficientNet(spec) if spec.family == "efficientnet" else ResNet(spec)
