"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-fp32 CPU restatement of the reference's train/validate step for the models on the hot
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker / the CPU baseline — never as the thing shipped.
The product path (`deepfake_detection_b200`) must not import it (tests/test_no_oracle_on_product_path.py).

Parity pin: the reference has NO tests, golden vectors or fixtures of its own, so
this restatement is pinned against outputs of the reference itself: `oracle/mint_goldens.py` imports the
unmodified reference from /root/reference (authoring container only), runs it on seeded synthetic weights
and inputs, and commits the results under tests/golden/; tests/test_oracle_vs_reference_goldens.py checks
this file against them on CPU.

What each function follows (file:line under /root/reference):
  efficientnet_forward : dfd/timm/models/efficientnet.py:320-348
  _mb_block            : dfd/timm/models/efficientnet_blocks.py:177-194 (DS), :314-348 (IR)
  _squeeze_excite      : dfd/timm/models/efficientnet_blocks.py:104-110
  swish                : dfd/timm/models/layers/activations.py:19-27 (x * sigmoid(x))
  resnet_forward       : dfd/timm/models/resnet.py:450-468, BasicBlock :150-175, Bottleneck :215-246
  batch norm           : torch.nn.BatchNorm2d semantics (biased var to normalise, unbiased var into
                         running_var, momentum 0.1, eps 1e-5 unless overridden, efficientnet_blocks.py:22-30)
  cross_entropy        : dfd/timm/loss/cross_entropy.py:20-36 and nn.CrossEntropyLoss (train.py:509-520)
  accuracy             : dfd/timm/utils.py:170-186

`act_dtype` (None | torch.bfloat16 | torch.float16): when set, every tensor that the native path stores
in 16-bit (conv outputs, block outputs, SE-gated activations, the network input) is rounded to that type
and back at the same point.  Arithmetic stays fp32.  This keeps the comparison with the CUDA path tight
enough to localise a wrong kernel; parity against the *reference* uses act_dtype=None.
"""
import torch
import torch.nn.functional as F


def _q(x, act_dtype):
    if act_dtype is None:
        return x
    return x.to(act_dtype).to(torch.float32)


class _QuantSTE(torch.autograd.Function):
    """Round-trip through a 16-bit type in forward AND backward (the native path stores activation
    gradients in 16-bit too)."""

    @staticmethod
    def forward(ctx, x, dt):
        ctx.dt = dt
        return x.to(dt).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dt).to(torch.float32), None


def q(x, act_dtype, grad_too=True):
    if act_dtype is None:
        return x
    if grad_too:
        return _QuantSTE.apply(x, act_dtype)
    return x.to(act_dtype).to(torch.float32)


def swish(x):
    return x * torch.sigmoid(x)


class BNState:
    """Holds bn hyper-parameters and whether running stats are updated (training) or used (eval)."""

    def __init__(self, training=True, momentum=0.1, eps=1e-5):
        self.training = training
        self.momentum = momentum
        self.eps = eps


def batch_norm(x, sd, prefix, bn):
    rm = sd[prefix + ".running_mean"]
    rv = sd[prefix + ".running_var"]
    if bn.training:
        nbt = sd.get(prefix + ".num_batches_tracked")
        if nbt is not None:
            nbt += 1
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], bn.training, bn.momentum, bn.eps)


def _squeeze_excite(x, sd, p):
    x_se = x.mean((2, 3), keepdim=True)
    x_se = F.conv2d(x_se, sd[p + ".se.conv_reduce.weight"], sd[p + ".se.conv_reduce.bias"])
    x_se = swish(x_se)
    x_se = F.conv2d(x_se, sd[p + ".se.conv_expand.weight"], sd[p + ".se.conv_expand.bias"])
    return x * torch.sigmoid(x_se)


def _mb_block(x, sd, b, bn, act_dtype, taps, drop_mask=None):
    """drop_mask: optional [N] tensor = floor(keep + U) / keep of drop_path (layers/drop.py:84-100), applied to the block's
    main path before the residual add (efficientnet_blocks.py:343-346); injected by the tests (torch's own random stream is
    not reproducible across implementations)."""
    p = b.name
    residual = x
    if b.kind == "ir":
        x = q(F.conv2d(x, sd[p + ".conv_pw.weight"]), act_dtype)
        if taps is not None:
            taps[p + ".conv_pw"] = x
        # the native depthwise kernel stages its activated input in 16-bit shared memory (as apex AMP O1 casts
        # the conv input to half in the reference's GPU path): one more rounding point in emulation mode
        x = q(swish(batch_norm(x, sd, p + ".bn1", bn)), act_dtype)
        x = q(F.conv2d(x, sd[p + ".conv_dw.weight"], stride=b.stride, padding=b.pad, groups=b.cmid), act_dtype)
        if taps is not None:
            taps[p + ".conv_dw"] = x
        x = swish(batch_norm(x, sd, p + ".bn2", bn))
        if b.cse:
            x = _squeeze_excite(x, sd, p)
        x = q(x, act_dtype)
        x = q(F.conv2d(x, sd[p + ".conv_pwl.weight"]), act_dtype)
        if taps is not None:
            taps[p + ".conv_pwl"] = x
        x = batch_norm(x, sd, p + ".bn3", bn)
    else:
        x = q(F.conv2d(x, sd[p + ".conv_dw.weight"], stride=b.stride, padding=b.pad, groups=b.cmid), act_dtype)
        if taps is not None:
            taps[p + ".conv_dw"] = x
        x = swish(batch_norm(x, sd, p + ".bn1", bn))
        if b.cse:
            x = _squeeze_excite(x, sd, p)
        x = q(x, act_dtype)
        x = q(F.conv2d(x, sd[p + ".conv_pw.weight"]), act_dtype)
        if taps is not None:
            taps[p + ".conv_pw"] = x
        x = batch_norm(x, sd, p + ".bn2", bn)
    if b.has_residual:
        if drop_mask is not None and bn.training:
            x = x * drop_mask.view(-1, 1, 1, 1)
        x = x + residual
    x = q(x, act_dtype)
    if taps is not None:
        taps[p + ".out"] = x
    return x


def efficientnet_forward(spec, sd, x, bn=None, act_dtype=None, taps=None, drop_masks=None, dropout_mask=None):
    """x: [N, C, H, W] fp32 -> logits [N, num_classes] fp32. `sd` maps reference state_dict names to
    fp32 CPU tensors (parameters may require grad; running stats are updated in place when bn.training)."""
    bn = bn or BNState()
    x = q(x, act_dtype, grad_too=False)
    x = q(F.conv2d(x, sd["conv_stem.weight"], stride=2, padding=1), act_dtype)
    if taps is not None:
        taps["conv_stem"] = x
    x = q(swish(batch_norm(x, sd, "bn1", bn)), act_dtype)
    if taps is not None:
        taps["stem.out"] = x
    for b in spec.blocks:
        x = _mb_block(x, sd, b, bn, act_dtype, taps, None if drop_masks is None else drop_masks.get(b.name))
    x = q(F.conv2d(x, sd["conv_head.weight"]), act_dtype)
    if taps is not None:
        taps["conv_head"] = x
    x = swish(batch_norm(x, sd, "bn2", bn))
    x = x.mean((2, 3))
    if taps is not None:
        taps["pooled"] = x
    if dropout_mask is not None and bn.training:
        x = x * dropout_mask            # F.dropout(x, p, training) with the mask (already / keep) injected, efficientnet.py:346-347
    return F.linear(x, sd["classifier.weight"], sd["classifier.bias"])


def _res_block(x, sd, b, bn, act_dtype, taps):
    p = b.name
    residual = x
    if b.kind == "basic":
        x = q(F.conv2d(x, sd[p + ".conv1.weight"], stride=b.stride, padding=1), act_dtype)
        x = q(F.relu(batch_norm(x, sd, p + ".bn1", bn)), act_dtype)
        x = q(F.conv2d(x, sd[p + ".conv2.weight"], padding=1), act_dtype)
        x = batch_norm(x, sd, p + ".bn2", bn)
    else:
        x = q(F.conv2d(x, sd[p + ".conv1.weight"]), act_dtype)
        x = q(F.relu(batch_norm(x, sd, p + ".bn1", bn)), act_dtype)
        x = q(F.conv2d(x, sd[p + ".conv2.weight"], stride=b.stride, padding=1), act_dtype)
        x = q(F.relu(batch_norm(x, sd, p + ".bn2", bn)), act_dtype)
        x = q(F.conv2d(x, sd[p + ".conv3.weight"]), act_dtype)
        x = batch_norm(x, sd, p + ".bn3", bn)
    if b.downsample:
        residual = q(F.conv2d(residual, sd[p + ".downsample.0.weight"], stride=b.stride), act_dtype)
        residual = batch_norm(residual, sd, p + ".downsample.1", bn)
    x = q(F.relu(x + residual), act_dtype)
    if taps is not None:
        taps[p + ".out"] = x
    return x


def resnet_forward(spec, sd, x, bn=None, act_dtype=None, taps=None):
    bn = bn or BNState()
    x = q(x, act_dtype, grad_too=False)
    x = q(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), act_dtype)
    x = q(F.relu(batch_norm(x, sd, "bn1", bn)), act_dtype)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["stem.out"] = x
    for b in spec.blocks:
        x = _res_block(x, sd, b, bn, act_dtype, taps)
    x = x.mean((2, 3))
    if taps is not None:
        taps["pooled"] = x
    return F.linear(x, sd["fc.weight"], sd["fc.bias"])


def forward(spec, sd, x, bn=None, act_dtype=None, taps=None, drop_masks=None, dropout_mask=None):
    if spec.family == "efficientnet":
        return efficientnet_forward(spec, sd, x, bn, act_dtype, taps, drop_masks, dropout_mask)
    return resnet_forward(spec, sd, x, bn, act_dtype, taps)


# ------------------------------------------------------------------------------------------------
# losses and metrics
# ------------------------------------------------------------------------------------------------

def cross_entropy(logits, target, smoothing=0.0):
    """Covers the reference's three losses: nn.CrossEntropyLoss (hard target, smoothing 0),
    LabelSmoothingCrossEntropy (cross_entropy.py:20-26) and SoftTargetCrossEntropy (:34-36, float target)."""
    logp = F.log_softmax(logits, dim=-1)
    if target.dtype.is_floating_point:
        return torch.sum(-target * logp, dim=-1).mean()
    nll = -logp.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
    smooth = -logp.mean(dim=-1)
    return ((1.0 - smoothing) * nll + smoothing * smooth).mean()


def bce_two_class(logits, target, smoothing=0.0):
    """The sigmoid-BCE form the fused CUDA head computes:
    with d = z1 - z0 and soft target t1, CE(z, t) == t1*softplus(-d) + (1-t1)*softplus(d)."""
    d = logits[:, 1] - logits[:, 0]
    if target.dtype.is_floating_point:
        t1 = target[:, 1]
    else:
        t1 = target.to(torch.float32) * (1.0 - smoothing) + 0.5 * smoothing
    return (t1 * F.softplus(-d) + (1.0 - t1) * F.softplus(d)).mean()


def accuracy_top1(logits, target):
    # utils.py:170-186 (soft targets: argmax of the target row, :177-178)
    if target.dtype.is_floating_point:
        target = target.argmax(dim=1)
    pred = logits.argmax(dim=1)
    return (pred == target).to(torch.float32).mean() * 100.0
