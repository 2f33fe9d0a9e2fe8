"""CPU ORACLE — TEST INFRASTRUCTURE (see oracle/model.py header).

Restates one iteration of the reference's hot loop (dfd/runners/train.py:610-649) and the optimizers
the hot path names, as plain torch-fp32 CPU arithmetic:

  sgd_nesterov_step : torch.optim.SGD(nesterov=True) as built by dfd/timm/optim/optim_factory.py:48-50
  adam_step         : torch.optim.Adam (optim_factory.py:51-53; L2 decay added to the gradient)
  adamw_step        : dfd/timm/optim/adamw.py:55-117 (decoupled decay first, :72)
  rmsprop_tf_step   : dfd/timm/optim/rmsprop_tf.py:57-122 (square_avg init ONES :80, eps inside sqrt :107,
                      lr folded into the momentum buffer :112-114)
  param groups      : optim_factory.py:11-23 (1-D tensors and *.bias get weight_decay 0)
  train_step        : train.py:621-637 (forward, loss, prec1, zero_grad, backward, step)
  validate_step     : train.py:719-731
"""
import math

import torch

from deepfake_detection_b200.arch import is_no_decay, param_entries

from . import model as M


class OptState:
    """Optimizer state keyed by parameter name (momentum_buffer / exp_avg / exp_avg_sq / square_avg)."""

    def __init__(self, kind="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4, eps=1e-8, betas=(0.9, 0.999), alpha=0.9):
        self.kind = kind
        self.lr = lr
        self.momentum = momentum
        self.weight_decay = weight_decay
        self.eps = eps
        self.betas = betas
        self.alpha = alpha
        self.step = 0
        self.state = {}


def _wd_for(opt, name, p):
    # for adamw the caller passes weight_decay already divided by lr (optim_factory.py:29-33)
    return 0.0 if is_no_decay(name, tuple(p.shape)) else opt.weight_decay


@torch.no_grad()
def optimizer_step(opt, params, grads):
    """params/grads: dict name -> tensor. Updates params in place."""
    opt.step += 1
    for name, p in params.items():
        g = grads[name]
        wd = _wd_for(opt, name, p)
        st = opt.state.setdefault(name, {})
        if opt.kind == "sgd":
            if wd != 0:
                g = g.add(p, alpha=wd)
            if "momentum_buffer" not in st:
                buf = st["momentum_buffer"] = g.clone()
            else:
                buf = st["momentum_buffer"]
                buf.mul_(opt.momentum).add_(g)
            g = g.add(buf, alpha=opt.momentum)  # nesterov
            p.add_(g, alpha=-opt.lr)
        elif opt.kind in ("adam", "adamw"):
            b1, b2 = opt.betas
            if opt.kind == "adamw":
                p.mul_(1 - opt.lr * wd)
            elif wd != 0:
                g = g.add(p, alpha=wd)
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
            st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1 = 1 - b1 ** opt.step
            bc2 = 1 - b2 ** opt.step
            denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(opt.eps)
            p.addcdiv_(st["exp_avg"], denom, value=-opt.lr / bc1)
        elif opt.kind == "rmsproptf":
            if wd != 0:
                g = g.add(p, alpha=wd)
            if "square_avg" not in st:
                st["square_avg"] = torch.ones_like(p)
                st["momentum_buffer"] = torch.zeros_like(p)
            sq = st["square_avg"]
            sq.add_(g.pow(2) - sq, alpha=1 - opt.alpha)
            avg = sq.add(opt.eps).sqrt_()
            if opt.momentum > 0:
                buf = st["momentum_buffer"]
                buf.mul_(opt.momentum).addcdiv_(g, avg, value=opt.lr)
                p.add_(-buf)
            else:
                p.addcdiv_(g, avg, value=-opt.lr)
        else:
            raise ValueError(opt.kind)


def split_state(spec, sd):
    """-> (params dict in named_parameters order, buffers dict)"""
    pnames = [n for n, _, _ in param_entries(spec)]
    params = {n: sd[n] for n in pnames}
    buffers = {n: t for n, t in sd.items() if n not in params}
    return params, buffers


def train_step(spec, sd, x, target, opt=None, smoothing=0.0, bn=None, act_dtype=None, taps=None,
               grad_hook=None, drop_masks=None, dropout_mask=None):
    """One iteration of train.py:621-637 on CPU. `sd` tensors are updated in place.
    Returns dict(logits, loss, prec1, grads)."""
    params, _ = split_state(spec, sd)
    for p in params.values():
        p.requires_grad_(True)
        p.grad = None
    logits = M.forward(spec, sd, x, bn or M.BNState(training=True), act_dtype, taps, drop_masks, dropout_mask)
    loss = M.cross_entropy(logits, target, smoothing)
    prec1 = M.accuracy_top1(logits.detach(), target)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in params.items()}
    for p in params.values():
        p.requires_grad_(False)
        p.grad = None
    if grad_hook is not None:
        grad_hook(grads)  # e.g. the DDP mean all-reduce
    if opt is not None:
        optimizer_step(opt, params, grads)
    return dict(logits=logits.detach(), loss=loss.detach(), prec1=prec1, grads=grads)


@torch.no_grad()
def validate_step(spec, sd, x, target, bn_eps=1e-5, act_dtype=None):
    logits = M.forward(spec, sd, x, M.BNState(training=False, eps=bn_eps), act_dtype)
    return dict(logits=logits, loss=M.cross_entropy(logits, target, 0.0), prec1=M.accuracy_top1(logits, target))
