"""Flat-arena optimizers for the native engine (drop-in for the `torch.optim.Optimizer` protocol the runner uses).

Mirrors `dfd.timm.optim.create_optimizer` (dfd/timm/optim/optim_factory.py:26-100) for the optimizers the hot
path names — `sgd` (nesterov), `adam`, `adamw`, `rmsproptf` — including its parameter-group split
(`add_weight_decay`, optim_factory.py:11-23: 1-D tensors and biases get weight_decay 0; group order is
[no_decay, decay]) and the AdamW weight-decay rescale (optim_factory.py:29-33).  `param_groups[i]['lr']` is
re-read on every step because the schedulers mutate it (dfd/timm/scheduler/scheduler.py:81-85).

One kernel launch per parameter group updates fp32 master weights, optimizer state and the 16-bit copies the
conv kernels read; a second launch refreshes the transposed 1x1 weights.
"""
from collections import OrderedDict

import torch

from . import _lib
from .arch import is_no_decay
from .engine import _ptr


class ArenaOptimizer:
    def __init__(self, engine, opt="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4, eps=1e-8, betas=(0.9, 0.999),
                 alpha=0.9, filter_bias_and_bn=True):
        kind = opt.lower().split("_")[-1]
        if kind not in ("sgd", "adam", "adamw", "rmsproptf"):
            raise ValueError("optimizer %r is not on the native hot path (sgd, adam, adamw, rmsproptf)" % opt)
        self.engine = e = engine
        self.kind = kind
        wd = float(weight_decay)
        if kind == "adamw":
            wd /= lr  # optim_factory.py:29-33
        names_nd = [n for n in e.param_names if is_no_decay(n, e.p_off[n][1])]
        names_d = [n for n in e.param_names if not is_no_decay(n, e.p_off[n][1])]
        if not (wd and filter_bias_and_bn):
            raise ValueError("native optimizer expects weight_decay > 0 with the bias/BN filter (runner default)")
        base = dict(lr=float(lr), momentum=float(momentum), eps=float(eps), betas=tuple(betas), alpha=float(alpha))
        # ranges of the flat arena: decay tensors first, then no-decay tensors (engine._layout_params)
        self.param_groups = [
            dict(base, params=names_nd, weight_decay=0.0, _range=(e.n_decay, e.n_params)),
            dict(base, params=names_d, weight_decay=wd, _range=(0, e.n_decay)),
        ]
        dev = e.device
        n = e.n_params
        self.state_a = torch.zeros(n, dtype=torch.float32, device=dev)          # momentum / exp_avg / square_avg
        self.state_b = torch.zeros(n, dtype=torch.float32, device=dev) if kind != "sgd" else None
        if kind == "rmsproptf":
            self.state_a.fill_(1.0)                                            # rmsprop_tf.py:80
        self.step_count = 0
        self.grad_scale = 1.0          # 1/world (DDP mean) * 1/loss_scale
        self.skip_flag = None          # device int* (fp16 overflow)
        self.gscale_dev = None         # device float*: 1/loss_scale (fp16 dynamic loss scaling)
        e.n_launch["opt"] = 3

    def zero_grad(self, set_to_none=False):
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("dfd_memset_async", _ptr(self.engine.grads32), 0, self.engine.grads32.numel() * 4, st)

    def step(self, closure=None, stream=None):
        e = self.engine
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        self.step_count += 1
        for g in self.param_groups:
            lo, hi = g["_range"]
            n = hi - lo
            if n <= 0:
                continue
            p, gr, p16 = _ptr(e.params32, lo), _ptr(e.grads32, lo), _ptr(e.params16, lo)
            a = _ptr(self.state_a, lo)
            if self.kind == "sgd":
                _lib.call("dfd_sgd_step", p, gr, a, n, g["lr"], g["momentum"], g["weight_decay"], 1, self.grad_scale,
                          self.gscale_dev, self.skip_flag, p16, e.dt, st)
            elif self.kind in ("adam", "adamw"):
                _lib.call("dfd_adam_step", p, gr, a, _ptr(self.state_b, lo), n, g["lr"], g["betas"][0], g["betas"][1],
                          g["eps"], g["weight_decay"], 1 if self.kind == "adamw" else 0, self.step_count,
                          self.grad_scale, self.gscale_dev, self.skip_flag, p16, e.dt, st)
            else:
                _lib.call("dfd_rmsprop_tf_step", p, gr, a, _ptr(self.state_b, lo), n, g["lr"], g["alpha"], g["eps"],
                          g["weight_decay"], g["momentum"], self.grad_scale, self.gscale_dev, self.skip_flag, p16, e.dt, st)
        e.refresh_weight_layouts(st)

    # ---- torch-compatible (de)serialisation so `--resume` works across backends -----------------
    _KEYS = {"sgd": ("momentum_buffer", None), "adam": ("exp_avg", "exp_avg_sq"), "adamw": ("exp_avg", "exp_avg_sq"),
             "rmsproptf": ("square_avg", "momentum_buffer")}

    def state_dict(self):
        e = self.engine
        ka, kb = self._KEYS[self.kind]
        state = OrderedDict()
        groups = []
        idx = 0
        for g in self.param_groups:
            ids = []
            for name in g["params"]:
                o, s, n = e.p_off[name]
                st = {ka: self.state_a[o:o + n].view(s).clone()}
                if kb is not None:
                    st[kb] = self.state_b[o:o + n].view(s).clone()
                if self.kind != "sgd":
                    st["step"] = self.step_count
                state[idx] = st
                ids.append(idx)
                idx += 1
            groups.append({k: v for k, v in g.items() if k not in ("params", "_range")} | {"params": ids})
        return {"state": state, "param_groups": groups, "step_count": self.step_count}

    def load_state_dict(self, sd):
        e = self.engine
        ka, kb = self._KEYS[self.kind]
        idx = 0
        for g, gs in zip(self.param_groups, sd["param_groups"]):
            for k, v in gs.items():
                if k != "params":
                    g[k] = v
            for name in g["params"]:
                o, s, n = e.p_off[name]
                st = sd["state"].get(idx)
                if st is not None:
                    self.state_a[o:o + n].copy_(st[ka].reshape(-1))
                    if kb is not None and kb in st:
                        self.state_b[o:o + n].copy_(st[kb].reshape(-1))
                    if "step" in st:
                        self.step_count = int(st["step"])
                idx += 1
        self.step_count = int(sd.get("step_count", self.step_count))


def create_optimizer(args, model, filter_bias_and_bn=True):
    """Same signature as dfd.timm.optim.create_optimizer (optim_factory.py:26); `model` is a NativeModel/Engine."""
    engine = getattr(model, "engine", model)
    return ArenaOptimizer(engine, opt=args.opt, lr=args.lr, momentum=getattr(args, "momentum", 0.9),
                          weight_decay=args.weight_decay, eps=getattr(args, "opt_eps", 1e-8),
                          filter_bias_and_bn=filter_bias_and_bn)
