"""Flat-arena optimizers for the native engine (drop-in for the `torch.optim.Optimizer` protocol the runner uses).

Mirrors `dfd.timm.optim.create_optimizer` (dfd/timm/optim/optim_factory.py:26-100) for the optimizers the hot
path names — `sgd` (always nesterov, optim_factory.py:48-50), `adam`, `adamw`, `rmsproptf` —
including its parameter-group split (`add_weight_decay`, optim_factory.py:11-23: 1-D tensors and biases get
weight_decay 0; group order is [no_decay, decay]; with `weight_decay == 0` or `filter_bias_and_bn=False` the
reference passes ONE group of `model.parameters()`, optim_factory.py:34-38, and so does this class) and the AdamW
weight-decay rescale (optim_factory.py:29-33).  `lookahead_*` (optim_factory.py:96-98) is not on the native path
and is rejected by name instead of being silently reduced to its base optimizer.

`param_groups[i]['lr']` is re-read on every step because the schedulers mutate it
(dfd/timm/scheduler/scheduler.py:81-85).  The value travels to the device in a 1-block launch (`dfd_set_floats`,
`push_hyper`) and the update kernels read it from device memory, so a CUDA graph captured around `step()` stays
valid across every scheduler update; Adam's step count lives on the device too (`dfd_opt_tick`) and does not
advance on an fp16-overflow-skipped step (apex semantics).

One kernel launch per arena range updates fp32 master weights, optimizer state and the 16-bit copies the conv
kernels read; further launches refresh the derived weight layouts.
"""
from collections import OrderedDict

import torch

from . import _lib
from .arch import is_no_decay
from .engine import _ptr

_KINDS = {"sgd": ("sgd", 1), "adam": ("adam", 0), "adamw": ("adamw", 0), "rmsproptf": ("rmsproptf", 0)}


class ArenaOptimizer:
    def __init__(self, engine, opt="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4, eps=1e-8, betas=(0.9, 0.999),
                 alpha=0.9, filter_bias_and_bn=True):
        opt_lower = opt.lower()
        opt_split = opt_lower.split("_")
        if len(opt_split) > 1:
            # optim_factory.py:96-98 wraps the base optimizer in Lookahead for `lookahead_<name>`: a different algorithm
            raise ValueError("optimizer %r: Lookahead wrappers are not on the native hot path "
                             "(supported: sgd, adam, adamw, rmsproptf)" % opt)
        if opt_lower not in _KINDS:
            raise ValueError("optimizer %r is not on the native hot path (sgd, adam, adamw, rmsproptf)" % opt)
        self.engine = e = getattr(engine, "arena", engine)      # the owner of the parameter / gradient arenas
        self.kind, self.nesterov = _KINDS[opt_lower]
        wd = float(weight_decay)
        if self.kind == "adamw" and wd and lr:
            wd /= lr  # optim_factory.py:29-33
        base = dict(lr=float(lr), momentum=float(momentum), eps=float(eps), betas=tuple(betas), alpha=float(alpha))
        if wd and filter_bias_and_bn:
            # ranges of the flat arena: decay tensors first, then no-decay tensors (engine._layout_params)
            names_nd = [n for n in e.param_names if is_no_decay(n, e.p_off[n][1])]
            names_d = [n for n in e.param_names if not is_no_decay(n, e.p_off[n][1])]
            self.param_groups = [
                dict(base, params=names_nd, weight_decay=0.0, _ranges=[(e.n_decay, e.n_params)]),
                dict(base, params=names_d, weight_decay=wd, _ranges=[(0, e.n_decay)]),
            ]
        else:
            # optim_factory.py:34-38: a single group over model.parameters() (named_parameters order)
            self.param_groups = [dict(base, params=list(e.param_names), weight_decay=wd,
                                      _ranges=[(0, e.n_decay), (e.n_decay, e.n_params)])]
        dev = e.device
        n = e.n_params
        self.state_a = torch.zeros(n, dtype=torch.float32, device=dev)          # momentum / exp_avg / square_avg
        self.state_b = torch.zeros(n, dtype=torch.float32, device=dev) if self.kind != "sgd" else None
        if self.kind == "rmsproptf":
            self.state_a.fill_(1.0)                                            # rmsprop_tf.py:80
        self.grad_scale = 1.0          # extra factor on the gradients (the DDP mean is taken by the reducer itself)
        self.skip_flag = None          # device int* (fp16 overflow)
        self.gscale_dev = None         # device float*: 1/loss_scale (fp16 dynamic loss scaling)
        # device-resident hyper-parameters: lr of every group, and the step counter of Adam's bias corrections
        self.hyper_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._host_steps = 0           # launches of step(); the device counter is the authority (skips do not advance it)
        e.n_launch["opt"] = 3

    # `step_count` mirrors the device counter (reads synchronise; used by state_dict / tests, not on the hot path)
    @property
    def step_count(self):
        if self.engine._plan_only:
            return self._host_steps
        return int(self.step_dev.item())

    @step_count.setter
    def step_count(self, v):
        self._host_steps = int(v)
        if not self.engine._plan_only:
            self.step_dev.fill_(int(v))

    def zero_grad(self, set_to_none=False):
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("dfd_memset_async", _ptr(self.engine.grads32), 0, self.engine.grads32.numel() * 4, st)

    def hyper_signature(self):
        """everything EXCEPT lr that the captured launches bake in (a change re-captures the graph)"""
        return (self.kind, self.nesterov, self.grad_scale,
                tuple((g["momentum"], g["weight_decay"], g["eps"], g["betas"], g["alpha"]) for g in self.param_groups))

    def push_hyper(self, stream=None):
        """current param_groups[i]['lr'] -> device (one tiny launch; call it OUTSIDE a captured graph, before replay)"""
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        lrs = [float(g["lr"]) for g in self.param_groups]
        if len(lrs) > 8:
            raise ValueError("at most 8 parameter groups")
        lrs += [0.0] * (8 - len(lrs))
        _lib.call("dfd_set_floats", _ptr(self.hyper_dev), len(self.param_groups), *lrs, st)

    def step(self, closure=None, stream=None, push=True):
        e = self.engine
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        if push:
            self.push_hyper(st)
        self._host_steps += 1
        if self.kind in ("adam", "adamw"):
            _lib.call("dfd_opt_tick", _ptr(self.step_dev), self.skip_flag, st)
        for gi, g in enumerate(self.param_groups):
            lr_dev = _ptr(self.hyper_dev, gi)
            for lo, hi in g["_ranges"]:
                n = hi - lo
                if n <= 0:
                    continue
                p, gr, p16 = _ptr(e.params32, lo), _ptr(e.grads32, lo), _ptr(e.params16, lo)
                a = _ptr(self.state_a, lo)
                if self.kind == "sgd":
                    _lib.call("dfd_sgd_step", p, gr, a, n, g["lr"], g["momentum"], g["weight_decay"], self.nesterov,
                              self.grad_scale, self.gscale_dev, self.skip_flag, p16, e.dt, lr_dev, st)
                elif self.kind in ("adam", "adamw"):
                    _lib.call("dfd_adam_step", p, gr, a, _ptr(self.state_b, lo), n, g["lr"], g["betas"][0], g["betas"][1],
                              g["eps"], g["weight_decay"], 1 if self.kind == "adamw" else 0, self._host_steps,
                              self.grad_scale, self.gscale_dev, self.skip_flag, p16, e.dt, lr_dev, _ptr(self.step_dev), st)
                else:
                    _lib.call("dfd_rmsprop_tf_step", p, gr, a, _ptr(self.state_b, lo), n, g["lr"], g["alpha"], g["eps"],
                              g["weight_decay"], g["momentum"], self.grad_scale, self.gscale_dev, self.skip_flag, p16, e.dt,
                              lr_dev, st)
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
        steps = self.step_count
        for g in self.param_groups:
            ids = []
            for name in g["params"]:
                o, s, n = e.p_off[name]
                st = {ka: self.state_a[o:o + n].view(s).clone()}
                if kb is not None:
                    st[kb] = self.state_b[o:o + n].view(s).clone()
                if self.kind != "sgd":
                    st["step"] = steps
                state[idx] = st
                ids.append(idx)
                idx += 1
            groups.append({k: v for k, v in g.items() if k not in ("params", "_ranges")} | {"params": ids})
        return {"state": state, "param_groups": groups, "step_count": steps}

    def load_state_dict(self, sd):
        e = self.engine
        ka, kb = self._KEYS[self.kind]
        idx = 0
        steps = None
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
                        steps = int(st["step"])
                idx += 1
        steps = sd.get("step_count", steps)
        if steps is not None:
            self.step_count = int(steps)


def create_optimizer(args, model, filter_bias_and_bn=True):
    """Same signature as dfd.timm.optim.create_optimizer (optim_factory.py:26); `model` is a NativeModel / NativeDDP /
    Engine."""
    model = getattr(model, "module", model)
    engine = getattr(model, "engine", model)
    return ArenaOptimizer(engine, opt=args.opt, lr=args.lr, momentum=getattr(args, "momentum", 0.9),
                          weight_decay=args.weight_decay, eps=getattr(args, "opt_eps", 1e-8),
                          filter_bias_and_bn=filter_bias_and_bn)
