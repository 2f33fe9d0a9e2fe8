"""Mint the committed golden fixtures under tests/golden/ FROM THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Run in the authoring container only (needs /root/reference, read-only):

    python -m oracle.mint_goldens

It imports the reference through oracle/ref_shims.py, loads the formula-generated synthetic weights of
oracle/weights.py into the reference's own modules (`dfd.timm.models.create_model`), drives the
reference's own optimizers (`dfd.timm.optim.create_optimizer`), losses and — for BASELINE config 1 — its
own `dfd.runners.train.train_epoch` / `validate`, and stores the results.  The fixtures are small
(norms + sampled elements, not full tensors) and are what pins oracle/ to the reference.
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from deepfake_detection_b200.arch import get_spec, param_entries, state_entries  # noqa: E402
from oracle import ref_shims  # noqa: E402
from oracle.weights import synth_batch, synth_state  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (arch, batch, H, W): small resolutions so the CPU suite stays in minutes; 76 exercises odd sizes
STEP_CASES = [
    ("efficientnet_b0", 4, 64, 64),
    ("efficientnet_b4", 2, 76, 76),
    ("resnet18", 4, 64, 64),
    ("resnet50", 2, 64, 64),
]


def _summ(t):
    """norm + first/last few elements: enough to catch any semantic slip, tiny on disk"""
    f = t.detach().reshape(-1).to(torch.float64)
    idx = torch.linspace(0, f.numel() - 1, steps=min(8, f.numel())).round().long()
    return dict(norm=float(f.norm()), sum=float(f.sum()), samples=f[idx].to(torch.float32).tolist(),
                idx=idx.tolist())


def _args(**kw):
    d = dict(opt="sgd", lr=0.05, momentum=0.9, weight_decay=1e-4, opt_eps=1e-8)
    d.update(kw)
    return SimpleNamespace(**d)


def mint_state_keys():
    from dfd.timm.models import create_model
    from dfd.timm.models.factory import create_deepfake_model_v4
    out = {}
    for arch in ("efficientnet_b0", "efficientnet_b4", "resnet18", "resnet50"):
        m = create_model(arch, num_classes=2)
        out[arch] = dict(state=[[k, list(v.shape)] for k, v in m.state_dict().items()],
                         params=[[k, list(v.shape)] for k, v in m.named_parameters()],
                         n_params=sum(p.numel() for p in m.parameters()))
    m = create_deepfake_model_v4("efficientnet_deepfake_v4", num_classes=2, in_chans=12)
    out["efficientnet_deepfake_v4"] = dict(
        state=[[k, list(v.shape)] for k, v in m.state_dict().items()],
        params=[[k, list(v.shape)] for k, v in m.named_parameters()],
        n_params=sum(p.numel() for p in m.parameters()))
    with open(os.path.join(GOLDEN, "state_keys.json"), "w") as f:
        json.dump(out, f)
    print("state_keys.json:", {k: v["n_params"] for k, v in out.items()})


def mint_step(arch, batch, H, W, n_steps=2, smoothing=0.0, opt_name="sgd", soft=False, tag=""):
    from dfd.timm.loss import LabelSmoothingCrossEntropy, SoftTargetCrossEntropy
    from dfd.timm.models import create_model
    from dfd.timm.optim import create_optimizer
    from dfd.timm.utils import accuracy
    torch.manual_seed(0)
    spec = get_spec(arch)
    model = create_model(arch, num_classes=2)
    sd = synth_state(spec, seed=7)
    model.load_state_dict(sd, strict=True)
    model.train()
    lr = 0.01 if opt_name == "sgd" else 1e-3
    wd = 1e-4
    args = _args(opt=opt_name, lr=lr, weight_decay=wd)
    optimizer = create_optimizer(args, model)
    if soft:
        loss_fn = SoftTargetCrossEntropy()
    elif smoothing > 0:
        loss_fn = LabelSmoothingCrossEntropy(smoothing)
    else:
        loss_fn = torch.nn.CrossEntropyLoss()
    rec = dict(arch=arch, batch=batch, H=H, W=W, weight_seed=7, opt=opt_name, lr=lr, momentum=0.9,
               weight_decay=wd, smoothing=smoothing, soft=soft, torch=torch.__version__, steps=[])
    for step in range(n_steps):
        x, y = synth_batch(batch, 3, H, W, seed=1234 + step, soft=soft)
        out = model(x)
        loss = loss_fn(out, y)
        prec1 = accuracy(out, y, topk=(1,))
        optimizer.zero_grad()
        loss.backward()
        grads = {k: _summ(p.grad) for k, p in model.named_parameters()}
        optimizer.step()
        rec["steps"].append(dict(
            logits=out.detach().tolist(), loss=float(loss), prec1=float(prec1), grads=grads,
            params={k: _summ(p) for k, p in model.named_parameters()},
            buffers={k: _summ(b.float()) for k, b in model.named_buffers()}))
    # eval-mode forward with the post-training running stats (validate path)
    model.eval()
    with torch.no_grad():
        x, y = synth_batch(batch, 3, H, W, seed=999)
        out = model(x)
        rec["eval"] = dict(logits=out.tolist(), loss=float(torch.nn.CrossEntropyLoss()(out, y)))
    name = "step_%s%s.json" % (arch, tag)
    with open(os.path.join(GOLDEN, name), "w") as f:
        json.dump(rec, f)
    print(name, "loss", [s["loss"] for s in rec["steps"]], "eval", rec["eval"]["loss"])


def mint_step_dropped(arch="efficientnet_b0", batch=4, H=64, W=64, drop_rate=0.2, drop_path_rate=0.2):
    """One train step of the reference with stochastic depth + classifier dropout (the production configuration,
    scripts/train.sh: --drop 0.35 --drop-connect 0.2).  The reference draws its masks from torch's global generator
    (layers/drop.py:96, F.dropout at efficientnet.py:347); torch.rand / F.dropout are wrapped here ONLY to record what was
    drawn (the reference code runs unmodified), so that the oracle and the CUDA path can be given the same masks."""
    import torch.nn.functional as F
    from dfd.timm.models import create_model
    from dfd.timm.optim import create_optimizer
    torch.manual_seed(0)
    spec = get_spec(arch)
    model = create_model(arch, num_classes=2, drop_rate=drop_rate, drop_path_rate=drop_path_rate)
    model.load_state_dict(synth_state(spec, seed=7), strict=True)
    model.train()
    args = _args(opt="sgd", lr=0.01, weight_decay=1e-4)
    optimizer = create_optimizer(args, model)
    rates = [(b.name, float(m.drop_path_rate)) for b, m in zip(spec.blocks, [blk for st in model.blocks for blk in st])]
    rands, drops = [], []
    real_rand, real_dropout = torch.rand, F.dropout

    def rec_rand(*a, **k):
        t = real_rand(*a, **k)
        rands.append(t.clone())
        return t

    def rec_dropout(x, p=0.5, training=True, inplace=False):
        out = real_dropout(x, p, training, False)
        drops.append(((out != 0) | (x == 0)).float() / (1.0 - p) if training and p > 0 else torch.ones_like(x))
        return out

    x, y = synth_batch(batch, 3, H, W, seed=1234)
    torch.rand, F.dropout = rec_rand, rec_dropout
    try:
        torch.manual_seed(5)
        out = model(x)
    finally:
        torch.rand, F.dropout = real_rand, real_dropout
    loss = torch.nn.CrossEntropyLoss()(out, y)
    optimizer.zero_grad()
    loss.backward()
    grads = {k: _summ(p.grad) for k, p in model.named_parameters()}
    optimizer.step()
    # drop_path draws one torch.rand((N,1,1,1)) per residual block with rate > 0, in block order
    names = [n for (n, r), b in zip(rates, spec.blocks) if b.has_residual and r > 0]
    assert len(names) == len(rands), (len(names), len(rands))
    masks = {}
    for n, u in zip(names, rands):
        keep = 1.0 - dict(rates)[n]
        masks[n] = (torch.floor(keep + u) / keep).reshape(-1).tolist()
    rec = dict(arch=arch, batch=batch, H=H, W=W, weight_seed=7, lr=0.01, momentum=0.9, weight_decay=1e-4,
               drop_rate=drop_rate, drop_path_rate=drop_path_rate, block_rates=rates, drop_masks=masks,
               dropout_mask=drops[0].tolist(), logits=out.detach().tolist(), loss=float(loss), grads=grads,
               params={k: _summ(p) for k, p in model.named_parameters()}, torch=torch.__version__)
    with open(os.path.join(GOLDEN, "step_%s_dropped.json" % arch), "w") as f:
        json.dump(rec, f)
    print("step_%s_dropped.json loss" % arch, rec["loss"], "dropped samples per block", {n: m.count(0.0) for n, m in masks.items()})


def mint_aux():
    """Small formula fixtures from the reference's own helper code: ModelEma.update (utils.py:329-340) on a toy module, the
    prefetcher's normalisation (loader.py:229-253) on a uint8 batch, and drop_path itself (layers/drop.py:84-100)."""
    from dfd.timm.utils import ModelEma
    from dfd.timm.models.layers.drop import drop_path
    out = {}

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(2, 3, 1, bias=False)
            self.bn = torch.nn.BatchNorm2d(3)

    torch.manual_seed(3)
    m = Toy()
    ema = ModelEma(m, decay=0.9)
    hist = []
    for step in range(3):
        with torch.no_grad():
            for p_ in m.parameters():
                p_.add_(0.1 * (step + 1))
            m.bn.running_mean.add_(0.5)
            m.bn.num_batches_tracked.add_(7)
        ema.update(m)
        hist.append(dict(model={k: v.reshape(-1).tolist() for k, v in m.state_dict().items()},
                         ema={k: v.reshape(-1).tolist() for k, v in ema.ema.state_dict().items()}))
    out["ema"] = dict(decay=0.9, hist=hist)
    # PrefetchLoader arithmetic without CUDA: the same tensor expressions as loader.py:229-230,250-253
    from dfd.timm.data.constants import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
    g = torch.Generator().manual_seed(9)
    xb = torch.randint(0, 256, (2, 12, 5, 7), generator=g, dtype=torch.uint8)
    mean = torch.tensor([[v * 255 for v in IMAGENET_DEFAULT_MEAN] for _ in range(4)]).view(1, 12, 1, 1)
    std = torch.tensor([[v * 255 for v in IMAGENET_DEFAULT_STD] for _ in range(4)]).view(1, 12, 1, 1)
    out["normalize"] = dict(x=xb.tolist(), y=xb.float().sub_(mean).div_(std).tolist(), mean=list(IMAGENET_DEFAULT_MEAN),
                            std=list(IMAGENET_DEFAULT_STD), img_num=4)
    torch.manual_seed(11)
    xdp = torch.randn(6, 2, 2, 2)
    st = torch.random.get_rng_state()
    ydp = drop_path(xdp, 0.3, True)
    torch.random.set_rng_state(st)
    u = torch.rand((6, 1, 1, 1))
    out["drop_path"] = dict(x=xdp.tolist(), y=ydp.tolist(), u=u.reshape(-1).tolist(), drop_prob=0.3)
    with open(os.path.join(GOLDEN, "aux_formulas.json"), "w") as f:
        json.dump(out, f)
    print("aux_formulas.json ok")


def mint_optimizers():
    """3 steps of each optimizer on a toy parameter set with fixed gradients, via the reference factory."""
    from dfd.timm.optim import create_optimizer

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(3)
            self.w = torch.nn.Parameter(torch.randn(5, 7, generator=g))
            self.bias = torch.nn.Parameter(torch.randn(7, generator=g))
            self.k = torch.nn.Parameter(torch.randn(4, 1, 3, 3, generator=g))

    out = {}
    for name, lr in (("sgd", 0.1), ("adam", 1e-2), ("adamw", 1e-2), ("rmsproptf", 1e-2)):
        m = Toy()
        optimizer = create_optimizer(_args(opt=name, lr=lr, weight_decay=1e-2, opt_eps=1e-3), m)
        g = torch.Generator().manual_seed(11)
        hist = []
        for step in range(3):
            for p in m.parameters():
                p.grad = torch.randn(p.shape, generator=g)
            optimizer.step()
            hist.append({k: p.detach().reshape(-1).tolist() for k, p in m.named_parameters()})
        out[name] = dict(lr=lr, weight_decay=1e-2, eps=1e-3, momentum=0.9, hist=hist)
    with open(os.path.join(GOLDEN, "optimizers.json"), "w") as f:
        json.dump(out, f)
    print("optimizers.json ok")


def mint_runner_config1():
    """BASELINE config 1: ResNet-18, batch 8, 3x224x224, CPU/gloo world_size=1 through the reference's own
    train_epoch / validate (dfd/runners/train.py:594-766) with torch DDP."""
    import torch.distributed as dist
    T = ref_shims.import_train_runner()
    from dfd.timm.models import create_model
    from dfd.timm.optim import create_optimizer
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    dist.init_process_group("gloo", rank=0, world_size=1)
    torch.manual_seed(0)
    spec = get_spec("resnet18")
    model = create_model("resnet18", num_classes=2)
    model.load_state_dict(synth_state(spec, seed=7))
    args = SimpleNamespace(opt="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4, opt_eps=1e-8,
                           prefetcher=True, mixup=0.0, mixup_off_epoch=0, num_classes=2, smoothing=0.0,
                           distributed=True, world_size=1, local_rank=0, log_interval=1, save_images=False,
                           recovery_interval=0, tta=0)
    optimizer = create_optimizer(args, model)
    ddp = torch.nn.parallel.DistributedDataParallel(model)

    class Loader(list):
        mixup_enabled = False

    batches = Loader(synth_batch(8, 3, 224, 224, seed=1234 + i) for i in range(2))
    loss_fn = torch.nn.CrossEntropyLoss()
    tm = T.train_epoch(0, ddp, batches, optimizer, loss_fn, args)
    vm = T.validate(ddp, batches, loss_fn, args)
    rec = dict(train={k: float(v) for k, v in tm.items()}, validate={k: float(v) for k, v in vm.items()},
               params={k: _summ(p) for k, p in model.named_parameters()},
               buffers={k: _summ(b.float()) for k, b in model.named_buffers()}, torch=torch.__version__)
    with open(os.path.join(GOLDEN, "runner_config1_resnet18.json"), "w") as f:
        json.dump(rec, f)
    dist.destroy_process_group()
    print("runner_config1:", rec["train"], rec["validate"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    ref_shims.install()
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    if a.only in ("", "keys"):
        mint_state_keys()
    if a.only in ("", "steps"):
        for arch, b, h, w in STEP_CASES:
            mint_step(arch, b, h, w)
        mint_step("efficientnet_b0", 4, 64, 64, smoothing=0.1, tag="_ls")
        mint_step("efficientnet_b0", 4, 64, 64, soft=True, opt_name="rmsproptf", tag="_soft_rmsprop")
        mint_step("efficientnet_b0", 4, 64, 64, opt_name="adamw", tag="_adamw")
    if a.only in ("", "drop"):
        mint_step_dropped()
    if a.only in ("", "aux"):
        mint_aux()
    if a.only in ("", "opt"):
        mint_optimizers()
    if a.only in ("", "runner"):
        mint_runner_config1()


if __name__ == "__main__":
    main()
