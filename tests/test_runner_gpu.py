"""-m gpu: the reference-facing entry points (create_model / create_optimizer / train_epoch / validate / checkpoint
layout) driving the native engine, against the oracle's restatement of the same loop."""
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(**kw):
    d = dict(opt="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4, opt_eps=1e-8, prefetcher=True, mixup=0.0, mixup_off_epoch=0,
             num_classes=2, smoothing=0.0, distributed=False, world_size=1, local_rank=0, log_interval=1, save_images=False,
             recovery_interval=0, tta=0, model="efficientnet_b0")
    d.update(kw)
    return SimpleNamespace(**d)


class _Loader(list):
    mixup_enabled = False


def _setup(dtype="fp16"):
    from deepfake_detection_b200.arch import get_spec
    from deepfake_detection_b200.models import create_model
    from deepfake_detection_b200.optim import create_optimizer
    from oracle.weights import synth_batch, synth_state
    spec = get_spec("efficientnet_b0")
    sd0 = synth_state(spec, seed=7)
    model = create_model("efficientnet_b0", num_classes=2, dtype=dtype)
    model.load_state_dict(sd0)
    args = _args()
    opt = create_optimizer(args, model)
    batches = _Loader((x.cuda(), y.cuda()) for x, y in (synth_batch(16, 3, 96, 96, seed=1234 + i) for i in range(2)))
    return spec, sd0, model, opt, args, batches


def _oracle_epoch(spec, sd0, batches, act_dtype):
    from oracle import train as OT
    sd = {k: v.clone() for k, v in sd0.items()}
    ost = OT.OptState(kind="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4)
    losses, precs = [], []
    for x, y in batches:
        o = OT.train_step(spec, sd, x.cpu(), y.cpu(), ost, act_dtype=act_dtype)
        losses.append(float(o["loss"]))
        precs.append(float(o["prec1"]))
    return sd, sum(losses) / len(losses), sum(precs) / len(precs)


def test_train_epoch_protocol_and_fused_match_oracle():
    from deepfake_detection_b200 import loss as NL
    from deepfake_detection_b200.runners.train import train_epoch, validate
    res = {}
    for flavour in ("protocol", "fused"):
        spec, sd0, model, opt, args, batches = _setup()
        loss_fn = torch.nn.CrossEntropyLoss() if flavour == "protocol" else NL.CrossEntropyLoss()
        m = train_epoch(0, model, batches, opt, loss_fn, args)
        v = validate(model, batches, torch.nn.CrossEntropyLoss(), args)
        res[flavour] = (m, v, model.state_dict())
        assert set(m) == {"loss", "prec1", "learning_rate"} and m["learning_rate"] == 0.01
    sd_o, loss_o, prec_o = _oracle_epoch(spec, sd0, batches, torch.float16)
    for flavour, (m, v, sd) in res.items():
        assert abs(m["loss"] - loss_o) < 5e-3, (flavour, m, loss_o)
        assert abs(m["prec1"] - prec_o) <= 100.0 / 16 + 1e-6, (flavour, m, prec_o)
        worst = max(float((sd[k].cpu().float() - sd_o[k].float()).norm() / (sd_o[k].float().norm() + 1e-12))
                    for k in sd_o if sd_o[k].dtype.is_floating_point and k.endswith("weight") and sd_o[k].dim() > 1)
        assert worst < 5e-3, (flavour, worst)
    # the two flavours run the same kernels: same numbers up to atomics ordering
    assert abs(res["protocol"][0]["loss"] - res["fused"][0]["loss"]) < 2e-3
    assert abs(res["protocol"][1]["loss"] - res["fused"][1]["loss"]) < 5e-3


def test_checkpoint_layout_roundtrip(tmp_path):
    from deepfake_detection_b200.helpers import CheckpointSaver, load_checkpoint, resume_checkpoint
    from deepfake_detection_b200.models import create_model
    from deepfake_detection_b200.optim import create_optimizer
    from deepfake_detection_b200.runners.train import train_epoch
    spec, sd0, model, opt, args, batches = _setup()
    train_epoch(0, model, batches, opt, torch.nn.CrossEntropyLoss(), args)
    saver = CheckpointSaver(checkpoint_dir=str(tmp_path), recovery_dir=str(tmp_path))
    saver.save_checkpoint(model, opt, args, epoch=3, metric=71.0)
    path = os.path.join(str(tmp_path), "checkpoint-3.pth.tar")
    assert os.path.exists(path) and os.path.exists(os.path.join(str(tmp_path), "model_best.pth.tar"))
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert {"epoch", "arch", "state_dict", "optimizer", "args", "version", "metric"} <= set(ck) and ck["version"] == 2
    assert list(ck["state_dict"])[:3] == ["conv_stem.weight", "bn1.weight", "bn1.bias"]
    assert ck["state_dict"]["classifier.weight"].shape == (2, 1280)
    m2 = create_model("efficientnet_b0", num_classes=2, dtype="fp16")
    other, epoch = resume_checkpoint(m2, path)
    assert epoch == 4 and "optimizer" in other
    a, b = model.state_dict(), m2.state_dict()
    assert all(torch.equal(a[k].cpu(), b[k].cpu()) for k in a)
    o2 = create_optimizer(args, m2)
    o2.load_state_dict(other["optimizer"])
    assert torch.equal(o2.state_a.cpu(), opt.state_a.cpu())
    # model_half.pth.tar style: fp16 tensors in a bare state dict, 'module.' prefixes (helpers.py:18)
    half = {"module." + k: (v.half() if v.dtype.is_floating_point else v) for k, v in a.items()}
    hp = os.path.join(str(tmp_path), "model_half.pth.tar")
    torch.save({"state_dict": half}, hp)
    m3 = create_model("efficientnet_b0", num_classes=2, dtype="fp16")
    load_checkpoint(m3, hp, strict=False)
    c = m3.state_dict()
    assert float((c["conv_head.weight"].cpu() - a["conv_head.weight"].cpu()).abs().max()) < 1e-3
