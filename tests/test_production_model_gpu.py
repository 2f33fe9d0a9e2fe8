"""-m gpu: the model the reference actually ships and trains: `efficientnet_deepfake_v4`
(efficientnet.py:806-851: stem 256, B7-depth body of 55 blocks, head 256), 12 input channels = 4 RGB frames, 600 x 600.

  * N3 — one production-shaped train step, per-GPU batch 3 (scripts/train.sh:3-23: RMSpropTF, mixup soft targets, drop 0.35,
    drop-connect 0.2, fp16 under apex AMP O1 = dynamic loss scaling) against the oracle with the SAME stochastic masks;
  * N2 — `test_img(model_path, img_files)` (dfd/runners/test.py:29-60) on a synthetic `model_half.pth.tar`: checkpoint layout,
    fp16 eval forward with BN frozen into per-channel scale / shift, softmax score of column 0."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_deepfake_v4_production_train_step():
    from deepfake_detection_b200.arch import get_spec, param_entries
    from deepfake_detection_b200.trainer import Trainer
    from oracle import train as OT
    from oracle.weights import synth_batch, synth_state
    torch.set_num_threads(int(os.environ.get("DFD_ORACLE_THREADS", "32")))
    torch.manual_seed(77)
    arch, N, R = "efficientnet_deepfake_v4", 3, 600
    spec = get_spec(arch, in_chans=12)
    assert len(spec.blocks) == 55 and spec.stem == 256 and spec.num_features == 256
    sd0 = synth_state(spec, seed=7)
    x, y = synth_batch(N, 12, R, R, seed=1, soft=True)                    # mixup targets: float [N, 2]
    hp = dict(lr=1e-3, momentum=0.9, weight_decay=1e-5)
    tr = Trainer(arch, N, R, R, in_chans=12, dtype="fp16", opt="rmsproptf", opt_eps=1e-3, drop_rate=0.35, drop_path_rate=0.2, **hp)
    assert tr.dynamic_scale                                                # fp16 => apex-O1-style dynamic loss scaling on the device
    tr.load_state_dict(sd0)
    tr.train_step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    e = tr.engine
    assert float(e.loss_scale_state[0]) == 65536.0 and int(e.flags[0]) == 0          # no overflow, step applied
    masks = {k: v[:, 0].cpu().clone() for k, v in e.drop_masks.items()}
    n_res = sum(1 for i, b in enumerate(spec.blocks) if b.has_residual and i > 0)
    assert len(masks) == n_res == 48
    sd = {k: v.clone() for k, v in sd0.items()}
    ost = OT.OptState(kind="rmsproptf", eps=1e-3, **hp)
    out = OT.train_step(spec, sd, x, y, ost, act_dtype=torch.float16, drop_masks=masks, dropout_mask=e.dropout_mask.cpu().clone())
    loss_rel = abs(float(e.loss) - float(out["loss"])) / abs(float(out["loss"]))
    pn = [n for n, _, _ in param_entries(spec)]
    w_worst = max((_rel(e.param_view(n), sd[n]), n) for n in pn)
    logits_rel = _rel(e.logits, out["logits"])
    buf_worst = max(_rel(e.buffer_view(n).float(), sd[n].float()) for n in sd if n not in pn and not n.endswith("num_batches_tracked"))
    # batch 3 with train-mode BN: the last stages normalise over 3 x 19 x 19 values per channel; fp16 storage noise of ~330
    # stored tensors stays well inside the north_star 1e-2 on what the optimizer consumes
    assert loss_rel < 1e-2 and w_worst[0] < 1e-2 and buf_worst < 1e-2, (loss_rel, w_worst, buf_worst, logits_rel)
    assert logits_rel < 5e-2, logits_rel


def test_test_img_roundtrip_on_model_half(tmp_path):
    from PIL import Image
    from deepfake_detection_b200.arch import get_spec
    from deepfake_detection_b200.params import img_mean, img_std, padding_image, resize
    from deepfake_detection_b200.runners.test import test_img as native_test_img
    from oracle import train as OT
    from oracle.weights import synth_state
    torch.set_num_threads(int(os.environ.get("DFD_ORACLE_THREADS", "32")))
    spec = get_spec("efficientnet_deepfake_v4", in_chans=12)
    sd = synth_state(spec, seed=7)
    # model_half.pth.tar (test.py:64, helpers.py:8-44): fp16 tensors, DDP 'module.' prefixes, under 'state_dict'
    path = str(tmp_path / "model_half.pth.tar")
    torch.save({"state_dict": {"module." + k: (v.half() if v.dtype.is_floating_point else v) for k, v in sd.items()}}, path)
    rng = np.random.RandomState(0)
    files = []
    for i, (h, w) in enumerate(((300, 450), (640, 480))):
        f = str(tmp_path / ("img%d.png" % i))
        Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)).save(f)
        files.append(f)
    scores = native_test_img(path, files)
    assert len(scores) == 2 and all(0.0 <= s <= 1.0 for s in scores)
    sd_h = {k: (v.half().float() if v.dtype.is_floating_point else v) for k, v in sd.items()}      # what the file holds
    for f, s in zip(files, scores):
        img = np.transpose(padding_image(resize(np.array(Image.open(f).convert("RGB"), np.uint8))), (2, 0, 1))
        t = torch.from_numpy(img).float().sub_(img_mean).div_(img_std).half().float()
        clip = torch.cat([t, t, t, t], dim=0).unsqueeze(0)
        ev = OT.validate_step(spec, sd_h, clip, torch.zeros(1, dtype=torch.int64), act_dtype=torch.float16)
        ref = float(torch.softmax(ev["logits"], -1)[0, 0])
        assert abs(s - ref) < 1e-2, (f, s, ref)                       # a probability: absolute 1e-2


def test_eval_forward_freezes_bn_and_tracks_weight_changes():
    """inference plan: the per-channel BN scale / shift ('folded' BN) are derived once per weight state - a second eval forward
    launches no dfd_bn_finalize - and re-derived after training moved the running statistics"""
    from deepfake_detection_b200 import _lib
    from deepfake_detection_b200.arch import get_spec
    from deepfake_detection_b200.models import create_model
    from oracle.weights import synth_batch, synth_state
    spec = get_spec("efficientnet_b0")
    m = create_model("efficientnet_b0", num_classes=2, dtype="fp16")
    m.load_state_dict(synth_state(spec, seed=7))
    x, y = synth_batch(8, 3, 96, 96, seed=3)
    m.eval()
    with torch.no_grad():
        a = m(x.cuda()).clone()
        c0 = _lib.N_CALLS[0]
        b = m(x.cuda()).clone()
        second = _lib.N_CALLS[0] - c0
    e = m.engine_for(8, 96, 96)
    n_fin = sum(1 for op in e.fwd_ops if op[1] == "dfd_bn_finalize")
    assert torch.equal(a, b) and n_fin == 49
    assert second <= len(e.fwd_ops) - n_fin + 4, (second, len(e.fwd_ops))          # no BN kernel in the steady state
    m.train()
    out = m(x.cuda())
    torch.nn.CrossEntropyLoss()(out, y.cuda()).backward()                           # running statistics moved
    m.eval()
    with torch.no_grad():
        c = m(x.cuda()).clone()
    assert not torch.equal(a, c)
