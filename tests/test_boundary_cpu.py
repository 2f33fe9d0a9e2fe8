"""Host logic without a GPU: the reference-facing surface (names, signatures, config constants, checkpoint helpers)."""
import inspect
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch


def test_runner_signatures_match_reference():
    import deepfake_detection_b200.runners.train as T
    assert list(inspect.signature(T.train_epoch).parameters) == [
        "epoch", "model", "loader", "optimizer", "loss_fn", "args", "lr_scheduler", "saver", "output_dir", "use_amp", "model_ema"]
    assert list(inspect.signature(T.validate).parameters) == ["model", "loader", "loss_fn", "args", "log_suffix"]
    from deepfake_detection_b200.runners.test import test_img as ti
    assert list(inspect.signature(ti).parameters) == ["model_path", "img_files"]


def test_params_surface():
    from deepfake_detection_b200 import params as P
    assert P.img_mean.shape == (3, 1, 1) and abs(float(P.img_mean[0]) - 0.485 * 255) < 1e-4
    assert abs(float(P.img_std[2]) - 0.225 * 255) < 1e-4
    assert (P.image_max_height, P.image_max_width, P.image_max_w_h, P.img_num) == (600, 600, (600, 600), 4)
    img = np.full((300, 450, 3), 7, np.uint8)
    r = P.resize(img)
    assert r.shape == (400, 600, 3)
    p = P.padding_image(r)
    assert p.shape == (600, 600, 3) and p[0, 0, 0] == 0 and p[300, 300, 0] == 7 and p[99, 0, 0] == 0 and p[100, 0, 0] == 7
    m = P.DeepFakeModel(torch.nn.Linear(4, 2))
    out = m(torch.randn(3, 4))
    assert torch.allclose(out.sum(-1), torch.ones(3), atol=1e-6)


def test_factories_and_optimizer_names():
    from deepfake_detection_b200.models import create_deepfake_model_v4, create_model
    from deepfake_detection_b200.optim import create_optimizer
    with pytest.raises(RuntimeError):
        create_model("inception_v3")
    with pytest.raises(AssertionError):
        create_deepfake_model_v4("efficientnet_b0")
    assert list(inspect.signature(create_optimizer).parameters)[:2] == ["args", "model"]
    m = create_model("efficientnet_b0", num_classes=2)
    assert m.default_cfg["input_size"] == (3, 224, 224) and m.default_cfg["classifier"] == "classifier"


def test_losses_match_reference_formulas():
    from deepfake_detection_b200 import loss as NL
    from oracle import model as OM
    g = torch.Generator().manual_seed(0)
    z = torch.randn(16, 2, generator=g)
    y = torch.randint(0, 2, (16,), generator=g)
    assert torch.allclose(NL.LabelSmoothingCrossEntropy(0.1)(z, y), OM.cross_entropy(z, y, 0.1), atol=1e-6)
    soft = torch.softmax(torch.randn(16, 2, generator=g), -1)
    assert torch.allclose(NL.SoftTargetCrossEntropy()(z, soft), OM.cross_entropy(z, soft), atol=1e-6)
    assert torch.allclose(NL.CrossEntropyLoss()(z, y), OM.cross_entropy(z, y), atol=1e-6)


def test_checkpoint_helpers_layout(tmp_path):
    from deepfake_detection_b200.helpers import CheckpointSaver, load_state_dict, resume_checkpoint

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(3, 2)

    m, opt = Toy(), None
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    saver = CheckpointSaver(checkpoint_dir=str(tmp_path), recovery_dir=str(tmp_path), max_history=2)
    args = SimpleNamespace(model="toy")
    for ep, metric in enumerate([50.0, 60.0, 55.0]):
        best = saver.save_checkpoint(m, opt, args, ep, metric=metric)
    assert best == (60.0, 1)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("checkpoint")) == ["checkpoint-1.pth.tar", "checkpoint-2.pth.tar"]
    sd = load_state_dict(os.path.join(str(tmp_path), "model_best.pth.tar"))
    assert set(sd) == {"fc.weight", "fc.bias"}
    other, epoch = resume_checkpoint(Toy(), os.path.join(str(tmp_path), "checkpoint-2.pth.tar"))
    assert epoch == 3 and "optimizer" in other
    saver.save_recovery(m, opt, args, 2, batch_idx=7)
    assert saver.find_recovery().endswith("recovery-2-7.pth.tar")
