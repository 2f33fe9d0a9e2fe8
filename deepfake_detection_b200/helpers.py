"""Checkpoint I/O in the reference's on-disk layout, so `--resume`, `--initial-checkpoint` and the inference script's
`model_half.pth.tar` keep working across backends.

Layout (dfd/timm/utils.py:97-112): {'epoch', 'arch', 'state_dict', 'optimizer', 'args', 'version': 2,
['state_dict_ema'], ['metric']} with `state_dict` keyed by the reference module names (OIHW tensors, fp32 — or fp16
for model_half.pth.tar); a bare state-dict file is accepted too (dfd/timm/models/helpers.py:8-28).
"""
import glob
import logging
import operator
import os
import shutil
from collections import OrderedDict

import torch


def _strip_module(sd):
    out = OrderedDict()
    for k, v in sd.items():
        out[k[7:] if k.startswith("module") else k] = v       # helpers.py:18
    return out


def load_state_dict(checkpoint_path, use_ema=False):
    """dfd/timm/models/helpers.py:8-28"""
    if not (checkpoint_path and os.path.isfile(checkpoint_path)):
        logging.error("No checkpoint found at '%s'", checkpoint_path)
        raise FileNotFoundError()
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    key = "state_dict"
    if isinstance(ckpt, dict) and use_ema and "state_dict_ema" in ckpt:
        key = "state_dict_ema"
    if isinstance(ckpt, dict) and key in ckpt:
        return _strip_module(ckpt[key])
    return ckpt


def load_checkpoint(model, checkpoint_path, use_ema=False, strict=True, ignore_keys=None):
    """dfd/timm/models/helpers.py:31-44 (non-strict drops shape-mismatched keys)"""
    sd = load_state_dict(checkpoint_path, use_ema)
    for k in ignore_keys or ():
        sd.pop(k)
    if not strict:
        own = model.state_dict()
        for k in list(own):
            if k in sd and tuple(own[k].shape) != tuple(sd[k].shape):
                sd.pop(k)
    model.load_state_dict(sd, strict=strict)


def resume_checkpoint(model, checkpoint_path):
    """dfd/timm/models/helpers.py:47-73 -> (other_state, resume_epoch)"""
    if not os.path.isfile(checkpoint_path):
        logging.error("No checkpoint found at '%s'", checkpoint_path)
        raise FileNotFoundError()
    other, epoch = {}, None
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    if isinstance(ckpt, dict) and "state_dict" in ckpt:
        model.load_state_dict(_strip_module(ckpt["state_dict"]))
        for k in ("optimizer", "amp"):
            if k in ckpt:
                other[k] = ckpt[k]
        if "epoch" in ckpt:
            epoch = ckpt["epoch"]
            if ckpt.get("version", 0) > 1:
                epoch += 1
    else:
        model.load_state_dict(ckpt)
    return other, epoch


def unwrap_model(model):
    return model.module if hasattr(model, "module") else model


def get_state_dict(model):
    return unwrap_model(model).state_dict()


class CheckpointSaver:
    """Top-`max_history` checkpoints by metric + recovery files (dfd/timm/utils.py:36-149), same file names
    (`checkpoint-{epoch}.pth.tar`, `model_best.pth.tar`, `recovery-{epoch}-{batch}.pth.tar`) and payload."""

    def __init__(self, checkpoint_prefix="checkpoint", recovery_prefix="recovery", checkpoint_dir="", checkpoint_dir_bak="",
                 recovery_dir="", decreasing=False, max_history=10):
        assert max_history >= 1
        self.checkpoint_files = []
        self.best_epoch = self.best_metric = None
        self.curr_recovery_file = self.last_recovery_file = ""
        self.checkpoint_dir, self.checkpoint_dir_bak, self.recovery_dir = checkpoint_dir, checkpoint_dir_bak, recovery_dir
        self.save_prefix, self.recovery_prefix, self.extension = checkpoint_prefix, recovery_prefix, ".pth.tar"
        self.decreasing = decreasing
        self.cmp = operator.lt if decreasing else operator.gt
        self.max_history = max_history

    def _save(self, path, model, optimizer, args, epoch, model_ema=None, metric=None, use_amp=False):
        state = {"epoch": epoch, "arch": getattr(args, "model", None), "state_dict": get_state_dict(model),
                 "optimizer": optimizer.state_dict(), "args": args, "version": 2}
        if model_ema is not None:
            state["state_dict_ema"] = get_state_dict(model_ema)
        if metric is not None:
            state["metric"] = metric
        torch.save(state, path)

    def save_checkpoint(self, model, optimizer, args, epoch, model_ema=None, metric=None, use_amp=False):
        assert epoch >= 0
        worst = self.checkpoint_files[-1] if self.checkpoint_files else None
        if len(self.checkpoint_files) < self.max_history or metric is None or self.cmp(metric, worst[1]):
            if len(self.checkpoint_files) >= self.max_history:
                self._cleanup_checkpoints(1)
            path = os.path.join(self.checkpoint_dir, "-".join([self.save_prefix, str(epoch)]) + self.extension)
            self._save(path, model, optimizer, args, epoch, model_ema, metric, use_amp)
            self.checkpoint_files.append((path, metric))
            self.checkpoint_files.sort(key=lambda x: x[1], reverse=not self.decreasing)
            if metric is not None and (self.best_metric is None or self.cmp(metric, self.best_metric)):
                self.best_epoch, self.best_metric = epoch, metric
                shutil.copyfile(path, os.path.join(self.checkpoint_dir, "model_best" + self.extension))
                if self.checkpoint_dir_bak != "":
                    shutil.copyfile(path, os.path.join(self.checkpoint_dir_bak, "model_best" + self.extension))
        return (None, None) if self.best_metric is None else (self.best_metric, self.best_epoch)

    def _cleanup_checkpoints(self, trim=0):
        trim = min(len(self.checkpoint_files), trim)
        keep = self.max_history - trim
        if keep <= 0 or len(self.checkpoint_files) <= keep:
            return
        for path, _ in self.checkpoint_files[keep:]:
            try:
                os.remove(path)
            except OSError as e:
                logging.error("Exception '%s' while deleting checkpoint", e)
        self.checkpoint_files = self.checkpoint_files[:keep]

    def save_recovery(self, model, optimizer, args, epoch, model_ema=None, use_amp=False, batch_idx=0):
        assert epoch >= 0
        path = os.path.join(self.recovery_dir, "-".join([self.recovery_prefix, str(epoch), str(batch_idx)]) + self.extension)
        self._save(path, model, optimizer, args, epoch, model_ema, use_amp=use_amp)
        if os.path.exists(self.last_recovery_file):
            try:
                os.remove(self.last_recovery_file)
            except OSError as e:
                logging.error("Exception '%s' while removing %s", e, self.last_recovery_file)
        self.last_recovery_file, self.curr_recovery_file = self.curr_recovery_file, path

    def find_recovery(self):
        files = sorted(glob.glob(os.path.join(self.recovery_dir, self.recovery_prefix) + "*" + self.extension))
        return files[0] if files else ""
