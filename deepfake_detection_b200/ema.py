"""`ModelEma` with the reference's interface (dfd/timm/utils.py:276-340) for NativeModel.

The reference keeps a deep copy of the model and, every step, walks its whole state_dict in Python:
`ema_v.copy_(ema_v * decay + (1 - decay) * model_v)` — two temporaries and ~3 kernels per tensor, >600 launches for
EfficientNet-B0.  Here the copy is another set of flat arenas and the update is ONE multi-tensor launch per arena
(`dfd_ema_update`: parameters, running statistics, and the int64 `num_batches_tracked` counters with the reference's
float arithmetic + truncation), enqueued on the caller's stream right behind the optimizer kernels.
"""
import logging
from collections import OrderedDict
from copy import deepcopy

import torch

from . import _lib
from .engine import _ptr


class ModelEma:
    def __init__(self, model, decay=0.9999, device="", resume=""):
        if device and str(device) not in ("cuda", str(getattr(getattr(model, "module", model).engine, "device", "cuda"))):
            raise _lib.NativeError("native ModelEma keeps the average on the model's GPU (device=%r requested)" % (device,))
        self.ema = deepcopy(model.module if hasattr(model, "module") else model)      # utils.py:300
        self.ema.eval()
        self.decay = decay
        self.device = device
        self.ema_has_module = hasattr(self.ema, "module")
        if resume:
            self._load_checkpoint(resume)
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def _load_checkpoint(self, checkpoint_path):
        checkpoint = torch.load(checkpoint_path, map_location="cpu")
        assert isinstance(checkpoint, dict)
        if "state_dict_ema" in checkpoint:
            sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in checkpoint["state_dict_ema"].items())
            self.ema.load_state_dict(sd)
            logging.info("Loaded state_dict_ema")
        else:
            logging.warning("Failed to find state_dict_ema, starting from loaded model weights")

    def update(self, model):
        m = model.module if hasattr(model, "module") else model
        src, dst = m.engine, self.ema.engine
        if src.n_params != dst.n_params:
            raise _lib.NativeError("ModelEma.update: architecture mismatch")
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("dfd_ema_update", _ptr(dst.params32), _ptr(src.params32), src.n_params, _ptr(dst.nbt), _ptr(src.nbt),
                  int(src.nbt.numel()), float(self.decay), st)
        _lib.call("dfd_ema_update", _ptr(dst.buffers32), _ptr(src.buffers32), int(src.buffers32.numel()), None, None, 0,
                  float(self.decay), st)
        dst.state_version += 1
        self.ema._weights_dirty = True        # the 16-bit kernel copies are refreshed lazily, when the EMA model is evaluated
