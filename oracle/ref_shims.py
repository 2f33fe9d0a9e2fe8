"""Import shims that let the UNMODIFIED reference (`/root/reference`, read-only) import on a
CPU-only box with torch 2.x.  TEST INFRASTRUCTURE ONLY: used by `oracle/mint_goldens.py` (run in
the authoring container, where `/root/reference` exists) to mint the committed fixtures under
`tests/golden/`.  Nothing on the product path, in `-m gpu` tests, `smoke()` or `bench.py` imports
this file: the reference tree does not exist on the GPU box.

The six shims (none edits a reference file) and the breakage each one papers over:
  1. `torch._six` removed in torch 2.x           (dfd/timm/models/layers/helpers.py:6)
  2. absolute `import timm...` inside the vendored copy while only `dfd.timm` exists
                                                  (dfd/timm/models/efficientnet.py:32-33, resnet.py:16, ...)
  3. `assert has_apex`                            (dfd/runners/train.py:27-37)
  4. `import matplotlib.pyplot`                   (dfd/timm/utils.py:14)
  5. `import xmltodict`                           (dfd/utils.py:10)
  6. `torch.cuda.synchronize()` in the step body  (dfd/runners/train.py:639,740)
"""
import collections.abc
import importlib
import importlib.abc
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DFD_REFERENCE_ROOT", "/root/reference")


class _TimmAlias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name == "timm" or name.startswith("timm."):
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        return importlib.import_module("dfd." + spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Install the shims (idempotent). Raises if the reference tree is absent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "dfd")):
        raise RuntimeError("reference tree not found at %s (goldens are minted in the authoring "
                           "container only)" % REFERENCE_ROOT)
    import torch

    six = types.ModuleType("torch._six")
    six.container_abcs = collections.abc
    sys.modules["torch._six"] = six
    sys.meta_path.insert(0, _TimmAlias())
    sys.path.insert(0, REFERENCE_ROOT)
    apex, amp, par = (types.ModuleType(n) for n in ("apex", "apex.amp", "apex.parallel"))
    par.DistributedDataParallel = torch.nn.parallel.DistributedDataParallel
    par.convert_syncbn_model = torch.nn.SyncBatchNorm.convert_sync_batchnorm
    apex.amp, apex.parallel = amp, par
    sys.modules.update({
        "apex": apex, "apex.amp": amp, "apex.parallel": par,
        "xmltodict": types.ModuleType("xmltodict"),
        "matplotlib": types.ModuleType("matplotlib"),
        "matplotlib.pyplot": types.ModuleType("matplotlib.pyplot"),
    })
    torch.cuda.synchronize = lambda *a, **k: None
    _installed = True


def import_train_runner():
    """Returns the reference's `dfd.runners.train` module (train_epoch / validate callable on CPU)."""
    install()
    argv = sys.argv
    sys.argv = ["x"]  # train.py builds its argparse parsers at import time
    try:
        import dfd.runners.train as T
    finally:
        sys.argv = argv
    return T
