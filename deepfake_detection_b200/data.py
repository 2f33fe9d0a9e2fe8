"""The step immediately before the hot path: the prefetcher's image normalisation.

`NativePrefetchLoader` has the interface and the stream structure of the reference's `PrefetchLoader`
(dfd/timm/data/loader.py:213-279): it wraps a loader of (uint8 NCHW batch, target), and on a side stream uploads the
NEXT batch and normalises it — `(x - mean*255) / (std*255)` with mean / std repeated per frame (`img_num` x RGB,
dfd/params.py:24-27) — while the current batch trains; `torch.cuda.current_stream().wait_stream(stream)` joins them
exactly where the reference does (loader.py:259).  Differences, all on purpose:
  * the three full-tensor passes of the reference (`.half()`/`.float()`, `sub_`, `div_`) are ONE kernel
    (`dfd_input_normalize`) that reads the uint8 batch once and writes the 16-bit tensor the stem kernel consumes;
  * the upload is the uint8 batch (1 byte / value), not a float tensor; pinned host batches make it asynchronous;
  * two device staging buffers are recycled instead of allocating every iteration.
Random erasing (loader.py:253-254) is not on the native path: `re_prob > 0` raises.
"""
import torch

from . import _lib
from .engine import _ptr

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


class InputNormalizer:
    """uint8 NCHW -> 16-bit NCHW on the current stream; holds the per-channel mean*255 / std*255 vectors on the device."""

    def __init__(self, mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD, img_num=1, dtype="bf16", device=None):
        dev = device if device is not None else "cuda:%d" % torch.cuda.current_device()
        self.mean255 = torch.tensor([x * 255 for x in mean] * img_num, dtype=torch.float32, device=dev)     # loader.py:229
        self.std255 = torch.tensor([x * 255 for x in std] * img_num, dtype=torch.float32, device=dev)      # loader.py:230
        if dtype in ("bf16", "bfloat16", torch.bfloat16):
            self.dt, self.tdtype = _lib.DT_BF16, torch.bfloat16
        elif dtype in ("fp16", "float16", "half", torch.float16):
            self.dt, self.tdtype = _lib.DT_FP16, torch.float16
        else:
            raise ValueError("dtype %r" % (dtype,))

    def __call__(self, x_u8, out=None):
        if x_u8.dtype != torch.uint8 or x_u8.dim() != 4 or not x_u8.is_cuda or not x_u8.is_contiguous():
            raise ValueError("InputNormalizer expects a contiguous CUDA uint8 NCHW batch")
        n, c, h, w = x_u8.shape
        if c != self.mean255.numel():
            raise ValueError("batch has %d channels, normaliser %d" % (c, self.mean255.numel()))
        if out is None:
            out = torch.empty(x_u8.shape, dtype=self.tdtype, device=x_u8.device)
        _lib.call("dfd_input_normalize", _ptr(x_u8), _ptr(self.mean255), _ptr(self.std255), _ptr(out), n, c, h, w, self.dt,
                  torch.cuda.current_stream().cuda_stream)
        return out


class NativePrefetchLoader:
    def __init__(self, loader, mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD, fp16=False, re_prob=0., re_mode="const",
                 re_count=1, re_num_splits=0, re_max=0.1, img_num=4, dtype=None):
        if re_prob > 0.:
            raise _lib.NativeError("random erasing (loader.py:237-241) is not on the native prefetch path")
        self.loader = loader
        self.norm = InputNormalizer(mean, std, img_num, dtype if dtype is not None else ("fp16" if fp16 else "bf16"))
        self._stage = [None, None]          # device uint8 staging, recycled
        self._out = [None, None]

    def __iter__(self):
        stream = torch.cuda.Stream()
        first = True
        slot = 0
        input = target = None
        for next_input, next_target in self.loader:
            with torch.cuda.stream(stream):
                if next_input.dtype != torch.uint8:
                    raise ValueError("NativePrefetchLoader expects uint8 batches (fast_collate, loader.py:14-41)")
                st = self._stage[slot]
                if st is None or st.shape != next_input.shape:
                    st = self._stage[slot] = torch.empty(next_input.shape, dtype=torch.uint8, device="cuda")
                    self._out[slot] = torch.empty(next_input.shape, dtype=self.norm.tdtype, device="cuda")
                st.copy_(next_input, non_blocking=True)
                next_target = next_target.cuda(non_blocking=True)
                next_input = self.norm(st, out=self._out[slot])
            slot ^= 1
            if not first:
                yield input, target
            else:
                first = False
            torch.cuda.current_stream().wait_stream(stream)
            # the consumer of `input` runs on the current stream: the side stream may not recycle its buffer before that
            stream.wait_stream(torch.cuda.current_stream())
            input, target = next_input, next_target
        yield input, target

    def __len__(self):
        return len(self.loader)

    @property
    def sampler(self):
        return self.loader.sampler

    @property
    def mixup_enabled(self):
        cf = getattr(self.loader, "collate_fn", None)
        return getattr(cf, "mixup_enabled", False)

    @mixup_enabled.setter
    def mixup_enabled(self, x):
        cf = getattr(self.loader, "collate_fn", None)
        if cf is not None and hasattr(cf, "mixup_enabled"):
            cf.mixup_enabled = x
