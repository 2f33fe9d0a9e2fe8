"""CPU ORACLE — TEST INFRASTRUCTURE (see oracle/model.py header).

Plain torch restatements of the small formulas around the step, pinned to the reference by tests/golden/aux_formulas.json
(tests/test_oracle_vs_reference_goldens.py::test_aux_formulas_match_reference):

  drop_path_mask : dfd/timm/models/layers/drop.py:94-99   mask = floor(keep + u) / keep, one u per sample
  normalize_u8   : dfd/timm/data/loader.py:229-230,250-253  uint8 -> float, (x - mean*255) / (std*255), per-frame repeat
  ema_update     : dfd/timm/utils.py:329-340             ema*decay + (1-decay)*model, copy_() back into the buffer's dtype
"""
import torch


def drop_path_mask(u, drop_prob):
    keep = 1.0 - drop_prob
    return torch.floor(keep + u) / keep


def normalize_u8(x_u8, mean, std, img_num):
    c = 3 * img_num
    m = torch.tensor([[v * 255 for v in mean] for _ in range(img_num)]).view(1, c, 1, 1)
    s = torch.tensor([[v * 255 for v in std] for _ in range(img_num)]).view(1, c, 1, 1)
    return x_u8.float().sub_(m).div_(s)


def ema_update(ema_v, model_v, decay):
    out = ema_v.clone()
    out.copy_(ema_v * decay + (1.0 - decay) * model_v)
    return out
