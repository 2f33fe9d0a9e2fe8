"""Inference-side configuration surface of the reference, preserved in name and meaning (dfd/params.py:24-67):
`img_mean`, `img_std` (ImageNet statistics x255, shape [3,1,1]), `image_max_height`, `image_max_width`,
`image_max_w_h` (width, height), `img_num`, `DeepFakeModel`, `resize`, `padding_image`."""
import numpy as np
import torch
import torch.nn as nn

img_mean = torch.tensor([v * 255 for v in (0.485, 0.456, 0.406)]).view(3, 1, 1)      # params.py:24-27
img_std = torch.tensor([v * 255 for v in (0.229, 0.224, 0.225)]).view(3, 1, 1)
image_max_height = 600
image_max_width = 600
image_max_w_h = (image_max_width, image_max_height)
img_num = 4                                                                          # frames per clip (12 channels)


class DeepFakeModel(nn.Module):
    """softmax over the base model's logits; column 0 is the fake probability (params.py:34-42, test.py:59)"""

    def __init__(self, model):
        super().__init__()
        self.basemodel = model
        self.softmax = nn.Softmax(-1)

    def forward(self, x):
        return self.softmax(self.basemodel(x))


def resize(image):
    """aspect-preserving resize so that the image fits the 600x600 canvas (params.py:45-55). image: HxWxC uint8."""
    import cv2
    h, w = image.shape[0:2]
    if float(h) / w > float(image_max_w_h[1]) / image_max_w_h[0]:
        th = image_max_w_h[1]
        tw = int(w * float(th) / h)
    else:
        tw = image_max_w_h[0]
        th = int(h * float(tw) / w)
    return cv2.resize(image, (tw, th))


def padding_image(image):
    """centre the image on a zero 600x600 canvas (params.py:58-67)"""
    h, w = image.shape[0:2]
    if h == image_max_height and w == image_max_width:
        return image
    top = int((image_max_height - h) / 2)
    left = int((image_max_width - w) / 2)
    return np.pad(image, ((top, image_max_height - h - top), (left, image_max_width - w - left), (0, 0)), "constant",
                  constant_values=0)
