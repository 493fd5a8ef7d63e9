"""Paired crop of (x, y) (reference: src/crop.py).  Torch slicing glue, no kernel.

Reproduces the reference's behaviour exactly, including its batched-input quirk: MinSizePadding
reads x.shape[1] / x.shape[2] as height / width, so for the 4-D tensors Loss.forward passes
(B,C,H,W) it compares the crop size with C and H, and pads `size - C` zero rows at the bottom and
`size - H` zero columns on the right before cropping (src/crop.py:49-57)."""
from math import ceil

import torch
import torch.nn.functional as F
from torch.nn import Module

from sei_b200 import draws


class MinSizePadding(Module):
    def __init__(self, size, padding_mode="constant", fill=0):
        super().__init__()
        self.size = size
        self.padding_mode = padding_mode
        self.fill = fill

    def forward(self, x):
        h_padding = max(0, self.size - x.shape[1])
        w_padding = max(0, self.size - x.shape[2])
        if h_padding == 0 and w_padding == 0:
            return x
        # torchvision TF.pad(x, [left=0, top=0, right=w_padding, bottom=h_padding]) acts on the last two dims
        return F.pad(x, (0, w_padding, 0, h_padding), mode=self.padding_mode, value=self.fill)


class CropPair(Module):
    def __init__(self, location, size):
        super().__init__()
        assert location in ["random", "center"]
        self.location = location
        self.size = size

    def forward(self, x, y, xy_size_ratio=None):
        if xy_size_ratio is None:
            xy_size_ratio = int(ceil(x.shape[1] / y.shape[1]))
        r = xy_size_ratio
        x = MinSizePadding(self.size * r)(x)
        y = MinSizePadding(self.size)(y)
        h, w = y.shape[-2:]
        if self.location == "random":
            # two draws from the CPU generator, rows first (reference :26-27)
            i = draws.randint(0, h - self.size + 1)
            j = draws.randint(0, w - self.size + 1)
        else:
            i = (h - self.size) // 2
            j = (w - self.size) // 2
        x_crop = _crop(x, i * r, j * r, self.size * r, self.size * r)
        y_crop = _crop(y, i, j, self.size, self.size)
        return x_crop, y_crop


def _crop(t, top, left, height, width):
    """torchvision TF.crop: a slice; regions beyond the image are zero-padded."""
    H, W = t.shape[-2:]
    if top + height > H or left + width > W:
        t = F.pad(t, (0, max(0, left + width - W), 0, max(0, top + height - H)))
    return t[..., top:top + height, left:left + width]
