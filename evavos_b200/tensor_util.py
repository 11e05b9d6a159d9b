"""Padding helpers on the propagation path (mivos/tensor_util.py:62-93); plain torch, not a hot spot."""
from __future__ import annotations

import torch.nn.functional as F


def pad_divide_by(in_img, d, in_size=None):
    """Symmetric zero-pad of the last two dims to multiples of ``d``; returns (padded, (lw, uw, lh, uh))."""
    h, w = in_img.shape[-2:] if in_size is None else in_size
    extra_h = (-h) % d
    extra_w = (-w) % d
    lh, lw = extra_h // 2, extra_w // 2
    pad_array = (int(lw), int(extra_w - lw), int(lh), int(extra_h - lh))
    return F.pad(in_img, pad_array), pad_array


def unpad(img, pad):
    """Inverse of pad_divide_by on a (..., H, W) 4-D tensor."""
    if pad[2] + pad[3] > 0:
        img = img[:, :, pad[2]:img.shape[2] - pad[3], :]
    if pad[0] + pad[1] > 0:
        img = img[:, :, :, pad[0]:img.shape[3] - pad[1]]
    return img


def unpad_3dim(img, pad):
    if pad[2] + pad[3] > 0:
        img = img[:, pad[2]:img.shape[1] - pad[3], :]
    if pad[0] + pad[1] > 0:
        img = img[:, :, pad[0]:img.shape[2] - pad[1]]
    return img
