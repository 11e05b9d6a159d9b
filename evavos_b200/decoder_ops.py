"""In-place elementwise tails of the decoder's convolutions on channels-last tensors (csrc/decoder_ops.cu).

The STCN decoder (prop_net.py:13-30; modules.py ResBlock / UpsampleBlock) has no normalisation layers: between its
convolutions sit bias adds, residual adds, ReLUs and two bilinear x2 upsamplings, one PyTorch kernel each.  The
convolutions are issued without bias and these two calls finish them.
"""
from __future__ import annotations

import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}          # EVAVOS_F32 / EVAVOS_BF16


def _nhwc(name, t):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor required (no CPU path in evavos_b200)")
    if t.dim() != 4 or not t.is_contiguous(memory_format=torch.channels_last):
        raise ValueError(f"{name}: a dense 4-D channels_last tensor is required")
    return t


def bias_residual_(y: torch.Tensor, bias: torch.Tensor, residual: torch.Tensor | None = None, relu: bool = False):
    """y <- [relu](y + bias[c] (+ residual)), in place; y (n,C,H,W) channels_last fp32 / bf16, bias fp32 (C)."""
    lib = _lib.load()
    _nhwc("bias_residual_", y)
    if residual is not None and (_nhwc("bias_residual_", residual).shape != y.shape or residual.dtype != y.dtype):
        raise ValueError("bias_residual_: residual must match y")
    n, c, h, w = y.shape
    if bias.dtype != torch.float32 or bias.numel() != c or not bias.is_contiguous():
        raise ValueError("bias_residual_: bias must be a contiguous fp32 vector of C entries")
    with torch.cuda.device(y.device):
        _lib.check(lib.evavos_bias_residual_nhwc(y.data_ptr(), bias.data_ptr(), residual.data_ptr() if residual is not None else None,
                                                 n * h * w, c, _DTYPES[y.dtype], int(bool(relu)), _lib.current_stream_ptr(y.device)))
    return y


def upsample2x_add_(y: torch.Tensor, bias: torch.Tensor, x: torch.Tensor):
    """y <- y + bias[c] + F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), in place;
    y (n,C,H,W), x (n,C,H/2,W/2), both channels_last of one dtype."""
    lib = _lib.load()
    _nhwc("upsample2x_add_", y)
    _nhwc("upsample2x_add_", x)
    n, c, h, w = y.shape
    if tuple(x.shape) != (n, c, h // 2, w // 2) or h % 2 or w % 2 or x.dtype != y.dtype:
        raise ValueError("upsample2x_add_: x must be (n, C, H/2, W/2) of y's dtype")
    if bias.dtype != torch.float32 or bias.numel() != c or not bias.is_contiguous():
        raise ValueError("upsample2x_add_: bias must be a contiguous fp32 vector of C entries")
    with torch.cuda.device(y.device):
        _lib.check(lib.evavos_upsample2x_add_nhwc(y.data_ptr(), bias.data_ptr(), x.data_ptr(), n, h, w, c, _DTYPES[y.dtype],
                                                  _lib.current_stream_ptr(y.device)))
    return y
