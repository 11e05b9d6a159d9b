"""aggregate_wbg with the reference's signature (mivos/model/aggregate.py:22-37), one fused kernel."""
from __future__ import annotations

import torch

from . import _lib


def aggregate_wbg(prob: torch.Tensor, keep_bg: bool = False, hard: bool = False) -> torch.Tensor:
    """prob (K,1,h,w) fp32 object probabilities -> (K+1,1,h,w) if keep_bg else (K,1,h,w).

    Background = prod_k (1 - p_k); all channels clamped to [1e-7, 1-1e-7], turned into logits
    (x1000 when ``hard``) and soft-maxed over the object axis.
    """
    lib = _lib.load()
    if not prob.is_cuda:
        raise RuntimeError("aggregate_wbg: CUDA tensor required (no CPU path in evavos_b200)")
    if prob.dim() != 4:
        raise ValueError("prob must be (K,1,h,w)")
    k, c, h, w = prob.shape
    p = prob.to(torch.float32).contiguous()
    npix = c * h * w
    out = torch.empty(((k + 1) if keep_bg else k, c, h, w), dtype=torch.float32, device=prob.device)
    with torch.cuda.device(prob.device):
        _lib.check(lib.evavos_aggregate_wbg(p.data_ptr(), out.data_ptr(), k, npix, int(bool(keep_bg)), int(bool(hard)),
                                            _lib.current_stream_ptr(prob.device)))
    return out


def argmax_unpad(prob: torch.Tensor, pad, h: int, w: int):
    """prob (C,T,1,nh,nw) fp32 -> (masks uint8 (T,1,nh,nw), unpadded uint8 (T,h,w)), one kernel for all frames.

    ``pad`` is the (lw, uw, lh, uh) tuple of pad_divide_by (mivos/tensor_util.py:62-80); replaces
    inference_core.py:247-257.
    """
    lib = _lib.load()
    if not prob.is_cuda:
        raise RuntimeError("argmax_unpad: CUDA tensor required (no CPU path in evavos_b200)")
    c, t, _, nh, nw = prob.shape
    p = prob.to(torch.float32).contiguous()
    masks = torch.empty((t, 1, nh, nw), dtype=torch.uint8, device=prob.device)
    out = torch.empty((t, h, w), dtype=torch.uint8, device=prob.device)
    with torch.cuda.device(prob.device):
        _lib.check(lib.evavos_argmax_unpad(p.data_ptr(), c, t, nh, nw, masks.data_ptr(), out.data_ptr(), int(pad[2]),
                                           int(pad[0]), int(h), int(w), _lib.current_stream_ptr(prob.device)))
    return masks, out


def get_segmentations(processor, rgb=None, device="cuda"):
    """interactions/eval.py:8-24: un-padded per-frame argmax of ``processor.prob`` times 255, uint8 (T,h,w) numpy.

    One fused argmax + un-padding launch over all frames instead of T argmax launches into a float buffer.  With
    more than one object the reference's ``argmax * 255`` leaves the uint8 range (its float -> uint8 cast wraps);
    the same wrap is kept here.  ``rgb`` is only used for its spatial size, as in the reference.
    """
    h, w = (processor.h, processor.w) if rgb is None else tuple(rgb.shape[-2:])
    _, unpadded = argmax_unpad(processor.prob.to(device), processor.pad, h, w)
    return (unpadded * 255).cpu().numpy()   # uint8 arithmetic: (k * 255) mod 256
