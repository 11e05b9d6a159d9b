"""Per-frame J / J&F quality on the GPU with the reference's interfaces.

``eval_processor_metric`` mirrors interactions/eval.py:27-81 (same arguments, same 4-tuple); ``get_j_and_f`` /
``compute_iou`` / ``f_measure`` mirror interactions/metrics.py:9-36, 100-160.  The reference scores every annotation
round on the CPU, one frame at a time (argmax -> D2H -> numpy / cv2: two dilations per frame); here the masks of all
frames stay on the device and ``evavos_jf_metrics`` produces every frame's numbers in three launches.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .aggregate import argmax_unpad
from .memory_reader import _workspace


def bound_pixels(h: int, w: int, bound_th: float = 0.008) -> int:
    """metrics.py:120-121: the dilation radius of the boundary F-measure."""
    return int(bound_th if bound_th >= 1 else math.ceil(bound_th * math.hypot(h, w)))


def frame_metrics(pred: torch.Tensor, gt: torch.Tensor, bound_th: float = 0.008):
    """pred, gt: (T,h,w) CUDA tensors (bool / uint8 / float; non-zero = foreground).

    Returns a dict of (T,) tensors on the device: ``j`` (smoothed IoU of compute_iou), ``jaccard``, ``f``,
    ``j_and_f`` (fp64) and ``gt_empty`` (bool).
    """
    lib = _lib.load()
    if not (pred.is_cuda and gt.is_cuda):
        raise RuntimeError("frame_metrics: CUDA tensors required (no CPU path in evavos_b200)")
    if pred.dim() != 3 or pred.shape != gt.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and gt {tuple(gt.shape)} must both be (T,h,w)")
    dev = pred.device
    p = (pred != 0).to(torch.uint8).contiguous()
    g = (gt.to(dev) != 0).to(torch.uint8).contiguous()
    t, h, w = p.shape
    out = torch.empty((t, 4), dtype=torch.float64, device=dev)
    empty = torch.empty((t,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        need = lib.evavos_jf_workspace_bytes(t, h, w)
        ws = _workspace.get(dev, int(need))
        _lib.check(lib.evavos_jf_metrics(p.data_ptr(), g.data_ptr(), t, h, w, bound_pixels(h, w, bound_th), ws.data_ptr(),
                                         ws.numel(), out.data_ptr(), empty.data_ptr(), _lib.current_stream_ptr(dev)))
    return {"j": out[:, 0], "jaccard": out[:, 1], "f": out[:, 2], "j_and_f": out[:, 3], "gt_empty": empty.bool()}


def compute_iou(outputs: torch.Tensor, labels: torch.Tensor) -> float:
    """metrics.py:9-20: mean smoothed IoU over the leading axis of two (N,h,w) boolean tensors."""
    assert outputs.ndim == labels.ndim == 3
    return float(frame_metrics(outputs, labels)["j"].to(torch.float32).mean().item())


def f_measure(true_mask, pred_mask, bound_th: float = 0.008) -> float:
    """metrics.py:100-160 for one (h,w) pair."""
    tm, pm = torch.as_tensor(true_mask), torch.as_tensor(pred_mask)
    dev = tm.device if tm.is_cuda else (pm.device if pm.is_cuda else torch.device("cuda"))
    return float(frame_metrics(pm.to(dev)[None], tm.to(dev)[None], bound_th)["f"].item())


def get_j_and_f(gt_mask: torch.Tensor, pred_mask: torch.Tensor) -> float:
    """metrics.py:24-36 for one (1,h,w) pair (both measures are symmetric in their arguments)."""
    assert gt_mask.ndim == pred_mask.ndim == 3
    dev = gt_mask.device if gt_mask.is_cuda else (pred_mask.device if pred_mask.is_cuda else torch.device("cuda"))
    return float(frame_metrics(pred_mask.to(dev), gt_mask.to(dev))["j_and_f"].item())


def eval_processor_metric(processor, data, interacted_frames, frame_iteraction_type, masks_from_sam=None, metric="j",
                          device="cuda"):
    """interactions/eval.py:27-81.  ``frame_iteraction_type[f]``: 0 no interaction, 1 ground-truth mask, 2 click/bbox
    (the SAM mask of that frame replaces the propagated one).  Returns
    (mean quality over frames with a non-empty ground truth, gen_masks (T,h,w) float64 numpy, frame_quality,
    frame_quality_all) - frames with an empty ground truth carry the token 20 in frame_quality_all."""
    assert metric in {"j", "j_and_f"}
    dev = torch.device(device)
    gt = data["gt"].squeeze().to(dev)                 # (T,h,w)
    if gt.dim() == 2:
        gt = gt[None]
    h, w = gt.shape[-2:]
    _, unpadded = argmax_unpad(processor.prob.to(dev), processor.pad, h, w)     # (T,h,w) uint8 object ids
    out_masks = unpadded * 255                        # uint8 wrap for more than one object, as the reference's cast
    pred = out_masks != 0
    gen = out_masks.to(torch.float64) / 255
    gtb = gt != 0
    for f in interacted_frames:
        kind = frame_iteraction_type[f]
        if kind == 1:
            pred[f] = gtb[f]
            gen[f] = gt[f].to(torch.float64)
        elif kind == 2:
            m = masks_from_sam[f].to(dev).bool()
            pred[f] = m
            gen[f] = m.to(torch.float64)
    res = frame_metrics(pred, gtb)
    q = (res["j"].to(torch.float32).to(torch.float64) if metric == "j" else res["j_and_f"]).cpu().numpy()
    empty = res["gt_empty"].cpu().numpy()
    frame_quality = [float(v) for v, e in zip(q, empty) if not e]
    frame_quality_all = [20 if e else float(v) for v, e in zip(q, empty)]
    return np.mean(np.array(frame_quality)), gen.cpu().numpy(), frame_quality, frame_quality_all
