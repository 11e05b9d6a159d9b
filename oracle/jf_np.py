"""numpy + cv2 restatement of the reference's per-frame J and J&F quality (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/interactions/metrics.py:
  compute_iou      :9-20    smoothed IoU (intersection + 1e-6) / (union + 1e-6)
  get_j_and_f      :24-36   0.5 * binary Jaccard (torchmetrics JaccardIndex(task="binary")) + 0.5 * f_measure
  _seg2bmap        :40-97   one-pixel boundary map (davisinteractive)
  f_measure        :100-160 boundary precision / recall after dilating with a disk of ceil(0.008 * |shape|) pixels
and interactions/eval.py:27-81 for which frames are scored.  Third-party pieces the reference imports and that are
absent offline are restated from their published definitions:
  skimage.morphology.disk(r)   (2r+1)^2 footprint of x^2 + y^2 <= r^2
  torchmetrics binary Jaccard  TP / (TP + FP + FN) (0 when the union is empty)
oracle/make_golden_jf.py executes the reference's own metrics.py (with exactly these two stubbed) and asserts that
this file reproduces it bit for bit; tests/test_oracle_golden.py repeats the check against the committed vectors.
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

SMOOTH = 1e-6


def disk(radius: int) -> np.ndarray:
    """skimage.morphology.disk: uint8 (2r+1, 2r+1) footprint."""
    r = int(radius)
    ax = np.arange(-r, r + 1)
    xx, yy = np.meshgrid(ax, ax)
    return (xx * xx + yy * yy <= r * r).astype(np.uint8)


def bound_pixels(shape, bound_th: float = 0.008) -> int:
    return int(bound_th if bound_th >= 1 else math.ceil(bound_th * np.linalg.norm(shape)))


def seg2bmap(seg: np.ndarray) -> np.ndarray:
    """metrics.py:40-97 at full resolution (width/height None)."""
    seg = seg.astype(bool)
    e = np.zeros_like(seg)
    s = np.zeros_like(seg)
    se = np.zeros_like(seg)
    e[:, :-1] = seg[:, 1:]
    s[:-1, :] = seg[1:, :]
    se[:-1, :-1] = seg[1:, 1:]
    b = seg ^ e | seg ^ s | seg ^ se
    b[-1, :] = seg[-1, :] ^ e[-1, :]
    b[:, -1] = seg[:, -1] ^ s[:, -1]
    b[-1, -1] = 0
    return b


def f_measure(true_mask: np.ndarray, pred_mask: np.ndarray, bound_th: float = 0.008) -> float:
    true_mask = np.asarray(true_mask, dtype=bool)
    pred_mask = np.asarray(pred_mask, dtype=bool)
    bp = bound_pixels(true_mask.shape, bound_th)
    fg_b, gt_b = seg2bmap(pred_mask), seg2bmap(true_mask)
    fg_dil = cv2.dilate(fg_b.astype(np.uint8), disk(bp))
    gt_dil = cv2.dilate(gt_b.astype(np.uint8), disk(bp))
    gt_match, fg_match = gt_b * fg_dil, fg_b * gt_dil
    n_fg, n_gt = np.sum(fg_b), np.sum(gt_b)
    if n_fg == 0 and n_gt > 0:
        precision, recall = 1, 0
    elif n_fg > 0 and n_gt == 0:
        precision, recall = 0, 1
    elif n_fg == 0 and n_gt == 0:
        precision, recall = 1, 1
    else:
        precision, recall = np.sum(fg_match) / float(n_fg), np.sum(gt_match) / float(n_gt)
    return 0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)


def compute_iou(pred: np.ndarray, gt: np.ndarray) -> float:
    """metrics.py:9-20 for one frame (fp32 arithmetic like the torch original)."""
    inter = np.float32(np.logical_and(pred, gt).sum())
    union = np.float32(np.logical_or(pred, gt).sum())
    return float((inter + np.float32(SMOOTH)) / (union + np.float32(SMOOTH)))


def binary_jaccard(a: np.ndarray, b: np.ndarray) -> float:
    inter = np.logical_and(a, b).sum()
    union = np.logical_or(a, b).sum()
    return float(np.float32(inter) / np.float32(union)) if union > 0 else 0.0


def j_and_f(pred: np.ndarray, gt: np.ndarray) -> float:
    """get_j_and_f(pred, gt) as eval.py:72 calls it (its first parameter, named gt_mask there, receives pred)."""
    return binary_jaccard(pred, gt) * 0.5 + f_measure(pred, gt) * 0.5


def frame_qualities(pred: np.ndarray, gt: np.ndarray, metric: str = "j"):
    """eval.py:52-79 without the interaction overrides: (T,h,w) bool masks -> (frame_quality, frame_quality_all);
    frames with an empty ground truth get the token 20 and are left out of frame_quality."""
    fq, fq_all = [], []
    for p, g in zip(pred.astype(bool), gt.astype(bool)):
        if not g.any():
            fq_all.append(20)
            continue
        q = compute_iou(p, g) if metric == "j" else j_and_f(p, g)
        fq.append(q)
        fq_all.append(q)
    return fq, fq_all
