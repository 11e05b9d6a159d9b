"""Golden vectors for the J / J&F metric, produced by the reference's OWN interactions/metrics.py.

    PYTHONPATH=/root/reference python oracle/make_golden_jf.py   -> tests/golden/jf.npz

skimage and torchmetrics are not installed offline, and metrics.py imports `disk` and `JaccardIndex` from them at
module level.  Exactly those two names are provided by stub modules (oracle.jf_np.disk, a binary Jaccard); every
other line that runs - compute_iou, get_j_and_f, _seg2bmap, f_measure - is the reference's, unmodified.  The script
also asserts that oracle/jf_np.py reproduces it bit for bit.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jf_np  # noqa: E402

REF = os.environ.get("EVAVOS_REFERENCE", "/root/reference")


def load_reference_metrics():
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.morphology")
    skm.disk = jf_np.disk
    sk.morphology = skm
    tm = types.ModuleType("torchmetrics")

    class JaccardIndex:      # torchmetrics.JaccardIndex(task="binary"): TP / (TP + FP + FN)
        def __init__(self, task="binary", num_classes=2):
            assert task == "binary"

        def __call__(self, a, b):
            a, b = a.bool(), b.bool()
            inter, union = (a & b).sum(), (a | b).sum()
            return (inter.float() / union.float()) if union > 0 else torch.zeros(())

    tm.JaccardIndex = JaccardIndex
    sys.modules.update({"skimage": sk, "skimage.morphology": skm, "torchmetrics": tm})
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "interactions", "metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def blobs(rng, t, h, w, n=3):
    """Smooth random blob masks: sums of a few Gaussians thresholded."""
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.zeros((t, h, w), dtype=bool)
    for f in range(t):
        acc = np.zeros((h, w))
        for _ in range(n):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            sy, sx = rng.uniform(h / 12, h / 4), rng.uniform(w / 12, w / 4)
            acc += np.exp(-((yy - cy) ** 2 / (2 * sy * sy) + (xx - cx) ** 2 / (2 * sx * sx)))
        out[f] = acc > rng.uniform(0.35, 0.8)
    return out


def main():
    ref = load_reference_metrics()
    rng = np.random.default_rng(2024)
    cases = {}
    for name, (t, h, w) in {"small": (6, 60, 107), "davis": (4, 480, 854), "odd": (5, 97, 41)}.items():
        gt = blobs(rng, t, h, w)
        pred = gt.copy()
        for f in range(t):      # perturb: shift + noise blobs, so boundaries are near but not on each other
            pred[f] = np.roll(gt[f], (rng.integers(-4, 5), rng.integers(-4, 5)), (0, 1)) ^ (blobs(rng, 1, h, w, 1)[0] & (rng.random((h, w)) > 0.7))
        if name == "small":
            gt[1] = False                     # empty ground truth -> token 20, frame skipped (eval.py:60-63)
            pred[2] = False                   # empty prediction, non-empty gt -> J = 0, F = 0
            pred[3] = gt[3]                   # perfect frame
            gt[4], pred[4] = True, True       # full frames: no boundary at all -> precision = recall = 1
            gt[4, 0, 0] = False
        j = np.array([ref.compute_iou(torch.from_numpy(pred[f:f + 1]), torch.from_numpy(gt[f:f + 1])) for f in range(t)])
        jf = np.full(t, 20.0)
        fm = np.full(t, 20.0)
        for f in range(t):
            if gt[f].any():
                jf[f] = ref.get_j_and_f(torch.from_numpy(pred[f:f + 1]), torch.from_numpy(gt[f:f + 1]))
                fm[f] = ref.f_measure(pred[f], gt[f])
                assert jf[f] == jf_np.j_and_f(pred[f], gt[f]), (name, f)
                assert fm[f] == jf_np.f_measure(pred[f], gt[f])
            assert j[f] == jf_np.compute_iou(pred[f], gt[f])
            assert (ref._seg2bmap(pred[f]) == jf_np.seg2bmap(pred[f])).all()
        cases[name] = (pred, gt, j, jf, fm)
    out = {}
    for name, (pred, gt, j, jf, fm) in cases.items():
        out[f"{name}_pred"] = np.packbits(pred, axis=-1)
        out[f"{name}_gt"] = np.packbits(gt, axis=-1)
        out[f"{name}_shape"] = np.array(pred.shape)
        out[f"{name}_j"], out[f"{name}_jf"], out[f"{name}_f"] = j, jf, fm
    path = os.path.join(ROOT, "tests", "golden", "jf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_j")})


if __name__ == "__main__":
    main()
