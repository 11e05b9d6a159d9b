"""GPU: per-frame J / J&F on the device against vectors produced by the reference's own
interactions/metrics.py (oracle/make_golden_jf.py) and against the numpy + cv2 restatement on random masks."""
import numpy as np
import pytest
import torch

from oracle import jf_np
from tests.helpers import load

pytestmark = pytest.mark.gpu


def _case(g, name):
    shape = tuple(int(x) for x in g[f"{name}_shape"])
    pred = np.unpackbits(g[f"{name}_pred"], axis=-1)[..., :shape[2]].astype(bool)
    gt = np.unpackbits(g[f"{name}_gt"], axis=-1)[..., :shape[2]].astype(bool)
    return pred, gt


@pytest.mark.parametrize("name", ["small", "davis", "odd"])
def test_frame_metrics_match_reference_golden(name):
    import evavos_b200 as ev
    g = load("jf.npz")
    pred, gt = _case(g, name)
    dev = torch.device("cuda:0")
    res = ev.frame_metrics(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev))
    j, jf, f = (res[k].cpu().numpy() for k in ("j", "j_and_f", "f"))
    empty = res["gt_empty"].cpu().numpy()
    assert (empty == ~gt.reshape(gt.shape[0], -1).any(1)).all()
    assert np.abs(j - g[f"{name}_j"]).max() <= 1e-6
    live = ~empty
    assert np.abs(jf[live] - g[f"{name}_jf"][live]).max() <= 1e-6      # |dJ|, |dF| <= 1e-6 (in fact bit-equal)
    assert np.abs(f[live] - g[f"{name}_f"][live]).max() <= 1e-6
    assert (g[f"{name}_jf"][empty] == 20).all()


def test_frame_metrics_random_masks_against_restatement():
    import evavos_b200 as ev
    rng = np.random.default_rng(7)
    dev = torch.device("cuda:0")
    for h, w in ((33, 65), (128, 96), (64, 31)):
        t = 7
        gt = rng.random((t, h, w)) > 0.6
        pred = gt ^ (rng.random((t, h, w)) > 0.9)
        pred[0] = False
        res = ev.frame_metrics(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev))
        for f in range(t):
            assert abs(res["f"][f].item() - jf_np.f_measure(pred[f], gt[f])) <= 1e-12
            assert abs(res["j_and_f"][f].item() - jf_np.j_and_f(pred[f], gt[f])) <= 1e-12
            assert abs(res["j"][f].item() - jf_np.compute_iou(pred[f], gt[f])) <= 1e-7


def test_eval_processor_metric_matches_eval_py():
    """The eval.py:27-81 flow: argmax of processor.prob, un-padding, interaction overrides, empty-gt token."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    t, h, w = 6, 50, 70
    nh, nw = 64, 80
    pad = ((nw - w) // 2, nw - w - (nw - w) // 2, (nh - h) // 2, nh - h - (nh - h) // 2)

    class Proc:      # only what eval.py reads: prob, pad, t
        pass
    proc = Proc()
    proc.prob, proc.pad, proc.t = torch.rand(2, t, 1, nh, nw, generator=g).to(dev), pad, t
    gt = (torch.rand(t, h, w, generator=g) > 0.5)
    gt[4] = False
    sam = {2: torch.rand(h, w, generator=g) > 0.4}
    data = {"gt": gt[None, :, None].float(), "rgb": torch.zeros(1, t, 3, h, w)}
    inter, kinds = [1, 2], [0, 1, 2, 0, 0, 0]
    for metric in ("j", "j_and_f"):
        mean, gen, fq, fq_all = ev.eval_processor_metric(proc, data, inter, kinds, masks_from_sam=sam, metric=metric)
        pred = proc.prob[:, :, 0, pad[2]:nh - pad[3], pad[0]:nw - pad[1]].argmax(0).bool().cpu().numpy()
        pred[1] = gt[1].numpy()
        pred[2] = sam[2].numpy()
        ref_fq, ref_all = jf_np.frame_qualities(pred, gt.numpy(), metric)
        assert len(fq) == len(ref_fq) == t - 1 and fq_all[4] == 20 and ref_all[4] == 20
        assert np.abs(np.array(fq) - np.array(ref_fq)).max() <= 1e-6
        assert abs(mean - np.mean(ref_fq)) <= 1e-6
        assert gen.shape == (t, h, w) and (gen[1] == gt[1].numpy()).all() and (gen[2] == sam[2].numpy()).all()
