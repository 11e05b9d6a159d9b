"""GPU parity of the fused soft aggregation against the reference's outputs (golden) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import memread_np as onp
from tests.helpers import load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["k1", "k3", "k5_odd"])
@pytest.mark.parametrize("keep_bg", [False, True])
@pytest.mark.parametrize("hard", [False, True])
def test_aggregate_golden(name, keep_bg, hard):
    import evavos_b200 as ev
    g = load("aggregate_wbg.npz")
    p = torch.from_numpy(g[f"{name}_prob"]).cuda()
    out = ev.aggregate_wbg(p, keep_bg=keep_bg, hard=hard).cpu().numpy()
    ref = g[f"{name}_bg{int(keep_bg)}_hard{int(hard)}"]
    assert out.shape == ref.shape
    if not hard:
        assert np.abs(out - ref).max() <= 1e-6      # SURVEY.md 8d parity gate
    else:
        # logits * 1000: a 1-ulp logit difference moves a near-tied pixel; everywhere else one-hot agrees
        close = np.abs(out - ref) <= 1e-4
        assert close.mean() > 0.999
        assert np.abs(out.sum(0) - 1).max() < 1e-5 if keep_bg else True


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8, 11])
def test_aggregate_vs_oracle_full_size(k):
    """480x864 (cfg2 aggregate size) and an unaligned size, K up to the generic kernel."""
    import evavos_b200 as ev
    g = torch.Generator().manual_seed(4321 + k)
    for h, w in ((480, 864), (33, 47)):
        p = torch.rand(k, 1, h, w, generator=g)
        out = ev.aggregate_wbg(p.cuda(), keep_bg=True).cpu().numpy()
        ref = onp.aggregate_wbg(p.numpy(), keep_bg=True)
        assert np.abs(out - ref).max() <= 1e-6
        assert np.abs(out.sum(0) - 1).max() < 1e-5
        out2 = ev.aggregate_wbg(p.cuda(), keep_bg=False).cpu().numpy()
        assert np.array_equal(out2, out[1:])


@pytest.mark.parametrize("shape,hw", [((3, 7, 1, 48, 64), (41, 57)), ((2, 5, 1, 32, 48), (32, 48)), ((4, 3, 1, 34, 50), (30, 47))])
def test_argmax_unpad_matches_torch(shape, hw):
    """Fused channel argmax + un-padding == torch.argmax per frame + slicing (inference_core.py:247-257)."""
    import evavos_b200 as ev
    from evavos_b200.tensor_util import pad_divide_by
    c, t, _, nh, nw = shape
    h, w = hw
    g = torch.Generator().manual_seed(c * 100 + t)
    prob = torch.rand(shape, generator=g)
    prob[:, :, :, ::3, ::5] = 0.25                      # exact ties: the first maximal channel must win
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    masks, out = ev.argmax_unpad(prob.cuda(), pad, h, w)
    ref = torch.argmax(prob, dim=0).to(torch.uint8)      # (t,1,nh,nw)
    assert torch.equal(masks.cpu(), ref)
    assert torch.equal(out.cpu(), ref[:, 0, lh:lh + h, lw:lw + w])
    assert out.is_contiguous() and out.dtype == torch.uint8


@pytest.mark.parametrize("k", [1, 2])
def test_get_segmentations_matches_reference_formula(k):
    """interactions/eval.py:8-24 restated with torch ops on the CPU (the module itself needs skimage/torchmetrics)."""
    import types
    import evavos_b200 as ev
    g = torch.Generator().manual_seed(9 + k)
    t, h, w = 5, 100, 140
    nh, nw = 112, 144
    pad = ((nw - w) // 2, nw - w - (nw - w) // 2, (nh - h) // 2, nh - h - (nh - h) // 2)   # (lw, uw, lh, uh)
    prob = torch.rand(k + 1, t, 1, nh, nw, generator=g)
    proc = types.SimpleNamespace(prob=prob.cuda(), pad=pad, t=t, h=h, w=w)
    got = ev.get_segmentations(proc, torch.zeros(3, h, w))
    want = np.zeros((t, h, w), dtype=np.uint8)
    for ti in range(t):
        p = prob[:, ti]
        p = p[:, :, pad[2]:-pad[3], :]
        p = p[:, :, :, pad[0]:-pad[1]]
        want[ti] = ((torch.argmax(p, dim=0)[0].numpy().astype(np.int64) * 255) % 256).astype(np.uint8)
    assert got.dtype == np.uint8 and got.shape == (t, h, w)
    assert (got == want).all()
