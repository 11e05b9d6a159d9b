"""Fused attention read of the fusion path (evavos_attention_readout) against the reference's outputs and the oracle."""
import numpy as np
import pytest
import torch

import evavos_b200 as ev
from oracle import memread_np as onp
from tests.helpers import load

pytestmark = pytest.mark.gpu

# fp32 everywhere; the kernel sums the softmax in a different order than torch (online, split over the memory
# axis), so agreement is to rounding: attention maps are convex combinations of mask values in [0, 1].
ATOL = 2e-6


@pytest.fixture(params=["auto", "simt", "tensor"])
def form(request, monkeypatch):
    """Both forms of the partial pass on every shape: CUDA cores (fp32) and tensor cores (3 x TF32 mma.sync); "auto" is
    the size rule of csrc/attention.cu."""
    if request.param == "auto":
        monkeypatch.delenv("EVAVOS_ATTENTION_PATH", raising=False)
    else:
        monkeypatch.setenv("EVAVOS_ATTENTION_PATH", request.param)
    return request.param


@pytest.mark.parametrize("name", ["b2", "b4_peaky", "b1_flat"])
def test_get_attention_golden(name, form):
    """PropagationNetwork.get_attention on the GPU vs the reference's get_attention (tests/golden/attention.npz)."""
    g = load("attention.npz")
    dev = torch.device("cuda:0")
    mk, qk = torch.from_numpy(g[f"{name}_mk"]).to(dev), torch.from_numpy(g[f"{name}_qk"]).to(dev)
    pos, neg = torch.from_numpy(g[f"{name}_pos"]).to(dev), torch.from_numpy(g[f"{name}_neg"]).to(dev)
    b, _, h, w = pos.shape
    nh, nw = h // 16, w // 16
    vec = torch.cat([torch.nn.functional.interpolate(pos, size=(nh, nw), mode="area").view(b, 1, -1),
                     torch.nn.functional.interpolate(neg, size=(nh, nw), mode="area").view(b, 1, -1)], 1)
    low = ev.attention_readout(mk, qk, vec.view(2 * b, -1)).view(b, 2, nh, nw)
    assert np.abs(low.cpu().numpy() - g[f"{name}_lowres"]).max() < ATOL
    # the method itself (bilinear upsampling included); unbound call: it only needs tensors
    attn = ev.PropagationNetwork.get_attention(None, mk, pos, neg, qk)
    assert attn.shape == (b, 2, h, w)
    assert np.abs(attn.cpu().numpy() - g[f"{name}_attn"]).max() < ATOL


@pytest.mark.parametrize("shape,n_vec,scale", [((30, 54), 8, 1.0), ((30, 54), 2, 3.0), ((68, 120), 12, 1.0),
                                               ((5, 7), 33 - 1, 1.0), ((1, 3), 1, 1.0), ((48, 90), 5, 1.5),
                                               ((12, 20), 3, 2.0e4), ((12, 20), 3, 1.0e-3)])   # beyond fp16 range / deep inside its
                                                                                           # subnormals: the tensor form normalises
def test_attention_vs_oracle(shape, n_vec, scale, form):
    """Full sizes (480p: 1620 x 1620; 1080p: 8160 x 8160), ragged tiny grids, up to the 32-row limit."""
    h, w = shape
    g = torch.Generator().manual_seed(77 + h)
    mk = torch.randn(1, 64, 1, h, w, generator=g) * scale
    qk = torch.randn(1, 64, h, w, generator=g) * scale
    vec = torch.rand(n_vec, h * w, generator=g)
    out = ev.attention_readout(mk.cuda(), qk.cuda(), vec.cuda()).cpu().numpy()
    want = onp.attention_readout(mk.reshape(64, -1).numpy(), qk.reshape(64, -1).numpy(), vec.numpy())
    assert out.shape == want.shape and np.isfinite(out).all()
    # fp32 scores carry an absolute rounding error of ~eps * |score|, which the softmax turns into a relative error
    # of the weights: the tolerance grows with the score magnitude (the reference's own fp32 path does the same)
    smax = np.abs(onp.affinity_scores(mk.reshape(64, -1).numpy(), qk.reshape(64, -1).numpy())).max()
    assert np.abs(out - want).max() < max(ATOL, 1.2e-7 * smax)
    # rows of ones must come back as ones (the weights of every query sum to 1)
    ones = ev.attention_readout(mk.cuda(), qk.cuda(), torch.ones(1, h * w).cuda()).cpu().numpy()
    assert np.abs(ones - 1).max() < 1e-6


def test_attention_strided_inputs_and_errors(form):
    """Channel-strided views (a frame sliced out of a (1,CK,T,H,W) bank) are read in place; bad shapes fail loudly."""
    g = torch.Generator().manual_seed(5)
    bank = torch.randn(1, 64, 3, 6, 9, generator=g).cuda()
    qk = torch.randn(1, 64, 6, 9, generator=g).cuda()
    vec = torch.rand(4, 54, generator=g).cuda()
    out = ev.attention_readout(bank[:, :, 1:2], qk, vec).cpu().numpy()
    want = onp.attention_readout(bank[0, :, 1].reshape(64, -1).cpu().numpy(), qk.reshape(64, -1).cpu().numpy(),
                                 vec.cpu().numpy())
    assert np.abs(out - want).max() < ATOL
    with pytest.raises(ev.EvavosError):
        ev.attention_readout(torch.randn(1, 32, 1, 6, 9).cuda(), torch.randn(1, 32, 6, 9).cuda(), vec)  # CK != 64
    with pytest.raises(ValueError):
        ev.attention_readout(bank[:, :, 1:2], qk, vec[:, :50])
    with pytest.raises(RuntimeError):
        ev.attention_readout(bank[:, :, 1:2].cpu(), qk.cpu(), vec.cpu())
