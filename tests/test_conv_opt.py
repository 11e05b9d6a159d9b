"""Host-side checks of the conv-pass rewrites around the memory read (evavos_b200/conv_opt.py, staging.py):
BatchNorm folding is the same function, caches follow the parameters, deep copies do not drag caches along."""
import copy

import pytest
import torch

import evavos_b200 as ev
from evavos_b200.conv_opt import ConvPasses, conv_passes, fold_batchnorm, folded_encoders
from evavos_b200.networks import seeded_init


@pytest.fixture(scope="module")
def prop():
    torch.set_grad_enabled(False)
    net = ev.PropagationNetwork().eval()
    seeded_init(net, 1001)
    yield net
    torch.set_grad_enabled(True)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def test_folded_encoders_compute_the_same_function(prop):
    x = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(1))
    ke, ve = folded_encoders(prop, channels_last=False)
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in ke.modules())
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in ve.modules())
    ref = prop.key_encoder(x)
    got = ke(x)
    assert all(_rel(g, r) < 1e-5 for g, r in zip(got, ref))
    masks = torch.rand(3, 1, 64, 96, generator=torch.Generator().manual_seed(2))
    v_ref = prop.encode_value(x[:1], ref[0][:1], masks)
    v_got = prop.encode_value(x[:1], ref[0][:1], masks, ve)
    assert v_got.shape == v_ref.shape == (3, 512, 1, 4, 6) and _rel(v_got, v_ref) < 1e-5
    # the caller's network is untouched: same state dict keys, BatchNorms still there
    assert len(prop.state_dict()) == 405 and any(isinstance(m, torch.nn.BatchNorm2d) for m in prop.modules())


def test_fold_needs_eval_mode():
    net = ev.PropagationNetwork()
    with pytest.raises(RuntimeError, match="eval"):
        fold_batchnorm(net.key_encoder.train())


def test_caches_follow_the_parameters_and_stay_out_of_deep_copies(prop):
    a = folded_encoders(prop, False)
    assert folded_encoders(prop, False)[0] is a[0]
    p1 = conv_passes(prop, False, False, True, False)
    assert conv_passes(prop, False, False, True, False) is p1 and isinstance(p1, ConvPasses)
    assert conv_passes(prop, False, False, False, False) is not p1          # other options: other passes
    prop.key_encoder.conv1.weight.mul_(1.0)                                 # an in-place write bumps the version
    assert folded_encoders(prop, False)[0] is not a[0]
    assert conv_passes(prop, False, False, True, False) is not p1
    clone = copy.deepcopy(prop)                                             # policies deep-copy whole processors
    assert clone.__dict__["_evavos_folded"].value is None and clone.__dict__["_evavos_passes"].value is None
    assert len(clone.state_dict()) == 405
    assert conv_passes(clone, False, False, True, False) is not conv_passes(prop, False, False, True, False)


def test_eager_passes_match_the_network(prop):
    x = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(3))
    passes = conv_passes(prop, False, False, True, False)
    ref = prop.encode_key(x)
    got = passes.encode_key(x)
    assert all(_rel(g, r) < 1e-5 for g, r in zip(got, ref))
    k16, v16, f16, f8, f4 = ref
    m4 = torch.rand(2, 2, 1024, 4, 6, generator=torch.Generator().manual_seed(4))
    assert _rel(passes.decode(m4, f8, f4), prop.decode_input(m4, f8, f4)) < 1e-5


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_other_object_masks_match_the_reference_formulation(prop, k):
    """encode_value sums every OTHER object's mask with a cached index tensor (no boolean indexing: it would synchronise
    and cannot be captured in a CUDA graph); same numbers as the reference's per-object gather (prop_net.py:158-163)."""
    masks = torch.rand(k, 1, 32, 48, generator=torch.Generator().manual_seed(10 + k))
    if k == 1:
        expect = torch.zeros_like(masks)
    else:
        expect = torch.cat([torch.sum(masks[[j for j in range(k) if j != i]], dim=0, keepdim=True) for i in range(k)], 0)
    seen = {}

    class Spy(torch.nn.Module):
        def forward(self, frame, kf16, m, others):
            seen["others"] = others
            return torch.zeros(k, 512, 2, 3)

    prop.encode_value(torch.rand(1, 3, 32, 48), torch.rand(1, 1024, 2, 3), masks, Spy())
    assert torch.equal(seen["others"], expect)


def test_staging_falls_back_for_small_and_cpu_tensors():
    from evavos_b200.staging import download_numpy
    t = torch.arange(12, dtype=torch.uint8).view(3, 4)
    out = download_numpy(t)
    assert out.shape == (3, 4) and (out == t.numpy()).all()


@pytest.mark.parametrize("k", [1, 3])
def test_fused_decoder_is_the_same_graph(prop, k, monkeypatch):
    """conv_opt.FusedDecoder restructures decode_input (bias-free convolutions + two in-place tails); with torch doubles
    for the two CUDA kernels (the GPU suite checks the kernels themselves) it is the same function on CPU."""
    import torch.nn.functional as F

    from evavos_b200 import decoder_ops
    from evavos_b200.conv_opt import FusedDecoder, fused_decoder

    def bias_residual_(y, bias, residual=None, relu=False):
        assert y.is_contiguous(memory_format=torch.channels_last) and bias.dtype == torch.float32
        y += bias.view(1, -1, 1, 1)
        if residual is not None:
            y += residual
        return y.relu_() if relu else y

    def upsample2x_add_(y, bias, x):
        assert y.is_contiguous(memory_format=torch.channels_last) and x.is_contiguous(memory_format=torch.channels_last)
        y += bias.view(1, -1, 1, 1) + F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        return y

    monkeypatch.setattr(decoder_ops, "bias_residual_", bias_residual_)
    monkeypatch.setattr(decoder_ops, "upsample2x_add_", upsample2x_add_)
    g = torch.Generator().manual_seed(5)
    f, hh, ww = 2, 3, 5
    m4 = torch.randn(f, k, 1024, hh, ww, generator=g)
    qf8, qf4 = torch.randn(f, 512, 2 * hh, 2 * ww, generator=g), torch.randn(f, 256, 4 * hh, 4 * ww, generator=g)
    want = prop.decode_input(m4, qf8, qf4)
    fd = FusedDecoder(prop.decoder, torch.float32)
    got = fd(m4, qf8, qf4)
    assert got.shape == want.shape == (f, k, 1, 16 * hh, 16 * ww)
    assert (got - want).abs().max().item() < 1e-5
    assert not fd.fused                                             # no cuDNN on the CPU: the unfused conv + ReLU form
    assert fused_decoder(prop) is fused_decoder(prop)               # cached on the network ...
    assert "_evavos_decoder" not in prop.state_dict() and len(prop.state_dict()) == 405
    assert copy.deepcopy(prop).__dict__["_evavos_decoder"].value is None    # ... and not copied with it


def test_decoder_tail_kernels_have_no_cpu_path():
    """evavos_b200.decoder_ops fails loudly on CPU tensors (north_star: no CPU fallback in the product package)."""
    from evavos_b200.decoder_ops import bias_residual_, upsample2x_add_
    y = torch.zeros(1, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        bias_residual_(y, torch.zeros(8))
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        upsample2x_add_(y, torch.zeros(8), torch.zeros(1, 8, 2, 2).contiguous(memory_format=torch.channels_last))
