"""GPU: the decoder's fused elementwise tails (csrc/decoder_ops.cu) against the PyTorch ops they replace."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(*shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("shape", [(3, 256, 12, 20), (1, 8, 2, 2), (5, 512, 30, 54)])
def test_bias_residual(dtype, tol, shape):
    from evavos_b200.decoder_ops import bias_residual_
    y, r = _cl(*shape, dtype=dtype, seed=1), _cl(*shape, dtype=dtype, seed=2)
    bias = torch.randn(shape[1], device="cuda")
    for res in (None, r):
        for relu in (False, True):
            want = y.float() + bias.view(1, -1, 1, 1) + (0 if res is None else res.float())
            want = torch.relu(want) if relu else want
            got = bias_residual_(y.clone(memory_format=torch.preserve_format), bias, res, relu=relu)
            assert got.dtype == dtype and got.is_contiguous(memory_format=torch.channels_last)
            assert (got.float() - want).abs().max().item() <= tol * (1 + want.abs().max().item())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("shape", [(2, 256, 12, 20), (1, 8, 2, 2), (3, 512, 60, 108), (1, 16, 6, 2)])
def test_upsample2x_add(dtype, tol, shape):
    from evavos_b200.decoder_ops import upsample2x_add_
    n, c, h, w = shape
    y, x = _cl(n, c, h, w, dtype=dtype, seed=3), _cl(n, c, h // 2, w // 2, dtype=dtype, seed=4)
    bias = torch.randn(c, device="cuda")
    want = y.float() + bias.view(1, -1, 1, 1) + F.interpolate(x.float(), scale_factor=2, mode="bilinear", align_corners=False)
    got = upsample2x_add_(y.clone(memory_format=torch.preserve_format), bias, x)
    assert (got.float() - want).abs().max().item() <= tol * (1 + want.abs().max().item())


def test_argument_checks():
    from evavos_b200._lib import EvavosError
    from evavos_b200.decoder_ops import bias_residual_, upsample2x_add_
    y = _cl(1, 8, 4, 4, dtype=torch.float32, seed=5)
    with pytest.raises(ValueError):
        bias_residual_(y.contiguous(), torch.zeros(8, device="cuda"))                  # NCHW
    with pytest.raises(ValueError):
        upsample2x_add_(y, torch.zeros(8, device="cuda"), _cl(1, 8, 4, 4, dtype=torch.float32, seed=6))
    with pytest.raises(EvavosError):
        bias_residual_(_cl(1, 6, 4, 4, dtype=torch.float32, seed=7), torch.zeros(6, device="cuda"))   # C % 4


@pytest.mark.parametrize("dtype,tol,k", [(torch.float32, 1e-4, 1), (torch.float32, 1e-4, 3), (torch.bfloat16, 4e-2, 2)])
def test_fused_decoder_matches_decode_input(dtype, tol, k):
    """conv_opt.FusedDecoder against PropagationNetwork.decode_input (same weights; fp32 convolutions without TF32)."""
    import evavos_b200 as ev
    from evavos_b200.conv_opt import FusedDecoder
    from evavos_b200.networks import seeded_init
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            prop = ev.PropagationNetwork().eval().cuda()
            seeded_init(prop, 1001)
            f, hh, ww = 2, 6, 10
            g = torch.Generator(device="cuda").manual_seed(11)
            m4 = torch.randn(f, k, 1024, hh, ww, device="cuda", generator=g)
            qf8 = torch.randn(f, 512, 2 * hh, 2 * ww, device="cuda", generator=g)
            qf4 = torch.randn(f, 256, 4 * hh, 4 * ww, device="cuda", generator=g)
            want = prop.decode_input(m4, qf8, qf4)
            got = FusedDecoder(prop.decoder, dtype)(m4, qf8, qf4)
            assert got.shape == want.shape and got.dtype == torch.float32
            assert (got - want).abs().max().item() < tol
    finally:
        torch.backends.cudnn.allow_tf32 = old
