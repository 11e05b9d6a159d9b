"""GPU: InferenceCore.interact end to end against the reference's outputs (seeded random weights)."""
import copy

import numpy as np
import pytest
import torch

from tests.helpers import load

pytestmark = pytest.mark.gpu


def _nets():
    import evavos_b200 as ev
    from evavos_b200.networks import seeded_init
    prop, fuse = ev.PropagationNetwork().eval(), ev.FusionNet().eval()
    seeded_init(prop, 1001)
    seeded_init(fuse, 1002)
    return prop, fuse


@pytest.fixture(scope="module", autouse=True)
def _fp32_convs():
    # the golden vectors were produced with fp32 convolutions on the CPU; keep cuDNN out of TF32
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    yield
    torch.backends.cudnn.allow_tf32 = old
    torch.set_grad_enabled(True)


@pytest.mark.parametrize("tag", ["e2e_k1", "e2e_k2"])
def test_interact_matches_reference(tag):
    import evavos_b200 as ev
    g = load(f"{tag}.npz")
    prop, fuse = _nets()
    images = torch.from_numpy(g["images"])
    k = int(g["num_objects"])
    proc = ev.InferenceCore(prop, fuse, images, k, mem_freq=int(g["mem_freq"]), device="cuda:0")
    assert tuple(proc.pad) == tuple(int(x) for x in g["pad"])
    for n in range(int(g["n_interactions"])):
        mask = torch.from_numpy(g[f"mask_{n}"])
        out = proc.interact(mask, int(g[f"frame_{n}"]), scribble=bool(g[f"scribble_{n}"]))
        ref_prob, ref_masks = g[f"prob_{n}"], g[f"np_masks_{n}"]
        prob = proc.prob.cpu().numpy()
        assert prob.shape == ref_prob.shape and out.shape == ref_masks.shape and out.dtype == np.uint8
        # conv stacks run on different hardware (cuDNN fp32 vs MKL fp32): allow 2e-3 on probabilities
        assert np.abs(prob - ref_prob).max() < 2e-3, np.abs(prob - ref_prob).max()
        # masks may only differ where the reference itself is undecided (two channels within 4e-3)
        srt = np.sort(ref_prob, 0)
        undecided = (srt[-1] - srt[-2] < 4e-3)[:, 0]
        lw, uw, lh, uh = (int(x) for x in g["pad"])
        undecided = undecided[:, lh:undecided.shape[1] - uh or None, lw:undecided.shape[2] - uw or None]
        diff = out != ref_masks
        assert not (diff & ~undecided).any()
        assert diff.mean() < 1e-3        # mask IoU parity gate: >= 0.999 agreement
    # the state the callers read
    assert proc.certain_mem_k.shape[2] == int(g["n_interactions"]) and proc.certain_mem_v.shape[0] == k
    # policies deep-copy the processor (interactions/policies.py:103)
    clone = copy.deepcopy(proc)
    assert torch.equal(clone.prob, proc.prob) and clone.certain_mem_k.data_ptr() != proc.certain_mem_k.data_ptr()


def test_reference_style_prop_net_is_accepted():
    """A network that only offers the reference's interface (no read_memory/decode helpers) still works."""
    import evavos_b200 as ev
    prop, fuse = _nets()

    class Foreign(torch.nn.Module):   # stands in for the reference's own PropagationNetwork class
        def __init__(self, p):
            super().__init__()
            self.value_encoder, self.key_encoder, self.key_proj, self.key_comp = p.value_encoder, p.key_encoder, p.key_proj, p.key_comp
            self.decoder, self.attn_memory = p.decoder, p.attn_memory
            self.encode_key, self.encode_value, self.get_attention = p.encode_key, p.encode_value, p.get_attention

    g = load("e2e_k1.npz")
    images = torch.from_numpy(g["images"])
    a = ev.InferenceCore(prop, fuse, images, 1, mem_freq=2, device="cuda:0")
    b = ev.InferenceCore(Foreign(prop), fuse, images, 1, mem_freq=2, device="cuda:0")
    mask = torch.from_numpy(g["mask_0"])
    ma, mb = a.interact(mask, 0), b.interact(mask, 0)
    # `a` runs the key encoder / decoder once per segment on a batch of frames, `b` frame by frame: the same layers,
    # but cuDNN may pick another algorithm for another batch size, so agreement is to rounding, not bitwise
    assert (a.prob - b.prob).abs().max().item() < 1e-3
    assert (ma != mb).mean() < 1e-3


def test_cpu_device_rejected():
    import evavos_b200 as ev
    prop, fuse = _nets()
    with pytest.raises(RuntimeError, match="CUDA"):
        ev.InferenceCore(prop, fuse, torch.zeros(1, 2, 3, 128, 160), 1, device="cpu")


def test_bf16_channels_last_convs_agree_with_fp32():
    """``amp=True`` runs the conv encoders / decoder under bf16 autocast in channels_last; the keys,
    the values handed to the bank and the memory read stay fp32.  With seeded RANDOM weights (no checkpoint exists
    offline) the probabilities move by ~1e-2 and only pixels the fp32 run itself leaves near 0.5 may flip."""
    import evavos_b200 as ev
    prop, fuse = _nets()
    g = load("e2e_k1.npz")
    images = torch.from_numpy(g["images"])
    mask = torch.from_numpy(g["mask_0"])
    a = ev.InferenceCore(prop, fuse, images, 1, mem_freq=2, device="cuda:0")
    ma = a.interact(mask, 0)
    b = ev.InferenceCore(copy.deepcopy(prop), fuse, images, 1, mem_freq=2, device="cuda:0", amp=True)
    mb = b.interact(mask, 0)
    assert b.prob.dtype == torch.float32
    d = (a.prob - b.prob).abs()
    agree = float((ma == mb).mean())
    print(f"amp vs fp32: max |dp| {d.max().item():.4f}, mean |dp| {d.mean().item():.5f}, mask agreement {agree:.5f}")
    assert d.mean().item() < 2e-2
    decided = ((a.prob[1] - 0.5).abs() > 0.1).cpu().numpy()[:, 0]          # fp32 run at least 0.1 away from the boundary
    lw, uw, lh, uh = (int(x) for x in g["pad"])
    decided = decided[:, lh:decided.shape[1] - uh or None, lw:decided.shape[2] - uw or None]
    assert ((ma != mb) & decided).mean() < 1e-3
    assert agree > 0.97


@pytest.mark.parametrize("tag,opts", [
    ("e2e_k1", dict(fold_bn=False)),
    ("e2e_k2", dict(channels_last=True)),
    ("e2e_k2", dict(cuda_graphs=True)),
    ("e2e_k1", dict(cuda_graphs=True, channels_last=True)),
])
def test_engine_options_do_not_change_the_result(tag, opts):
    """BatchNorm folding + fused cuDNN conv-bias-ReLU (default), NHWC and CUDA-graph replay of the conv passes are
    re-orderings of the same arithmetic: probabilities agree to fp32 rounding with the plain engine, and with the
    reference's golden output."""
    import evavos_b200 as ev
    g = load(f"{tag}.npz")
    prop, fuse = _nets()
    images = torch.from_numpy(g["images"])
    k = int(g["num_objects"])
    plain = ev.InferenceCore(prop, fuse, images, k, mem_freq=int(g["mem_freq"]), device="cuda:0", fold_bn=False)
    tuned = ev.InferenceCore(prop, fuse, images, k, mem_freq=int(g["mem_freq"]), device="cuda:0", **opts)
    for i in range(int(g["n_interactions"])):
        mask, frame, scribble = torch.from_numpy(g[f"mask_{i}"]), int(g[f"frame_{i}"]), bool(g[f"scribble_{i}"])
        ma, mb = plain.interact(mask, frame, scribble=scribble), tuned.interact(mask, frame, scribble=scribble)
        assert (plain.prob - tuned.prob).abs().max().item() < 1e-3
        assert (ma != mb).mean() < 1e-3
        assert np.abs(tuned.prob.cpu().numpy() - g[f"prob_{i}"]).max() < 2e-3
    if opts.get("cuda_graphs"):
        # a second video through the same network replays the captured graphs; a deep copy of the processor
        # (interactions/policies.py:103) gets its own
        again = ev.InferenceCore(prop, fuse, images, k, mem_freq=int(g["mem_freq"]), device="cuda:0", **opts)
        mask, frame = torch.from_numpy(g["mask_0"]), int(g["frame_0"])
        first = ev.InferenceCore(prop, fuse, images, k, mem_freq=int(g["mem_freq"]), device="cuda:0", fold_bn=False)
        m1 = first.interact(mask, frame, scribble=bool(g["scribble_0"]))
        m2 = again.interact(mask, frame, scribble=bool(g["scribble_0"]))
        assert (first.prob - again.prob).abs().max().item() < 1e-3 and (m1 != m2).mean() < 1e-3
        clone = copy.deepcopy(again)
        assert torch.equal(clone.prob, again.prob)
