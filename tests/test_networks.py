"""CPU: networks keep the reference's state dict; engine classes keep the reference's signatures."""
import inspect
import json
import os

import pytest
import torch

from tests.helpers import GOLDEN


def test_state_dict_matches_reference_checkpoints():
    from evavos_b200.networks import FusionNet, PropagationNetwork
    spec = json.load(open(os.path.join(GOLDEN, "propnet_state_dict.json")))
    for net, key in ((PropagationNetwork(), "prop"), (FusionNet(), "fuse")):
        mine = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert mine == spec[key]
        assert list(mine) == list(spec[key])          # same order: seeded_init walks it
    p = PropagationNetwork(top_k=20)
    assert p.memory.top_k == 20
    for name in ("value_encoder", "key_encoder", "key_proj", "key_comp", "memory", "attn_memory", "decoder"):
        assert hasattr(p, name)


def test_signatures_match_reference_api():
    """constructor / method names and defaults the callers rely on."""
    from evavos_b200 import EvalMemoryReader, InferenceCore, PropagationNetwork, aggregate_wbg
    sig = inspect.signature(InferenceCore.__init__)
    assert list(sig.parameters)[1:8] == ["prop_net", "fuse_net", "images", "num_objects", "mem_profile", "mem_freq", "device"]
    assert all(sig.parameters[n].kind is inspect.Parameter.KEYWORD_ONLY and sig.parameters[n].default is not inspect.Parameter.empty
               for n in list(sig.parameters)[8:])     # engine-only options are optional keywords
    assert sig.parameters["mem_profile"].default == 0 and sig.parameters["mem_freq"].default == 5
    assert sig.parameters["device"].default == "cuda"
    assert list(inspect.signature(InferenceCore.interact).parameters)[1:] == ["mask", "idx", "scribble"]
    assert list(inspect.signature(InferenceCore.do_pass).parameters)[1:] == ["key_k", "key_v", "idx", "forward"]
    assert list(inspect.signature(InferenceCore.fuse_one_frame).parameters)[1:] == ["tc", "tr", "ti", "prev_mask", "curr_mask", "mk16", "qk16"]
    rd = inspect.signature(EvalMemoryReader.__init__).parameters
    assert list(rd)[1:3] == ["top_k", "km"]      # the reference's positional interface (prop_net.py:75)
    assert all(rd[n].kind is inspect.Parameter.KEYWORD_ONLY and rd[n].default is not inspect.Parameter.empty
               for n in list(rd)[3:])            # anything else is an optional keyword of this engine
    assert list(inspect.signature(EvalMemoryReader.get_affinity).parameters)[1:] == ["mk", "qk"]
    assert list(inspect.signature(EvalMemoryReader.readout).parameters)[1:] == ["affinity", "mv"]
    assert list(inspect.signature(PropagationNetwork.segment_with_query).parameters)[1:] == ["mk16", "mv16", "qf8", "qf4", "qk16", "qv16"]
    s = inspect.signature(aggregate_wbg)
    assert list(s.parameters) == ["prob", "keep_bg", "hard"] and s.parameters["keep_bg"].default is False


def test_networks_forward_shapes_cpu():
    """The conv stacks are plain torch: run them on the CPU at a tiny size (no memory read involved)."""
    from evavos_b200.networks import FusionNet, PropagationNetwork, seeded_init
    torch.set_grad_enabled(False)
    try:
        p = PropagationNetwork().eval()
        seeded_init(p, 3)
        frame = torch.rand(1, 3, 64, 96)
        k16, f16_thin, f16, f8, f4 = p.encode_key(frame)
        assert k16.shape == (1, 64, 4, 6) and f16_thin.shape == (1, 512, 4, 6) and f16.shape == (1, 1024, 4, 6)
        masks = torch.rand(2, 1, 64, 96)
        v = p.encode_value(frame, f16, masks)
        assert v.shape == (2, 512, 1, 4, 6)
        out = p.decode(torch.rand(2, 512, 4, 6), f8, f4, f16_thin)
        assert out.shape == (2, 1, 64, 96) and float(out.min()) >= 0 and float(out.max()) <= 1
        # the attention read exists as a CUDA kernel only (no CPU path in the product)
        with pytest.raises(RuntimeError, match="CUDA"):
            p.get_attention(k16.unsqueeze(2), torch.rand(3, 1, 64, 96), torch.rand(3, 1, 64, 96), k16)
        att = torch.rand(3, 2, 64, 96)
        f = FusionNet().eval()
        o = f(frame, masks[:1], masks[1:], att[:1], torch.tensor([[0.3, 0.7]]))
        assert o.shape == (1, 1, 64, 96)
    finally:
        torch.set_grad_enabled(True)


def test_import_path_shim():
    import sys
    from tests.helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "evavos_b200", "compat"))
    try:
        for mod in [m for m in sys.modules if m == "mivos" or m.startswith("mivos.")]:
            del sys.modules[mod]
        from mivos.inference_core import InferenceCore
        from mivos.model.aggregate import aggregate_wbg
        from mivos.model.fusion_net import FusionNet
        from mivos.model.propagation.prop_net import EvalMemoryReader, PropagationNetwork
        from mivos.tensor_util import pad_divide_by
        assert InferenceCore.__module__.startswith("evavos_b200")
        assert PropagationNetwork.__module__.startswith("evavos_b200") and FusionNet and aggregate_wbg and pad_divide_by
        assert EvalMemoryReader.__module__.startswith("evavos_b200")
    finally:
        sys.path.pop(0)
        for mod in [m for m in sys.modules if m == "mivos" or m.startswith("mivos.")]:
            del sys.modules[mod]
