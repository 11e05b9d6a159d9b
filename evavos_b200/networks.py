"""Propagation / fusion networks around the memory read (PyTorch + cuDNN; not hand kernels).

The convolutional encoders and the decoder are outside the hot path this repository rewrites
(SURVEY.md section 2, rows 5-6): they stay ordinary PyTorch modules.  They are defined here only so
that the engine is self-contained and so that the reference's checkpoints load unchanged: module
attribute names, parameter names and shapes match ``stcn.pth`` / ``fusion.pth``
(tests/golden/propnet_state_dict.json, generated from the reference; 405 + 12 entries).

Architecture (STCN): ResNet-50 key encoder to stride 16 (1024 ch), 3x3 key projection to 64 ch,
3x3 "key_comp" to 512 ch; ResNet-18-style value encoder on image + object mask + other-objects mask
(5 input channels, biased convolutions) fused with the key feature by two residual blocks around a
CBAM gate -> 512 ch; decoder = residual compress + two skip-connected x2 upsampling stages + 1-channel
prediction, bilinearly upsampled x4.  (mivos/model/propagation/{modules,mod_resnet,cbam,prop_net}.py,
mivos/model/fusion_net.py)
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

from .memory_bank import MemoryBank
from .memory_reader import EvalMemoryReader


# ----------------------------------------------------------------------------- building blocks
class ResBlock(nn.Module):
    """Pre-activation residual pair of 3x3 convolutions; a 3x3 ``downsample`` when widths differ."""

    def __init__(self, indim: int, outdim: int | None = None):
        super().__init__()
        outdim = indim if outdim is None else outdim
        self.downsample = None if indim == outdim else nn.Conv2d(indim, outdim, 3, padding=1)
        self.conv1 = nn.Conv2d(indim, outdim, 3, padding=1)
        self.conv2 = nn.Conv2d(outdim, outdim, 3, padding=1)

    def forward(self, x):
        y = self.conv2(F.relu(self.conv1(F.relu(x))))
        return (x if self.downsample is None else self.downsample(x)) + y


class _ChannelGate(nn.Module):
    def __init__(self, channels: int, reduction: int = 16):
        super().__init__()
        # indices 1 and 3 carry the weights (0 = flatten, 2 = ReLU), as in the checkpoint
        self.mlp = nn.Sequential(nn.Flatten(), nn.Linear(channels, channels // reduction), nn.ReLU(),
                                 nn.Linear(channels // reduction, channels))

    def forward(self, x):
        att = self.mlp(F.adaptive_avg_pool2d(x, 1)) + self.mlp(F.adaptive_max_pool2d(x, 1))
        return x * torch.sigmoid(att)[:, :, None, None]


class _ConvOnly(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=(k - 1) // 2)

    def forward(self, x):
        return self.conv(x)


class _SpatialGate(nn.Module):
    def __init__(self):
        super().__init__()
        self.spatial = _ConvOnly(2, 1, 7)

    def forward(self, x):
        pooled = torch.cat([x.amax(1, keepdim=True), x.mean(1, keepdim=True)], 1)
        return x * torch.sigmoid(self.spatial(pooled))


class CBAM(nn.Module):
    """Channel gate then spatial gate (Woo et al. 2018)."""

    def __init__(self, channels: int):
        super().__init__()
        self.ChannelGate = _ChannelGate(channels)
        self.SpatialGate = _SpatialGate()

    def forward(self, x):
        return self.SpatialGate(self.ChannelGate(x))


class FeatureFusionBlock(nn.Module):
    def __init__(self, indim: int, outdim: int):
        super().__init__()
        self.block1 = ResBlock(indim, outdim)
        self.attention = CBAM(outdim)
        self.block2 = ResBlock(outdim, outdim)

    def forward(self, x, f16):
        x = self.block1(torch.cat([x, f16], 1))
        return self.block2(x + self.attention(x))


class _BiasedBasicBlock(nn.Module):
    """ResNet basic block whose convolutions keep their bias (the value encoder's checkpoint has them)."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class ValueEncoder(nn.Module):
    """image (3) + mask (1) + other-objects mask (1) -> 512 ch at stride 16, fused with the key feature."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(5, 64, 7, stride=2, padding=3)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = nn.Sequential(_BiasedBasicBlock(64, 64, 1), _BiasedBasicBlock(64, 64, 1))
        self.layer2 = nn.Sequential(_BiasedBasicBlock(64, 128, 2), _BiasedBasicBlock(128, 128, 1))
        self.layer3 = nn.Sequential(_BiasedBasicBlock(128, 256, 2), _BiasedBasicBlock(256, 256, 1))
        self.fuser = FeatureFusionBlock(1024 + 256, 512)

    def forward(self, image, key_f16, mask, other_masks):
        x = torch.cat([image, mask, other_masks], 1)
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer3(self.layer2(self.layer1(x)))
        return self.fuser(x, key_f16)


class KeyEncoder(nn.Module):
    """torchvision ResNet-50 trunk up to stride 16; returns (f16, f8, f4)."""

    def __init__(self):
        super().__init__()
        trunk = torchvision.models.resnet50(weights=None)
        self.conv1, self.bn1, self.relu, self.maxpool = trunk.conv1, trunk.bn1, trunk.relu, trunk.maxpool
        self.res2, self.layer2, self.layer3 = trunk.layer1, trunk.layer2, trunk.layer3

    def forward(self, f):
        x = self.maxpool(self.relu(self.bn1(self.conv1(f))))
        f4 = self.res2(x)
        f8 = self.layer2(f4)
        return self.layer3(f8), f8, f4


class UpsampleBlock(nn.Module):
    def __init__(self, skip_c: int, up_c: int, out_c: int, scale_factor: int = 2):
        super().__init__()
        self.skip_conv = nn.Conv2d(skip_c, up_c, 3, padding=1)
        self.out_conv = ResBlock(up_c, out_c)
        self.scale_factor = scale_factor

    def forward(self, skip_f, up_f):
        up = F.interpolate(up_f, scale_factor=self.scale_factor, mode="bilinear", align_corners=False)
        return self.out_conv(self.skip_conv(skip_f) + up)


class KeyProjection(nn.Module):
    def __init__(self, indim: int, keydim: int):
        super().__init__()
        self.key_proj = nn.Conv2d(indim, keydim, 3, padding=1)
        nn.init.orthogonal_(self.key_proj.weight.data)
        nn.init.zeros_(self.key_proj.bias.data)

    def forward(self, x):
        return self.key_proj(x)


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.compress = ResBlock(1024, 512)
        self.up_16_8 = UpsampleBlock(512, 512, 256)
        self.up_8_4 = UpsampleBlock(256, 256, 256)
        self.pred = nn.Conv2d(256, 1, 3, padding=1)

    def forward(self, f16, f8, f4):
        x = self.up_8_4(f4, self.up_16_8(f8, self.compress(f16)))
        x = self.pred(F.relu(x))
        return F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)


class AttentionMemory(nn.Module):
    """Full-softmax affinity of one memory frame (fusion path only, prop_net.py:117-138).

    Kept with the reference's signature for callers of ``get_W``; ``get_attention`` itself uses the fused
    attention read (evavos_b200.attention) and never builds this matrix on the GPU.
    """

    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, mk, qk):
        ck = mk.shape[1]
        m = mk.flatten(start_dim=2)
        q = qk.flatten(start_dim=2)
        aff = (2 * (m.transpose(1, 2) @ q) - m.pow(2).sum(1).unsqueeze(2) - q.pow(2).sum(1).unsqueeze(1)) / math.sqrt(ck)
        return F.softmax(aff, dim=1)


# ----------------------------------------------------------------------------- the two networks
class PropagationNetwork(nn.Module):
    """Same sub-module names, methods and state dict as the reference's PropagationNetwork (prop_net.py:140-211);
    ``memory`` is the sm_100a reader and ``segment_with_query`` uses its fused read."""

    def __init__(self, top_k: int = 50):
        super().__init__()
        self.value_encoder = ValueEncoder()
        self.key_encoder = KeyEncoder()
        self.key_proj = KeyProjection(1024, keydim=64)
        self.key_comp = nn.Conv2d(1024, 512, kernel_size=3, padding=1)
        self.memory = EvalMemoryReader(top_k, km=None)
        self.attn_memory = AttentionMemory(top_k)
        self.decoder = Decoder()
        self._others_idx = {}

    def encode_value(self, frame, kf16, masks, value_encoder=None):
        """frame (1,3,h,w), kf16 (1,1024,H,W), masks (K,1,h,w) -> (K,512,1,H,W) (prop_net.py:153-170).
        ``value_encoder``: a stand-in for ``self.value_encoder`` (the BatchNorm-folded copy of conv_opt.py)."""
        k, _, h, w = masks.shape
        frame = frame.view(1, 3, h, w).expand(k, -1, -1, -1)
        kf16 = kf16.expand(k, -1, -1, -1)
        if k != 1:
            # per object, the sum of every OTHER object's mask (summed in index order, like the reference); gathered
            # with a cached index tensor: boolean indexing would synchronise (and cannot be captured in a CUDA graph)
            key = (k, masks.device)
            idx = self._others_idx.get(key)
            if idx is None:
                idx = torch.tensor([[j for j in range(k) if j != i] for i in range(k)], device=masks.device)
                self._others_idx[key] = idx
            others = masks[:, 0][idx].sum(1).unsqueeze(1)
        else:
            others = torch.zeros_like(masks)
        return (value_encoder or self.value_encoder)(frame, kf16, masks, others).unsqueeze(2)

    def encode_key(self, frame, key_encoder=None):
        f16, f8, f4 = (key_encoder or self.key_encoder)(frame)
        return self.key_proj(f16), self.key_comp(f16), f16, f8, f4

    def read_memory(self, mk16, mv16, qk16):
        """(K,512,[F,]H,W) readout of query key(s) against a MemoryBank or reference-layout tensors."""
        if isinstance(mk16, MemoryBank):
            return self.memory.read(mk16, qk16)
        return self.memory.read(mk16, qk16, mv16)

    def decode_frames(self, readout, qf8, qf4, qv16):
        """Batched ``decode`` for F query frames that were read against the same bank.

        readout (K,512,F,H,W), qf8 / qf4 / qv16 with batch F -> (F,K,1,h,w) probabilities.  The same layers as
        ``Decoder.forward`` (prop_net.py:13-30); the per-frame skip features, which the single-frame path broadcasts
        over the K objects inside ``skip_conv(skip) + up``, are repeated explicitly so that one pass covers F*K maps.
        """
        k = readout.shape[0]
        m4 = torch.cat([readout.permute(2, 0, 1, 3, 4), qv16.unsqueeze(1).expand(-1, k, -1, -1, -1)], 2)
        return self.decode_input(m4, qf8, qf4)

    def decode_input(self, m4, qf8, qf4):
        """The decoder on an assembled input: m4 (F,K,1024,H,W) = [memory readout | query value feature] per frame and
        object (prop_net.py:189-190).  InferenceCore lets the read kernel write the readout half of m4 in place, so
        the (K,1024,H,W) ``torch.cat`` of the reference never happens."""
        f, k, _, hh, ww = m4.shape
        dec = self.decoder
        x = dec.compress(m4.reshape(f * k, -1, hh, ww))
        for block, skip in ((dec.up_16_8, qf8), (dec.up_8_4, qf4)):
            up = F.interpolate(x, scale_factor=block.scale_factor, mode="bilinear", align_corners=False)
            x = block.out_conv(block.skip_conv(skip).repeat_interleave(k, 0) + up)
        x = dec.pred(F.relu(x))
        x = F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)
        return torch.sigmoid(x).view(f, k, 1, *x.shape[-2:])

    def decode(self, readout, qf8, qf4, qv16):
        """readout (K,512,H,W) + shared query value feature -> (K,1,h,w) probabilities (prop_net.py:189-192)."""
        k = readout.shape[0]
        m4 = torch.cat([readout, qv16.expand(k, -1, -1, -1)], 1)
        return torch.sigmoid(self.decoder(m4, qf8, qf4))

    def segment_with_query(self, mk16, mv16, qf8, qf4, qk16, qv16):
        return self.decode(self.read_memory(mk16, mv16, qk16), qf8, qf4, qv16)

    def get_W(self, mk16, qk16):
        return self.attn_memory(mk16, qk16)

    def get_attention(self, mk16, pos_mask, neg_mask, qk16):
        """prop_net.py:198-211.  W = get_W(mk16, qk16) is never built: one fused attention read
        (evavos_attention_readout, CUDA tensors only - there is no CPU path) produces the stride-16 positive /
        negative maps directly.  Any number of objects: the mask rows go through the kernel 32 at a time."""
        b, _, h, w = pos_mask.shape
        nh, nw = h // 16, w // 16
        pos = F.interpolate(pos_mask, size=(nh, nw), mode="area").view(b, 1, nh * nw)
        neg = F.interpolate(neg_mask, size=(nh, nw), mode="area").view(b, 1, nh * nw)
        from .attention import attention_readout
        attn = attention_readout(mk16, qk16, torch.cat([pos, neg], 1).view(2 * b, nh * nw)).view(b, 2, nh, nw)
        return F.interpolate(attn, mode="bilinear", size=(h, w), align_corners=False)


class FusionNet(nn.Module):
    """9-channel (image, two masks, two attention maps, two time scalars) residual mixer (fusion_net.py:8-50)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(9, 32, 3, padding=1), nn.ReLU())
        self.conv2 = nn.Sequential(nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 32, 3, padding=1))
        self.conv3 = nn.Sequential(nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 32, 3, padding=1))
        self.relu = nn.ReLU()
        self.final_conv = nn.Conv2d(32, 1, 3, padding=1)

    def forward(self, im, seg1, seg2, attn, time):
        h, w = im.shape[-2:]
        t = time[:, :, None, None].expand(-1, -1, h, w)
        x = self.conv1(torch.cat([im, seg1, seg2, attn, t], 1))
        x = self.relu(x + self.conv2(x))
        x = self.relu(x + self.conv3(x))
        return self.final_conv(x)


def seeded_init(module: nn.Module, seed: int, gain: float = 0.6) -> None:
    """Deterministic, well-conditioned random weights for tests (no checkpoints exist offline).

    Walks ``state_dict()`` in order with a CPU generator, so the reference's networks and ours receive
    identical values through their identical key sets.
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            if name.endswith("running_var"):
                v = 0.8 + 0.4 * torch.rand(t.shape, generator=g)
            elif name.endswith("running_mean"):
                v = 0.05 * torch.randn(t.shape, generator=g)
            elif t.dim() >= 2:
                fan_in = t[0].numel()
                v = torch.randn(t.shape, generator=g) * math.sqrt(gain / fan_in)
            elif name.endswith("weight"):      # norm scales
                v = 0.9 + 0.2 * torch.rand(t.shape, generator=g)
            else:                               # biases
                v = 0.02 * torch.randn(t.shape, generator=g)
            t.copy_(v.to(t.dtype))
