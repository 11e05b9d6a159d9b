"""Inference-time rewrites of the convolutional encoders around the memory read.

The encoders are ordinary PyTorch/cuDNN modules (they are not the path this repository rewrites), but every
BatchNorm2d of the two ResNet trunks is a separate memory-bound kernel per layer and frame - 234 launches and ~8 ms of a
32-frame 480p video (scripts/profile_cfg3.py).  In eval mode a BatchNorm that directly follows a convolution is an
affine map of that convolution's output and folds into its weights and bias:

    w' = w * gamma / sqrt(var + eps)        b' = (b - mean) * gamma / sqrt(var + eps) + beta

The caller's module is never modified (its state dict must keep loading ``stcn.pth``): a folded deep copy is cached on
it and rebuilt when a parameter changes, moves, or changes memory format.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
from torch.nn.utils.fusion import fuse_conv_bn_eval


def fold_batchnorm(module: nn.Module) -> nn.Module:
    """Deep copy of ``module`` (which must be in eval mode) with every BatchNorm2d that is registered right after a
    Conv2d in the same parent - ``conv1, bn1`` of a ResNet block, ``(0, 1)`` of a ``downsample`` Sequential, which is
    also the order their forwards apply them in - folded into that convolution."""
    if module.training:
        raise RuntimeError("fold_batchnorm needs eval mode (running statistics)")
    folded = copy.deepcopy(module)
    for parent in folded.modules():
        names = list(parent._modules.keys())
        for a, b in zip(names, names[1:]):
            conv, bn = parent._modules[a], parent._modules[b]
            if isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and bn.track_running_stats \
                    and conv.out_channels == bn.num_features:
                parent._modules[a] = fuse_conv_bn_eval(conv, bn)
                parent._modules[b] = nn.Identity()
    return folded.eval().requires_grad_(False)


# ---- conv + bias + ReLU (+ residual) in one cuDNN call ----------------------------------------------------------------
# After folding, every trunk convolution is "conv + bias -> ReLU" or "conv + bias + identity -> ReLU".  PyTorch's
# Conv2d runs those as a convolution, a broadcast bias add, (an add,) and a clamp - three or four kernels, the
# elementwise ones as expensive as the BatchNorm they replaced.  cuDNN's fused convolution-bias-activation does them
# in the convolution's epilogue (aten::cudnn_convolution_relu / cudnn_convolution_add_relu).
def _conv_relu(conv, x):
    return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)


def _conv_add_relu(conv, x, z):
    return torch.cudnn_convolution_add_relu(x, conv.weight, z, 1.0, conv.bias, conv.stride, conv.padding,
                                            conv.dilation, conv.groups)


def _probe_fused(device, dtype, channels_last) -> bool:
    """Does this cuDNN build run the fused ops for (dtype, layout)?  One tiny call, compared with the unfused form."""
    if device.type != "cuda":
        return False
    try:
        with torch.no_grad(), torch.autocast("cuda", enabled=False):
            conv = nn.Conv2d(16, 16, 3, padding=1).to(device=device, dtype=dtype)
            x = torch.randn(2, 16, 12, 20, device=device, dtype=dtype)
            if channels_last:
                conv, x = conv.to(memory_format=torch.channels_last), x.contiguous(memory_format=torch.channels_last)
            ref = torch.relu(conv(x) + x).float()
            got = _conv_add_relu(conv, x, x).float()
            ref1, got1 = torch.relu(conv(x)).float(), _conv_relu(conv, x).float()
            tol = 5e-2 if dtype != torch.float32 else 1e-2      # (TF32 may be on for one form and not the other)
            return bool((got - ref).abs().max() <= tol * (1 + ref.abs().max()) and
                        (got1 - ref1).abs().max() <= tol * (1 + ref1.abs().max()))
    except Exception:
        return False


class _FoldedTrunk(nn.Module):
    """Shared plumbing: weights in ``dtype`` (fp32, or bf16 for the amp engine - no per-call weight casts, which
    autocast would launch for every convolution), layout, and whether the fused cuDNN ops are usable."""

    def __init__(self, enc: nn.Module, dtype, channels_last: bool):
        super().__init__()
        self.enc = enc
        self.dtype, self.channels_last = dtype, bool(channels_last)
        self.fused = _probe_fused(next(enc.parameters()).device, dtype, channels_last)

    def _in(self, x):
        x = x.to(self.dtype)
        return x.contiguous(memory_format=torch.channels_last) if self.channels_last else x

    def _cr(self, conv, x):
        return _conv_relu(conv, x) if self.fused else torch.relu(conv(x))

    def _car(self, conv, x, z):
        return _conv_add_relu(conv, x, z) if self.fused else torch.relu(conv(x) + z)

    def _bottleneck(self, blk, x):          # torchvision Bottleneck with folded norms
        idt = x if blk.downsample is None else blk.downsample(x)
        return self._car(blk.conv3, self._cr(blk.conv2, self._cr(blk.conv1, x)), idt)

    def _basic(self, blk, x):               # networks._BiasedBasicBlock with folded norms
        idt = x if blk.downsample is None else blk.downsample(x)
        return self._car(blk.conv2, self._cr(blk.conv1, x), idt)


class FoldedKeyEncoder(_FoldedTrunk):
    """networks.KeyEncoder.forward on a BatchNorm-folded copy: (f16, f8, f4)."""

    def forward(self, f):
        e = self.enc
        with torch.autocast("cuda", enabled=False):
            x = e.maxpool(self._cr(e.conv1, self._in(f)))
            for blk in e.res2:
                x = self._bottleneck(blk, x)
            f4 = x
            for blk in e.layer2:
                x = self._bottleneck(blk, x)
            f8 = x
            for blk in e.layer3:
                x = self._bottleneck(blk, x)
        return x, f8, f4


class FoldedValueEncoder(_FoldedTrunk):
    """networks.ValueEncoder.forward on a BatchNorm-folded copy; the fusion block (no norms) runs as it is, under the
    caller's autocast."""

    def __init__(self, enc, dtype, channels_last):
        super().__init__(enc, dtype, channels_last)
        self.trunk_modules = nn.ModuleList([enc.conv1, enc.layer1, enc.layer2, enc.layer3])

    def forward(self, image, key_f16, mask, other_masks):
        e = self.enc
        with torch.autocast("cuda", enabled=False):
            x = self._in(torch.cat([image, mask, other_masks], 1))
            x = e.maxpool(self._cr(e.conv1, x))
            for layer in (e.layer1, e.layer2, e.layer3):
                for blk in layer:
                    x = self._basic(blk, x)
        if not torch.is_autocast_enabled():     # the fusion block's weights are fp32
            x, key_f16 = x.float(), key_f16.float()
        return e.fuser(x, key_f16)


class FusedDecoder(nn.Module):
    """``PropagationNetwork.decode_input`` (prop_net.py:13-30, batched over frames and objects) on a private NHWC copy of
    the decoder in ``dtype``.  The decoder has no normalisation layers to fold; what sits between its convolutions is
    bias adds, residual adds, ReLUs and two bilinear x2 upsamplings, one PyTorch kernel each over up to 5 x 256 x 120 x
    216 elements.  Here every convolution that feeds a ReLU is a fused cuDNN conv-bias-ReLU, the others run without
    bias, and the tails are the two in-place kernels of csrc/decoder_ops.cu:

        ResBlock        x + conv2(relu(conv1(relu(x))))        -> bias_residual_(conv2', b2 (+ b_ds), x or downsample'(x))
        UpsampleBlock   skip_conv(skip) + up2x(x)              -> upsample2x_add_(skip_conv', b_skip, x)

    (primes: without bias).  The ReLU in front of ``pred`` rides on the last ResBlock's tail."""

    def __init__(self, dec: nn.Module, dtype):
        super().__init__()
        self.dec = copy.deepcopy(dec).eval().requires_grad_(False).to(memory_format=torch.channels_last).to(dtype)
        self.dtype = dtype
        self.fused = _probe_fused(next(dec.parameters()).device, dtype, True)
        for name, rb in (("compress", dec.compress), ("up_16_8", dec.up_16_8.out_conv), ("up_8_4", dec.up_8_4.out_conv)):
            b = rb.conv2.bias.detach().float()
            if rb.downsample is not None:
                b = b + rb.downsample.bias.detach().float()
            self.register_buffer("res_bias_" + name, b.contiguous().clone())
        for name in ("up_16_8", "up_8_4"):
            self.register_buffer("skip_bias_" + name, getattr(dec, name).skip_conv.bias.detach().float().contiguous().clone())

    def _in(self, x):
        return x.to(self.dtype).contiguous(memory_format=torch.channels_last)

    @staticmethod
    def _bare(conv, x):
        y = torch.nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        return y.contiguous(memory_format=torch.channels_last)       # (a no-op: cuDNN answers NHWC inputs in NHWC)

    def _res(self, rb, x, bias, relu_out=False):
        from .decoder_ops import bias_residual_
        a = torch.relu(x)
        c1 = _conv_relu(rb.conv1, a) if self.fused else torch.relu(rb.conv1(a))
        idt = x if rb.downsample is None else self._bare(rb.downsample, x)
        return bias_residual_(self._bare(rb.conv2, c1), bias, idt, relu=relu_out)

    def _up(self, blk, skip, x, k, skip_bias, res_bias, relu_out=False):
        from .decoder_ops import upsample2x_add_
        s = self._bare(blk.skip_conv, skip)
        if k != 1:      # the per-frame skip feature, once per object (the reference broadcasts it in the add)
            s = s.repeat_interleave(k, 0).contiguous(memory_format=torch.channels_last)
        return self._res(blk.out_conv, upsample2x_add_(s, skip_bias, x), res_bias, relu_out)

    def forward(self, m4, qf8, qf4):
        f, k, _, hh, ww = m4.shape
        d = self.dec
        with torch.autocast("cuda", enabled=False):
            x = self._res(d.compress, self._in(m4.reshape(f * k, -1, hh, ww)), self.res_bias_compress)
            x = self._up(d.up_16_8, self._in(qf8), x, k, self.skip_bias_up_16_8, self.res_bias_up_16_8)
            x = self._up(d.up_8_4, self._in(qf4), x, k, self.skip_bias_up_8_4, self.res_bias_up_8_4, relu_out=True)
            x = torch.nn.functional.interpolate(d.pred(x).float(), scale_factor=4, mode="bilinear", align_corners=False)
            return torch.sigmoid(x).view(f, k, 1, *x.shape[-2:])


def fused_decoder(prop_net: nn.Module, dtype=torch.float32) -> FusedDecoder:
    """The FusedDecoder of ``prop_net.decoder``; cached on ``prop_net`` like the folded encoders."""
    stamp = _stamp((prop_net.decoder,), (dtype,))
    slot = prop_net.__dict__.setdefault("_evavos_decoder", _CacheSlot())
    if slot.stamp != stamp:
        with torch.no_grad():
            slot.value = FusedDecoder(prop_net.decoder, dtype)
        slot.stamp = stamp
    return slot.value


def _stamp(mods, extra):
    tensors = [t for m in mods for t in list(m.parameters()) + list(m.buffers())]
    first = tensors[0]
    return (first.data_ptr(), first.device, sum(t._version for t in tensors), len(tensors)) + tuple(extra)


class _CacheSlot:
    """(stamp, value) kept in a module's __dict__; a deepcopy of the module gets an EMPTY slot (policies deep-copy whole
    processors: folded weights and captured graphs are rebuilt for the copy on first use, never copied)."""

    def __init__(self, stamp=None, value=None):
        self.stamp, self.value = stamp, value

    def __deepcopy__(self, memo):
        return _CacheSlot()


def folded_encoders(prop_net: nn.Module, channels_last: bool, dtype=torch.float32):
    """(key encoder, value encoder) of ``prop_net`` as FoldedKeyEncoder / FoldedValueEncoder; cached on ``prop_net`` and
    rebuilt when any of their parameters / buffers was written, moved or re-laid-out since.  ``dtype``: precision of
    the two ResNet trunks (the value encoder's fusion block keeps fp32 weights)."""
    mods = (prop_net.key_encoder, prop_net.value_encoder)
    stamp = _stamp(mods, (bool(channels_last), dtype))
    slot = prop_net.__dict__.setdefault("_evavos_folded", _CacheSlot())
    if slot.stamp != stamp:
        with torch.no_grad():
            ke, ve = (fold_batchnorm(m) for m in mods)
            if channels_last:
                ke, ve = ke.to(memory_format=torch.channels_last), ve.to(memory_format=torch.channels_last)
            ke = ke.to(dtype)
            for part in (ve.conv1, ve.layer1, ve.layer2, ve.layer3):
                part.to(dtype)
            slot.value = (FoldedKeyEncoder(ke, dtype, channels_last), FoldedValueEncoder(ve, dtype, channels_last))
        slot.stamp = stamp
    return slot.value


# ---- the three conv passes of the engine, eager or replayed ------------------------------------------------------------
class ConvPasses:
    """``encode_key`` / ``decode`` / ``encode_value`` of one PropagationNetwork under one set of engine options
    (autocast, layout, folded norms), as plain closures or - ``cuda_graphs`` - as graphs.GraphedPass replays.
    Built once per InferenceCore (``conv_passes``): the parameters of ``prop_net`` must not be re-allocated while an
    engine that holds it is alive (in-place updates are fine without graphs, and with them as long as ``fold_bn`` is
    off - a folded copy is a snapshot)."""

    def __init__(self, prop_net, amp, channels_last, fold_bn, cuda_graphs, fused_tails=True):
        self.amp, self.channels_last, self.fold_bn, self.cuda_graphs = amp, channels_last, fold_bn, cuda_graphs
        dtype = torch.bfloat16 if amp else torch.float32
        ke, ve = folded_encoders(prop_net, channels_last, dtype) if fold_bn else (None, None)
        self.fused = bool(ke is not None and ke.fused)

        def autocast():
            return torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp)

        def encode_key(frames):
            if channels_last:
                frames = frames.contiguous(memory_format=torch.channels_last)
            with autocast():
                outs = prop_net.encode_key(frames, ke) if fold_bn else prop_net.encode_key(frames)
            # the memory key (and everything the read touches) stays fp32; the skip features keep the conv dtype
            return (outs[0].float(),) + tuple(outs[1:])

        # NHWC engine with private weight copies: the decoder's elementwise tails run in csrc/decoder_ops.cu
        fd = fused_decoder(prop_net, dtype) if fold_bn and channels_last and fused_tails else None
        self.fused_decoder = fd is not None

        def decode(m4, qf8, qf4):
            if fd is not None:
                return fd(m4, qf8, qf4).float()
            with autocast():
                return prop_net.decode_input(m4, qf8, qf4).float()

        def encode_value(frame, qf16, masks):
            with autocast():
                v = prop_net.encode_value(frame, qf16, masks, ve) if fold_bn else prop_net.encode_value(frame, qf16, masks)
            return v.float()

        if cuda_graphs:
            from .graphs import GraphedPass
            pool = torch.cuda.graph_pool_handle()
            gk, gd, gv = GraphedPass(encode_key, pool), GraphedPass(decode, pool), GraphedPass(encode_value, pool)
            self.encode_key = lambda frames: gk(frames.contiguous(), clone=True)   # results live in the key cache
            self.decode = lambda m4, qf8, qf4: gd(m4, qf8, qf4)                   # consumed before the next replay
            self.encode_value = lambda frame, qf16, masks: gv(frame.contiguous(), qf16.contiguous(), masks.contiguous())
        else:
            self.encode_key, self.decode, self.encode_value = encode_key, decode, encode_value


def conv_passes(prop_net, amp=False, channels_last=False, fold_bn=True, cuda_graphs=False, fused_tails=True) -> ConvPasses:
    """The ConvPasses of ``prop_net`` for these options; cached on ``prop_net`` (captured graphs are worth keeping
    across the InferenceCores of a dataset) and rebuilt when its parameters were written, moved or re-laid-out.
    ``fused_tails``: with ``fold_bn`` and ``channels_last``, decode through FusedDecoder."""
    opts = (bool(amp), bool(channels_last), bool(fold_bn), bool(cuda_graphs), bool(fused_tails))
    stamp = _stamp((prop_net,), opts)
    slot = prop_net.__dict__.setdefault("_evavos_passes", _CacheSlot(value={}))
    if slot.value is None:
        slot.value = {}
    ent = slot.value.get(opts)
    if ent is None or ent[0] != stamp:
        ent = slot.value[opts] = (stamp, ConvPasses(prop_net, *opts))
    return ent[1]
