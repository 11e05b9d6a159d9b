"""The reference's stock propagation loop as plain torch ops (TEST INFRASTRUCTURE ONLY).

``bench.py --workload cfg3`` times this on the B200 as ``gpu_baseline``: what a user of the reference gets on the same
GPU before switching - one query frame at a time, a dense (N, HW) affinity + ``topk`` + scatter, one ``bmm`` per
object, the (K,1024,H,W) ``cat``, BatchNorm modules as they are, ``argmax`` frame by frame.  The reference itself
(/root/reference, pure Python with un-vendored imports) does not travel to the GPU box; this restatement runs the same
ATen op sequence through ``oracle/torch_port.py`` (bit-identical to the reference's reader on CPU) and the plain
``forward``s of ``evavos_b200.networks`` (the reference's architecture and state dict).  It never touches the CUDA
extension.  Pinned: ``tests/test_oracle_golden.py::test_stock_engine_matches_reference`` replays the interactions of
``tests/golden/e2e_*.npz`` (recorded from the live reference ``InferenceCore``) on CPU.

Follows mivos/inference_core.py (memory profile 0 only): __init__ :36-99, key-feature cache :118-128, do_pass :130-200,
fuse_one_frame :202-218, interact :220-270; prop_net.py segment_with_query :177-192.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import torch_port as tp


def _pad16(x):
    """Zero-pad the last two axes to multiples of 16, split evenly (tensor_util.py:62-93). Returns (padded, lrtb)."""
    h, w = x.shape[-2:]
    ph, pw = (-h) % 16, (-w) % 16
    lrtb = (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)
    return F.pad(x, lrtb), lrtb


class StockEngine:
    KEY_CACHE = 105                                              # k_buf_size of memory profile 0

    def __init__(self, prop_net, fuse_net, images, num_objects, mem_freq=5, device="cuda"):
        self.prop, self.fuse, self.dev = prop_net.to(device), fuse_net.to(device) if fuse_net is not None else None, device
        self.mem_freq, self.k = mem_freq, num_objects
        self.t, (self.h, self.w) = images.shape[1], images.shape[-2:]
        padded, self.pad = _pad16(images)
        self.images = padded.to(device)
        self.nh, self.nw = self.images.shape[-2:]
        self.masks = torch.zeros((self.t, 1, self.nh, self.nw), dtype=torch.uint8, device=device)
        self.prob = torch.zeros((num_objects + 1, self.t, 1, self.nh, self.nw), dtype=torch.float32, device=device)
        self.prob[0] = 1e-7
        self.feats, self.interacted = {}, set()
        self.sure_k = self.sure_v = None                         # memory of the frames the user annotated

    def _features(self, ti):
        if ti not in self.feats:
            if len(self.feats) > self.KEY_CACHE:
                self.feats = {}
            self.feats[ti] = self.prop.encode_key(self.images[:, ti])
        return self.feats[ti]

    def _segment(self, mem_k, mem_v, qf8, qf4, qk16, qv16):
        """prop_net.py:177-192 with the reader spelled out: (K,1,nh,nw) object probabilities."""
        read = tp.memory_read(mem_k, qk16, mem_v, self.prop.memory.top_k)
        m4 = torch.cat([read, qv16.expand(mem_v.shape[0], -1, -1, -1)], 1)
        return torch.sigmoid(self.prop.decoder(m4, qf8, qf4))

    def _fuse(self, t_other, t_from, ti, before, now, key_from, qk16):
        """Blend this pass's result with the earlier one between two annotated frames (inference_core.py:202-218)."""
        span = abs(t_other - t_from)
        dist = torch.tensor([[abs(t_other - ti) / span, abs(t_from - ti) / span]], dtype=torch.float32, device=self.dev)
        attn = tp.get_attention(key_from, self.pos_diff, self.neg_diff, qk16)
        out = torch.zeros((self.k, 1, self.nh, self.nw), dtype=torch.float32, device=self.dev)
        for o in range(1, self.k + 1):
            out[o - 1] = torch.sigmoid(self.fuse(self.images[:, ti], before[o:o + 1], now[o:o + 1], attn[o:o + 1], dist))
        return tp.aggregate_wbg(out, keep_bg=True)

    def _sweep(self, key_k, idx, step):
        """One direction of do_pass (inference_core.py:130-200): propagate from idx until the next annotated frame."""
        if step > 0:
            stop = min([ti for ti in self.interacted if ti > idx] + [self.t])
        else:
            stop = max([ti for ti in self.interacted if ti < idx] + [-1])
        n_sure = self.sure_k.shape[2]
        n_slots = (abs(stop - idx) - 1) // self.mem_freq + 1 + n_sure
        keys = torch.empty((1, key_k.shape[1], n_slots) + tuple(key_k.shape[-2:]), dtype=torch.float32, device=self.dev)
        values = torch.empty(tuple(self.sure_v.shape[:2]) + (n_slots,) + tuple(key_k.shape[-2:]), dtype=torch.float32,
                             device=self.dev)
        keys[:, :, :n_sure], values[:, :, :n_sure] = self.sure_k, self.sure_v
        front, last_added, between = n_sure, idx, stop not in (self.t, -1)
        for ti in range(idx + step, stop, step):
            qk16, qv16, qf16, qf8, qf4 = self._features(ti)
            seg = tp.aggregate_wbg(self._segment(keys[:, :, :front], values[:, :, :front], qf8, qf4, qk16, qv16),
                                   keep_bg=True)
            if ti != stop - step and abs(ti - last_added) >= self.mem_freq:
                keys[:, :, front:front + 1] = qk16.unsqueeze(2)
                values[:, :, front:front + 1] = self.prop.encode_value(self.images[:, ti], qf16, seg[1:])
                front, last_added = front + 1, ti
            self.prob[:, ti] = self._fuse(stop, idx, ti, self.prob[:, ti], seg, key_k, qk16) if between else seg

    def interact(self, mask, idx, scribble=False):
        """inference_core.py:220-270: returns the (T,h,w) uint8 masks."""
        self.interacted.add(idx)
        mask, _ = _pad16(mask.to(self.dev))
        diff = mask - self.prob[:, idx]
        self.pos_diff, self.neg_diff = diff.clamp(0, 1), (-diff).clamp(0, 1)
        self.prob[:, idx] = mask
        qk16, _, qf16, _, _ = self._features(idx)
        key_k = qk16.unsqueeze(2)
        key_v = self.prop.encode_value(self.images[:, idx], qf16, mask[1:] if scribble else mask)
        first = self.sure_k is None
        self.sure_k = key_k if first else torch.cat([self.sure_k, key_k], 2)
        self.sure_v = key_v if first else torch.cat([self.sure_v, key_v], 2)
        self._sweep(key_k, idx, +1)
        self._sweep(key_k, idx, -1)
        for ti in range(self.t):
            self.masks[ti] = torch.argmax(self.prob[:, ti], dim=0)
        l, r, t, b = self.pad
        out = self.masks[:, 0, t:self.nh - b, l:self.nw - r]
        return out.cpu().numpy().astype(np.uint8)
