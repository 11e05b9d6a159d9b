"""InferenceCore with the reference's interface on top of the sm_100a memory read.

Mirrors mivos/inference_core.py:15-259 (constructor, ``interact``, ``do_pass``, ``fuse_one_frame``,
buffers and the externally read attributes ``prob``, ``masks``, ``np_masks``, ``pad``, ``t``, ``k``,
``h/w/nh/nw/kh/kw``, ``certain_mem_k/v``, ``interacted``), so ``interactions/eval.py:99`` and the policies
that call ``processor.interact(mask, frame)`` run unchanged.  What differs underneath:

* the per-pass memory bank is a :class:`MemoryBank` (reference-layout tensors + position-major shadow),
  seeded and appended by one kernel launch per frame instead of strided slice-assigns;
* certain memory lives in a capacity-doubling bank (no ``torch.cat`` re-allocation per ``interact``);
  ``certain_mem_k`` / ``certain_mem_v`` are views in the reference layout;
* query frames between two memory appends read the same bank, so they are read in ONE fused launch
  (up to ``mem_freq`` frames x HW queries) before being decoded frame by frame;
* affinity, top-k softmax and the readout of all objects are one C-ABI call; ``aggregate_wbg`` is one kernel;
* the key encoder and the decoder run once per segment on a batch of its frames (they only depend on the images
  and on the bank), not once per frame;
* the final per-frame argmax is one launch over all frames.

``prop_net`` may be this package's PropagationNetwork or the reference's own (same attribute names):
only ``encode_key``, ``encode_value``, ``decoder``, ``get_attention`` are used, never its torch memory reader.
"""
from __future__ import annotations

import numpy as np
import torch

from .aggregate import aggregate_wbg, argmax_unpad
from .memory_bank import MemoryBank
from .memory_reader import EvalMemoryReader
from .staging import download_numpy, upload
from .tensor_util import pad_divide_by


class InferenceCore:
    """
    images - unpadded, normalised, (1,T,3,H,W) (CPU or GPU)
    mem_profile - 0: everything on the GPU (the only profile any EVA-VOS caller uses); 1: frames stay on the
                  CPU and are staged per use; 2: small key-feature buffer.  Profile 3 of the reference spills the
                  results to the CPU and is broken upstream (inference_core.py:61 vs :118); it is rejected here.
    mem_freq - period at which propagated frames are added to the memory bank
    """

    def __init__(self, prop_net, fuse_net, images, num_objects, mem_profile=0, mem_freq=5, device="cuda", *,
                 amp=False, fold_bn=True, channels_last=None, cuda_graphs=False, fused_tails=True):
        """Keywords of this engine, not of the reference:
        ``amp``: run the conv encoders / decoder under bf16 autocast; keys, values and the memory read stay fp32.
        ``fold_bn``: run the two ResNet encoders through BatchNorm-folded copies (conv_opt.py; ``prop_net`` itself is
        not modified; only with this package's PropagationNetwork).
        ``channels_last``: NHWC conv stacks (default: same as ``amp``).
        ``cuda_graphs``: capture the key encoder, decoder and value encoder once per input shape and replay them
        (graphs.py; this package's PropagationNetwork only).  The graphs are cached on ``prop_net``.
        ``fused_tails``: with ``fold_bn`` and ``channels_last``, the decoder's bias / residual / ReLU / upsample tails
        run in the two kernels of csrc/decoder_ops.cu (conv_opt.FusedDecoder) instead of one PyTorch kernel each."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("evavos_b200.InferenceCore needs a CUDA device: the memory read has no CPU path")
        self.prop_net = prop_net.to(dev, non_blocking=True)
        if fuse_net is not None:
            self.fuse_net = fuse_net.to(dev, non_blocking=True)
        self.mem_profile = mem_profile
        self.mem_freq = mem_freq
        self.device = dev
        self.amp = bool(amp)
        self.channels_last = self.amp if channels_last is None else bool(channels_last)
        if self.channels_last:
            self.prop_net = self.prop_net.to(memory_format=torch.channels_last)
        ours = hasattr(prop_net, "decode_input")
        self.fold_bn = bool(fold_bn) and ours and not prop_net.training
        self.cuda_graphs = bool(cuda_graphs) and ours
        self.fused_tails = bool(fused_tails)

        if mem_profile == 0:
            self.data_dev, self.result_dev, self.k_buf_size, self.i_buf_size = dev, dev, 105, -1
        elif mem_profile == 1:
            self.data_dev, self.result_dev, self.k_buf_size, self.i_buf_size = torch.device("cpu"), dev, 105, 105
        elif mem_profile == 2:
            self.data_dev, self.result_dev, self.k_buf_size, self.i_buf_size = dev, dev, 3, -1
        else:
            raise NotImplementedError(f"mem_profile={mem_profile}: CPU-resident results are not supported "
                                      "(no caller uses them; the reference's profile 3 raises AttributeError)")

        t = images.shape[1]
        h, w = images.shape[-2:]
        self.k = num_objects

        if self.data_dev.type == "cuda" and not images.is_cuda:
            # upload first, pad on the GPU (the reference pads 150 MB of frames on the host, inference_core.py:63)
            images = upload(images, self.data_dev)
        self.images, self.pad = pad_divide_by(images, 16, images.shape[-2:])
        nh, nw = self.images.shape[-2:]
        self.images = self.images.to(self.data_dev, non_blocking=False)

        self.masks = torch.zeros((t, 1, nh, nw), dtype=torch.uint8, device=self.result_dev)
        self.np_masks = np.zeros((t, h, w), dtype=np.uint8)

        self.prob = torch.zeros((self.k + 1, t, 1, nh, nw), dtype=torch.float32, device=self.result_dev)
        self.prob[0] = 1e-7

        self.t, self.h, self.w = t, h, w
        self.nh, self.nw = nh, nw
        self.kh, self.kw = nh // 16, nw // 16

        self.key_buf = {}
        self.image_buf = {}
        self.interacted = set()

        # batched encoder / decoder passes over the frames of a segment need this package's PropagationNetwork
        # (decode_frames); a reference network passed in is driven frame by frame
        self._batch_frames = hasattr(prop_net, "decode_frames")
        self.key_batch = 16   # frames per key-encoder pass when a whole propagation pass is encoded ahead
        self._certain: MemoryBank | None = None
        mem = getattr(prop_net, "memory", None)
        top_k = getattr(mem, "top_k", 50)
        if top_k is None:
            raise NotImplementedError("prop_net.memory.top_k is None (full softmax over the bank): the reference never "
                                      "runs it (PropagationNetwork(top_k=50), prop_net.py:141) and the fused read "
                                      "implements the top-k form only")
        self._reader = mem if isinstance(mem, EvalMemoryReader) else EvalMemoryReader(top_k, None)

    # ------------------------------------------------------------------ reference-layout views of certain memory
    @property
    def certain_mem_k(self):
        return None if self._certain is None else self._certain.keys_view()

    @property
    def certain_mem_v(self):
        return None if self._certain is None else self._certain.values_view()

    # ------------------------------------------------------------------ buffers (inference_core.py:101-124)
    def get_image_buffered(self, idx):
        if self.data_dev == self.device:
            return self.images[:, idx]
        if idx not in self.image_buf and len(self.image_buf) > self.i_buf_size:
            self.image_buf = {}
        self.image_buf[idx] = self.images[:, idx].to(self.device)
        return self.image_buf[idx]

    def get_key_feat_buffered(self, idx):
        if idx not in self.key_buf:
            if len(self.key_buf) > self.k_buf_size:
                self.key_buf = {}
            self.key_buf[idx] = self._encode_key(self.get_image_buffered(idx))
        return self.key_buf[idx]

    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp)

    @property
    def _conv(self):
        """The three conv passes under this engine's options (conv_opt.ConvPasses), resolved on first use and after a
        deepcopy (policies deep-copy whole processors; captured graphs stay with the network they were made for)."""
        c = self.__dict__.get("_conv_cache")
        if c is None:
            from .conv_opt import conv_passes
            c = self.__dict__["_conv_cache"] = conv_passes(self.prop_net, self.amp, self.channels_last, self.fold_bn,
                                                           self.cuda_graphs, self.__dict__.get("fused_tails", True))
        return c

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_conv_cache", None)
        return state

    def _encode_key(self, frames):
        if hasattr(self.prop_net, "decode_input"):
            return self._conv.encode_key(frames)
        with self._autocast():      # a reference network passed in: its own encode_key, as it is
            outs = self.prop_net.encode_key(frames.contiguous(memory_format=torch.channels_last)
                                            if self.channels_last else frames)
        return (outs[0].float(),) + tuple(outs[1:])

    def _encode_value(self, frame, qf16, masks):
        if hasattr(self.prop_net, "decode_input"):
            return self._conv.encode_value(frame, qf16, masks)
        with self._autocast():
            return self.prop_net.encode_value(frame, qf16, masks).float()

    def _key_feats(self, frames):
        """Key features of several frames: the ones not cached yet go through the encoder as ONE batch (they only
        depend on the images), then through the same cache, with the same flush policy, as get_key_feat_buffered."""
        missing = [ti for ti in frames if ti not in self.key_buf]
        if len(missing) > 1 and self._batch_frames:
            batch = torch.cat([self.get_image_buffered(ti) for ti in missing], 0)
            outs = self._encode_key(batch)
            fresh = {}
            for j, ti in enumerate(missing):
                if len(self.key_buf) > self.k_buf_size:
                    self.key_buf = {}
                self.key_buf[ti] = fresh[ti] = tuple(o[j:j + 1] for o in outs)
            # a small cache (mem_profile 2) may have been flushed half-way: hand out what was just computed
            # instead of encoding those frames a second time
            return [fresh[ti] if ti in fresh else self.get_key_feat_buffered(ti) for ti in frames]
        return [self.get_key_feat_buffered(ti) for ti in frames]

    # ------------------------------------------------------------------ pieces of segment_with_query
    def _decode(self, readout, qf8, qf4, qv16):
        k = readout.shape[0]
        m4 = torch.cat([readout, qv16.expand(k, -1, -1, -1).to(readout.dtype)], 1)
        with self._autocast():
            return torch.sigmoid(self.prop_net.decoder(m4, qf8, qf4)).float()

    def _grow_certain(self, key_k, key_v):
        """certain_mem = cat(certain_mem, new frame) (inference_core.py:235-240) without re-allocating every time."""
        _, ck, _, hh, ww = key_k.shape
        kk, cv = key_v.shape[0], key_v.shape[1]
        if self._certain is None:
            self._certain = MemoryBank(kk, ck, cv, hh, ww, 4, self.device)
        elif self._certain.n_frames == self._certain.capacity_frames:
            bigger = MemoryBank(kk, ck, cv, hh, ww, 2 * self._certain.capacity_frames, self.device)
            bigger.write_frames(0, self._certain.keys_view(), self._certain.values_view())
            self._certain = bigger
        self._certain.append(key_k, key_v)

    # ------------------------------------------------------------------ propagation (inference_core.py:126-191)
    def do_pass(self, key_k, key_v, idx, forward=True):
        """Propagate from frame ``idx`` until the next interacted frame (exclusive) or the end of the video."""
        certain = self._certain
        num_certain_keys = certain.n_frames

        if forward:
            closest_ti = min([ti for ti in self.interacted if ti > idx] + [self.t])
            total_m = (closest_ti - idx - 1) // self.mem_freq + 1 + num_certain_keys
            this_range = list(range(idx + 1, closest_ti))
            end = closest_ti - 1
        else:
            closest_ti = max([ti for ti in self.interacted if ti < idx] + [-1])
            total_m = (idx - closest_ti - 1) // self.mem_freq + 1 + num_certain_keys
            this_range = list(range(idx - 1, closest_ti, -1))
            end = closest_ti + 1
        _, CK, _, H, W = key_k.shape
        K, CV = key_v.shape[0], key_v.shape[1]

        # pre-allocated bank, certain memory first (one import launch per tensor)
        bank = MemoryBank(K, CK, CV, H, W, total_m, self.device)
        bank.write_frames(0, certain.keys_view(), certain.values_view())
        last_ti = idx
        fuse = (closest_ti != self.t) and (closest_ti != -1)

        # key features only depend on the images: encode the frames of this pass ahead of time in larger batches,
        # as long as they all fit in the key-feature cache (k_buf_size, the reference's flush-all policy)
        if self._batch_frames:
            ahead = [ti for ti in this_range if ti not in self.key_buf]
            if len(self.key_buf) + len(ahead) <= self.k_buf_size:
                for c in range(0, len(ahead), self.key_batch):
                    self._key_feats(ahead[c:c + self.key_batch])

        pos = 0
        while pos < len(this_range):
            # frames up to (and including) the next memory frame all read the same bank: one fused read
            seg = []
            for ti in this_range[pos:]:
                seg.append(ti)
                if ti != end and abs(ti - last_ti) >= self.mem_freq:
                    break
            pos += len(seg)
            feats = self._key_feats(seg)
            qk = torch.stack([f[0] for f in feats], 2) if len(seg) > 1 else feats[0][0]
            decoded = readout = None
            if self._batch_frames:
                # The frames of a segment are independent given the bank: one read and one decoder pass for all of
                # them.  The read kernel writes straight into the readout half of the decoder input (F,K,2*CV,H,W);
                # the other half is the frames' query value feature - no torch.cat (prop_net.py:189-190).
                if self.channels_last:
                    # NHWC engine: the decoder's first convolutions take the input as it is (no 33 MB layout conversion
                    # per use), and the read kernel writes channel-contiguous rows straight from its accumulators
                    m4 = torch.empty((len(seg) * K, 2 * CV, H, W), dtype=torch.float32, device=self.device,
                                     memory_format=torch.channels_last).view(len(seg), K, 2 * CV, H, W)
                else:
                    m4 = torch.empty((len(seg), K, 2 * CV, H, W), dtype=torch.float32, device=self.device)
                self._read(bank, qk, out=m4 if len(seg) > 1 else m4[0])
                m4[:, :, CV:] = torch.cat([f[1] for f in feats], 0).unsqueeze(1)
                decoded = self._conv.decode(m4, torch.cat([f[3] for f in feats], 0), torch.cat([f[4] for f in feats], 0))
            else:
                readout, _ = self._read(bank, qk)
                if len(seg) == 1:
                    readout = readout.unsqueeze(2)
            for j, ti in enumerate(seg):
                k16, qv16, qf16, qf8, qf4 = feats[j]
                out_mask = decoded[j] if decoded is not None else self._decode(readout[:, :, j], qf8, qf4, qv16)
                out_mask = aggregate_wbg(out_mask, keep_bg=True)

                if ti != end and abs(ti - last_ti) >= self.mem_freq:
                    new_v = self._encode_value(self.get_image_buffered(ti), qf16, out_mask[1:])
                    bank.append(k16, new_v)
                    last_ti = ti

                if fuse:
                    self.prob[:, ti] = self.fuse_one_frame(closest_ti, idx, ti, self.prob[:, ti], out_mask,
                                                           key_k, k16).to(self.result_dev)
                else:
                    self.prob[:, ti] = out_mask.to(self.result_dev)
        return closest_ti

    def _read(self, bank, qk, out=None):
        from .memory_reader import memory_read
        return memory_read(bank, qk, self._reader.top_k, out=out)

    def fuse_one_frame(self, tc, tr, ti, prev_mask, curr_mask, mk16, qk16):
        assert (tc < ti < tr or tr < ti < tc)
        nc = abs(tc - ti) / abs(tc - tr)
        nr = abs(tr - ti) / abs(tc - tr)
        dist = torch.tensor([[nc, nr]], dtype=torch.float32, device=self.device)
        attn_map = self.prop_net.get_attention(mk16, self.pos_mask_diff, self.neg_mask_diff, qk16)
        # all objects in one batched pass of the fusion net (the reference loops k = 1..K with batch 1, :200-204)
        k = self.k
        prob = torch.sigmoid(self.fuse_net(self.get_image_buffered(ti).expand(k, -1, -1, -1),
                                           prev_mask[1:k + 1].to(self.device), curr_mask[1:k + 1].to(self.device),
                                           attn_map[1:k + 1], dist.expand(k, -1)))
        return aggregate_wbg(prob, keep_bg=True)

    # ------------------------------------------------------------------ interaction (inference_core.py:209-259)
    def interact(self, mask, idx, scribble=False):
        """Interact -> propagate both ways -> fuse.  Returns all masks as uint8 (T,h,w) numpy."""
        self.interacted.add(idx)

        mask = mask.to(self.device)
        mask, _ = pad_divide_by(mask, 16, mask.shape[-2:])
        self.mask_diff = mask - self.prob[:, idx].to(self.device)
        self.pos_mask_diff = self.mask_diff.clamp(0, 1)
        self.neg_mask_diff = (-self.mask_diff).clamp(0, 1)

        self.prob[:, idx] = mask
        key_k, _, qf16, _, _ = self.get_key_feat_buffered(idx)
        key_k = key_k.unsqueeze(2)
        key_v = self._encode_value(self.get_image_buffered(idx), qf16, mask[1:] if scribble else mask)

        self._grow_certain(key_k, key_v)

        self.do_pass(key_k, key_v, idx, True)
        self.do_pass(key_k, key_v, idx, False)

        # all frames, argmax + un-padding in one kernel (the reference loops T argmax calls and slices, :247-257)
        self.masks, unpadded = argmax_unpad(self.prob, self.pad, self.h, self.w)
        self.np_masks = download_numpy(unpadded)
        return self.np_masks
