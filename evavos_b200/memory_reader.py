"""EvalMemoryReader with the reference's signatures on top of the sm_100a kernels.

Mirrors mivos/model/propagation/prop_net.py:74-115:

    reader = EvalMemoryReader(top_k=50, km=None)
    affinity = reader.get_affinity(mk, qk)      # mk (1,CK,T,H,W), qk (1,CK,H,W)
    mem = reader.readout(affinity, mv)          # mv (1,CV,T,H,W) -> (1,CV,H,W)

The reference's ``get_affinity`` returns the dense (1, THW, HW) matrix; here it returns a
compact :class:`TopKAffinity` (indices + weights, ``to_dense()`` on demand) that ``readout``
accepts.  ``read`` is the fused fast path used by ``segment_with_query``.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from .memory_bank import MemoryBank


class TopKAffinity:
    """Sparse form of the reference's top-k-softmax affinity (prop_net.py:53-60).

    idx (HW, k) int32 memory positions best-first, weight (HW, k) fp32 (rows sum to 1),
    score (HW, k) fp32 affinities (-a+b-c)/sqrt(CK).
    """

    def __init__(self, idx, weight, score, n_pos, height, width):
        self.idx, self.weight, self.score = idx, weight, score
        self.n_pos, self.height, self.width = int(n_pos), int(height), int(width)

    @property
    def shape(self):
        return (1, self.n_pos, self.idx.shape[0])

    def to_dense(self) -> torch.Tensor:
        """The (1, THW, HW) tensor the reference returns (x.zero_().scatter_(1, indices, x_exp))."""
        lib = _lib.load()
        nq, k = self.idx.shape
        dense = torch.empty((1, self.n_pos, nq), dtype=torch.float32, device=self.idx.device)
        with torch.cuda.device(self.idx.device):
            _lib.check(lib.evavos_affinity_dense(self.idx.data_ptr(), self.weight.data_ptr(), nq, k, self.n_pos,
                                                 dense.data_ptr(), _lib.current_stream_ptr(self.idx.device)))
        return dense


class _Workspace:
    """Grow-only device scratch per device, handed to the C ABI (the library never allocates)."""

    def __init__(self):
        self._buf = {}

    def get(self, device, nbytes: int) -> torch.Tensor:
        key = (device.type, device.index)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty((int(nbytes * 1.25) + 4096,), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


_workspace = _Workspace()


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the evavos_b200 memory read has no CPU path "
                           "(use oracle/ for CPU checking)")


def memory_read(bank: MemoryBank, qk: torch.Tensor, top_k: int = 50, n_frames: int | None = None,
                want_readout: bool = True, want_topk: bool = False, path: int = _lib.PATH_AUTO,
                sample_stride: int | None = None, out: torch.Tensor | None = None, peers=None,
                peer_gather_offset: int = 0):
    """Fused read of ``qk`` (1,CK,H,W) or (1,CK,F,H,W) against the first ``n_frames`` of ``bank``.

    ``sample_stride`` (tensor path): the filter's threshold pass contracts every sample_stride-th key tile
    (None/0: the library's choice, or $EVAVOS_SAMPLE_STRIDE when set - a tuning knob, results do not depend on it).
    ``out``: optional pre-allocated fp32 destination - (K, >=CV, *spatial), e.g. the first CV channels of the
    decoder's (K, 2*CV, H, W) input, or, for a (1,CK,F,H,W) query batch, frame-major (F, K, >=CV, H, W) - so that no
    torch.cat is needed (prop_net.py:189-190).
    ``peers`` (an ``_lib.Peers``) / ``peer_gather_offset``: sharded read - the finalizer also stores every query's
    list into all ranks' exchange buffers (see include/evavos.h, EvavosMemReadArgs.peers).
    Returns (readout (K,CV,[F,]H,W) or None, TopKAffinity or None).
    """
    lib = _lib.load()
    _require_cuda(qk, "query key")
    if qk.shape[0] != 1 or qk.shape[1] != bank.CK:
        raise ValueError(f"query key {tuple(qk.shape)} does not match bank CK={bank.CK} (batch must be 1)")
    spatial = tuple(qk.shape[2:])
    q2 = qk.to(torch.float32).reshape(bank.CK, -1)
    if q2.stride(1) != 1:
        q2 = q2.contiguous()
    nq = q2.shape[1]
    n_frames = bank.n_frames if n_frames is None else int(n_frames)
    n_pos = n_frames * bank.HW
    dev = bank.device

    a = _lib.MemReadArgs()
    a.bank = bank.shadow()
    a.query = q2.data_ptr()
    a.query_ch_stride = q2.stride(0)
    a.n_pos, a.n_query, a.top_k, a.path = n_pos, nq, int(top_k), int(path)
    a.sample_stride = int(sample_stride if sample_stride else os.environ.get("EVAVOS_SAMPLE_STRIDE", 0))
    if peers is not None:
        a.peers = ctypes.pointer(peers)
        a.peer_gather_offset = int(peer_gather_offset)
    idx = weight = score = None
    user_out = out
    frame_major = False
    if want_readout:
        if out is None:
            out = torch.empty((bank.K, bank.CV, nq), dtype=torch.float32, device=dev)
        elif len(spatial) == 3 and out.dim() == 5 and tuple(out.shape) == (spatial[0], bank.K, out.shape[2]) + spatial[1:]:
            # frame-major (F, K, C>=CV, H, W): one destination block per query frame
            hw_q = spatial[1] * spatial[2]
            if out.dtype == torch.float32 and out.device == dev and out.shape[2] >= bank.CV and out.stride(2) == 1 \
                    and hw_q > 1 and out.stride(4) == out.shape[2] and out.stride(3) == spatial[2] * out.shape[2] \
                    and (bank.K == 1 or out.stride(1) == hw_q * out.shape[2]):
                # ... whose (C, H, W) blocks are channels-last (the decoder input of an NHWC engine): the rows go out
                # as they are accumulated, channel-contiguous (readout_ch_stride = 1, include/evavos.h)
                a.readout_obj_stride, a.readout_ch_stride = hw_q * out.shape[2], 1
            else:
                if out.dtype != torch.float32 or out.device != dev or out.shape[2] < bank.CV or not out[0, 0, 0].is_contiguous():
                    raise ValueError(f"out {tuple(out.shape)} cannot receive a frame-major readout")
                a.readout_obj_stride, a.readout_ch_stride = out.stride(1), out.stride(2)
            a.queries_per_frame, a.readout_frame_stride = hw_q, out.stride(0)
            frame_major = True
        elif len(spatial) == 2 and out.dim() == 4 and out.dtype == torch.float32 and out.device == dev \
                and out.shape[0] == bank.K and out.shape[1] >= bank.CV and tuple(out.shape[2:]) == spatial \
                and out.stride(1) == 1 and nq > 1 and out.stride(3) == out.shape[1] \
                and out.stride(2) == spatial[1] * out.shape[1] and (bank.K == 1 or out.stride(0) == nq * out.shape[1]):
            # (K, C>=CV, H, W) in channels_last memory format
            a.readout_obj_stride, a.readout_ch_stride = nq * out.shape[1], 1
        else:
            # (K, C>=CV, *spatial) fp32 with contiguous positions: the kernel takes object / channel strides
            if out.dtype != torch.float32 or out.device != dev or out.shape[0] != bank.K or out.shape[1] < bank.CV \
                    or tuple(out.shape[2:]) != spatial or not out[0, 0].is_contiguous():
                raise ValueError(f"out {tuple(out.shape)} cannot receive a ({bank.K},{bank.CV},{spatial}) readout")
            a.readout_obj_stride, a.readout_ch_stride = out.stride(0), out.stride(1)
        a.readout = out.data_ptr()
    if want_topk:
        idx = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
        weight = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        score = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        a.topk_idx, a.topk_weight, a.topk_score = idx.data_ptr(), weight.data_ptr(), score.data_ptr()
    with torch.cuda.device(dev):
        a.n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        # sizing needs non-null placeholders only; real pointers are already set
        need = lib.evavos_memread_workspace_bytes(ctypes.byref(a))
        if need == 0:
            # validation failed: surface the library's message (e.g. THW < top_k, prop_net.py:53)
            a.workspace, a.workspace_bytes = 1, 0
            _lib.check(lib.evavos_memread(ctypes.byref(a), None))
        ws = _workspace.get(dev, need)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.check(lib.evavos_memread(ctypes.byref(a), _lib.current_stream_ptr(dev)))
    _last_read[0] = (a, dev, q2, ws)
    aff = TopKAffinity(idx, weight, score, n_pos, bank.H, bank.W) if want_topk else None
    if want_readout:
        if user_out is None:
            out = out.view(bank.K, bank.CV, *spatial)
        else:
            out = user_out[:, :, :bank.CV] if frame_major else user_out[:, :bank.CV]
    else:
        out = None
    return out, aff


_last_read = [None]


def last_overflow_count() -> int:
    """Diagnostics: the number of queries of the most recent ``memory_read`` whose candidate list overflowed and that
    were therefore selected by the exact tiled pass (evavos_memread_overflow_count).  Synchronises the stream."""
    if _last_read[0] is None:
        raise RuntimeError("no memory_read has run yet")
    a, dev, _, _ = _last_read[0]
    n = ctypes.c_uint32(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().evavos_memread_overflow_count(ctypes.byref(a), ctypes.byref(n), _lib.current_stream_ptr(dev)))
    return int(n.value)


class EvalMemoryReader(nn.Module):
    """Drop-in for prop_net.py:74-115 (``km`` Gaussian re-weighting is dead code upstream: km=None, :149)."""

    def __init__(self, top_k, km=None, *, append_only=True):
        super().__init__()
        if km is not None:
            raise NotImplementedError("km (kernelised memory) is never enabled by the reference (prop_net.py:149)")
        if top_k is None:
            raise NotImplementedError("top_k=None (full softmax) is not on the propagation path; see AttentionMemory")
        self.top_k = int(top_k)
        self.km = km
        self.append_only = bool(append_only)
        self._shadow_cache: "OrderedDict[tuple, tuple]" = OrderedDict()

    # ---- shadow of foreign reference-layout tensors ------------------------------------------------
    # A caller that keeps its bank as plain tensors (the reference's do_pass, inference_core.py:150-177) hands
    # growing T-slices ``keys[:, :, :m_front]`` of ONE allocation to every read.  The shadow of such a bank is kept
    # per allocation: the cache entry holds strong references to the source tensors (so their address cannot be
    # recycled by the caching allocator while the entry lives - a recycled address with an equal version counter
    # would otherwise alias another pass's bank) and remembers how many frames it has shadowed at which version.
    # With ``append_only`` (the reference's contract: frames below m_front are never rewritten, :174-177) a longer
    # slice of the same allocation only shadows the NEW frames; anything else rebuilds the shadow.
    @staticmethod
    def _alloc_sig(t):
        return None if t is None else (t.data_ptr(), t.stride(0), t.stride(1), t.stride(2), t.stride(3), t.stride(4),
                                       t.shape[0], t.shape[1], t.shape[3], t.shape[4], t.dtype)

    def _bank_for(self, mk: torch.Tensor, mv: torch.Tensor | None) -> MemoryBank:
        key = (self._alloc_sig(mk), self._alloc_sig(mv))
        t = mk.shape[2]
        versions = (mk._version, None if mv is None else mv._version)
        ent = self._shadow_cache.get(key)
        if ent is not None:
            bank, n_done, seen, _refs = ent
            if seen == versions and n_done >= t:
                self._shadow_cache.move_to_end(key)
                bank.n_frames = t                       # a shorter slice of an unchanged bank reads a prefix
                return bank
            if self.append_only and n_done <= t and bank.capacity_frames >= t and n_done > 0:
                if t > n_done:
                    bank.n_frames = n_done
                    bank.write_frames(n_done, mk[:, :, n_done:t], None if mv is None else mv[:, :, n_done:t])
                bank.n_frames = t
                self._shadow_cache[key] = (bank, t, versions, (mk, mv))
                self._shadow_cache.move_to_end(key)
                return bank
        # capacity: whole allocation when the slice is a prefix view of a longer tensor (the base is visible)
        base = mk._base if mk._base is not None and mk._base.dim() == 5 and mk._base.data_ptr() == mk.data_ptr() else None
        cap = max(t, base.shape[2] if base is not None else t)
        k, cv = (mv.shape[0], mv.shape[1]) if mv is not None else (0, 0)
        bank = MemoryBank(k, mk.shape[1], cv, mk.shape[3], mk.shape[4], cap, mk.device, keep_reference_layout=False)
        bank.write_frames(0, mk, mv)
        self._shadow_cache[key] = (bank, t, versions, (mk, mv))
        while len(self._shadow_cache) > 2:
            self._shadow_cache.popitem(last=False)
        return bank

    def get_affinity(self, mk, qk) -> TopKAffinity:
        """mk: (1,CK,T,H,W) tensor (strided T-slices allowed) or a MemoryBank; qk: (1,CK,H,W)."""
        if isinstance(mk, MemoryBank):
            bank = mk
        else:
            _require_cuda(mk, "memory key")
            bank = self._bank_for(mk, None)
        _, aff = memory_read(bank, qk, self.top_k, want_readout=False, want_topk=True)
        aff.height, aff.width = qk.shape[-2], qk.shape[-1]
        return aff

    def readout(self, affinity, mv) -> torch.Tensor:
        """affinity: TopKAffinity (or the dense (1,THW,HW) tensor); mv: (B,CV,T,H,W) or a MemoryBank."""
        lib = _lib.load()
        if isinstance(affinity, torch.Tensor):
            # dense matrix produced by to_dense(): recover its top_k non-zeros per query column
            w, i = torch.topk(affinity[0].transpose(0, 1), k=self.top_k, dim=1)
            affinity = TopKAffinity(i.to(torch.int32).contiguous(), w.contiguous(), None, affinity.shape[1], 0, 0)
        if isinstance(mv, MemoryBank):
            bank = mv
            h, w_ = bank.H, bank.W
        else:
            _require_cuda(mv, "memory value")
            b, cv, t, h, w_ = mv.shape
            bank = self._values_bank(mv)
        nq, k = affinity.idx.shape
        out = torch.empty((bank.K, bank.CV, nq), dtype=torch.float32, device=bank.device)
        sh = bank.shadow()
        with torch.cuda.device(bank.device):
            _lib.check(lib.evavos_readout(ctypes.byref(sh), affinity.idx.data_ptr(), affinity.weight.data_ptr(), nq, k,
                                          out.data_ptr(), 0, 0, _lib.current_stream_ptr(bank.device)))
        return out.view(bank.K, bank.CV, -1, w_) if nq % w_ == 0 else out

    def _values_bank(self, mv: torch.Tensor) -> MemoryBank:
        key = ("v", mv.data_ptr(), tuple(mv.shape), tuple(mv.stride()), mv._version)
        ent = self._shadow_cache.get(key)
        bank = ent[0] if ent is not None else None
        if bank is None:
            k, cv, t, h, w = mv.shape
            bank = MemoryBank(k, 8, cv, h, w, t, mv.device, keep_reference_layout=False)
            lib = _lib.load()
            src = mv.to(torch.float32)
            if not (src.stride(4) == 1 and src.stride(3) == w and (t == 1 or src.stride(2) == h * w)):
                src = src.contiguous()
            sh = bank.shadow()
            with torch.cuda.device(mv.device):
                _lib.check(lib.evavos_bank_write_values(ctypes.byref(sh), src.data_ptr(), src.stride(0), src.stride(1),
                                                        0, t * h * w, None, 0, 0, _lib.current_stream_ptr(mv.device)))
            bank.n_frames = t
            self._shadow_cache[key] = (bank, t, (mv._version,), (mv,))   # the reference pins the address
            while len(self._shadow_cache) > 2:
                self._shadow_cache.popitem(last=False)
        return bank

    def read(self, mk, qk, mv=None, n_frames=None) -> torch.Tensor:
        """Fused affinity + top-k softmax + readout for all objects -> (K,CV,H,W).

        ``mk`` may be a MemoryBank (then ``mv`` is ignored) or reference-layout tensors.
        """
        bank = mk if isinstance(mk, MemoryBank) else self._bank_for(mk, mv)
        out, _ = memory_read(bank, qk, self.top_k, n_frames=n_frames)
        return out

    def forward(self, mk, qk, mv):
        return self.read(mk, qk, mv)

    def __deepcopy__(self, memo):
        new = EvalMemoryReader(self.top_k, self.km, append_only=self.append_only)
        memo[id(self)] = new
        return new
