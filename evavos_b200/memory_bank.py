"""The space-time memory bank: reference-layout tensors plus the engine-private shadow.

The reference allocates ``keys (1,CK,T,H,W)`` / ``values (K,CV,T,H,W)`` per pass and appends
with strided slice-assigns (mivos/inference_core.py:150-155, 174-177).  ``MemoryBank`` keeps
exactly those tensors (``.keys`` / ``.values``, sliceable ``[:, :, :m_front]`` like the
reference) and, written by the same kernel launch, a position-major shadow the read consumes
(see include/evavos.h).  All memory is owned by PyTorch; the C ABI only sees pointers.
"""
from __future__ import annotations

import copy
import ctypes

import torch

from . import _lib


def _positions_contiguous(t: torch.Tensor) -> bool:
    """True when dims (T,H,W) of a (B,C,T,H,W) view are mutually contiguous (one run per channel)."""
    _, _, tt, h, w = t.shape
    return t.stride(4) == 1 and t.stride(3) == w and (tt == 1 or t.stride(2) == h * w)


class MemoryBank:
    def __init__(self, num_objects: int, key_dim: int, value_dim: int, height: int, width: int,
                 capacity_frames: int, device, value_dtype: torch.dtype = torch.float32,
                 keep_reference_layout: bool = True):
        if value_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("value_dtype must be float32 or bfloat16")
        self.K, self.CK, self.CV, self.H, self.W = int(num_objects), int(key_dim), int(value_dim), int(height), int(width)
        self.HW = self.H * self.W
        self.capacity_frames = int(capacity_frames)
        self.device = torch.device(device)
        self.value_dtype = value_dtype
        self.n_frames = 0
        cap = self.capacity_frames * self.HW
        dev = self.device
        if keep_reference_layout:
            self.keys = torch.empty((1, self.CK, self.capacity_frames, self.H, self.W), dtype=torch.float32, device=dev)
            self.values = torch.empty((self.K, self.CV, self.capacity_frames, self.H, self.W), dtype=torch.float32,
                                      device=dev) if self.K > 0 else None
        else:
            self.keys = None
            self.values = None
        self.key_pm = torch.empty((cap, self.CK), dtype=torch.float32, device=dev)
        n_tile_bytes = int(_lib.load().evavos_key_tiles_bytes(cap))
        self.key_tiles = torch.empty((n_tile_bytes,), dtype=torch.uint8, device=dev) if self.CK == 64 else None
        self.key_maxnorm = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.val_pm = torch.empty((self.K, cap, self.CV), dtype=value_dtype, device=dev) if self.K > 0 else None

    # ------------------------------------------------------------------ C-ABI view
    def shadow(self) -> _lib.BankShadow:
        s = _lib.BankShadow()
        s.key_pm = self.key_pm.data_ptr()
        s.key_tiles = self.key_tiles.data_ptr() if self.key_tiles is not None else None
        s.key_maxnorm = self.key_maxnorm.data_ptr()
        s.val_pm = self.val_pm.data_ptr() if self.val_pm is not None else None
        s.capacity_pos = self.capacity_frames * self.HW
        s.K, s.CK, s.CV = self.K, self.CK, self.CV
        s.val_dtype = _lib.BF16 if self.value_dtype == torch.bfloat16 else _lib.F32
        return s

    @property
    def n_pos(self) -> int:
        return self.n_frames * self.HW

    # ------------------------------------------------------------------ writes
    def write_frames(self, slot: int, key_frames: torch.Tensor, value_frames: torch.Tensor | None) -> None:
        """Write ``t`` frames at ``slot``: key_frames (1,CK,t,H,W), value_frames (K,CV,t,H,W).

        The sources may be T-slices of another bank (strided views, inference_core.py:154-155).
        """
        lib = _lib.load()
        if key_frames.dim() == 4:
            key_frames = key_frames.unsqueeze(2)
        t = key_frames.shape[2]
        if key_frames.shape != (1, self.CK, t, self.H, self.W):
            raise ValueError(f"key frames {tuple(key_frames.shape)} do not match bank (1,{self.CK},t,{self.H},{self.W})")
        if slot < 0 or slot + t > self.capacity_frames:
            raise IndexError(f"frames [{slot}, {slot + t}) exceed bank capacity {self.capacity_frames}")
        key_frames = key_frames.to(device=self.device, dtype=torch.float32)
        if not _positions_contiguous(key_frames):
            key_frames = key_frames.contiguous()
        sh = self.shadow()
        stream = _lib.current_stream_ptr(self.device)
        pos0, n_pos = slot * self.HW, t * self.HW
        with torch.cuda.device(self.device):
            _lib.check(lib.evavos_bank_write_keys(
                ctypes.byref(sh), key_frames.data_ptr(), key_frames.stride(1), pos0, n_pos,
                self.keys.data_ptr() if self.keys is not None else None,
                self.keys.stride(1) if self.keys is not None else 0, stream))
            if self.K > 0:
                if value_frames is None:
                    raise ValueError("value_frames required for a bank with objects")
                if value_frames.dim() == 4:
                    value_frames = value_frames.unsqueeze(2)
                if value_frames.shape != (self.K, self.CV, t, self.H, self.W):
                    raise ValueError(f"value frames {tuple(value_frames.shape)} do not match bank "
                                     f"({self.K},{self.CV},{t},{self.H},{self.W})")
                value_frames = value_frames.to(device=self.device, dtype=torch.float32)
                if not _positions_contiguous(value_frames):
                    value_frames = value_frames.contiguous()
                _lib.check(lib.evavos_bank_write_values(
                    ctypes.byref(sh), value_frames.data_ptr(), value_frames.stride(0), value_frames.stride(1), pos0,
                    n_pos, self.values.data_ptr() if self.values is not None else None,
                    self.values.stride(0) if self.values is not None else 0,
                    self.values.stride(1) if self.values is not None else 0, stream))
        self.n_frames = max(self.n_frames, slot + t)

    def append(self, key_frame: torch.Tensor, value_frame: torch.Tensor | None) -> int:
        """keys[:,:,m_front] = k16 ; values[:,:,m_front] = v (inference_core.py:174-177). Returns the slot."""
        slot = self.n_frames
        self.write_frames(slot, key_frame, value_frame)
        return slot

    @classmethod
    def from_tensors(cls, mk: torch.Tensor, mv: torch.Tensor | None, value_dtype=torch.float32,
                     keep_reference_layout: bool = False) -> "MemoryBank":
        """Shadow of existing reference-layout tensors mk (1,CK,T,H,W), mv (K,CV,T,H,W)."""
        if mk.dim() != 5 or mk.shape[0] != 1:
            raise ValueError("memory keys must be (1,CK,T,H,W); the reference read is batch-1 (prop_net.py:54)")
        _, ck, t, h, w = mk.shape
        k, cv = (mv.shape[0], mv.shape[1]) if mv is not None else (0, 0)
        bank = cls(k, ck, cv, h, w, t, mk.device, value_dtype, keep_reference_layout)
        bank.write_frames(0, mk, mv)
        return bank

    # ------------------------------------------------------------------ reference-style views
    def keys_view(self) -> torch.Tensor:
        return self.keys[:, :, :self.n_frames]

    def values_view(self) -> torch.Tensor:
        return self.values[:, :, :self.n_frames]

    def __deepcopy__(self, memo):
        # interactions/policies.py:103 deep-copies the whole processor; tensors are cloned, nothing else is held.
        new = object.__new__(MemoryBank)
        memo[id(self)] = new
        for name, val in self.__dict__.items():
            setattr(new, name, val.clone() if isinstance(val, torch.Tensor) else copy.deepcopy(val, memo))
        return new
