"""Host <-> device copies of the video and of the result masks through cached pinned staging buffers.

A pageable ``tensor.to("cuda")`` of a 32-frame 480p video (159 MB) takes ~14 ms on the GPU box - one thread copying
into the driver's bounce buffer - and the pageable read-back of the masks another ~5 ms; both sit on the critical path
of every ``InferenceCore`` (inference_core.py:63 / :257 of the reference do the same).  Here the host memcpy (torch's
multi-threaded ``copy_`` into pinned memory) of chunk i+1 overlaps the DMA of chunk i, and the staging buffers are
allocated once per process.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

_CHUNK = 32 << 20
_lock = threading.Lock()
_bufs: dict = {}      # slot -> pinned uint8 tensor
_events: dict = {}    # slot -> event recorded after the last DMA that touched the slot's buffer


def _slot(slot: int, nbytes: int) -> torch.Tensor:
    ev = _events.get(slot)
    if ev is not None:
        ev.synchronize()
    buf = _bufs.get(slot)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(nbytes, _CHUNK),), dtype=torch.uint8).pin_memory()
        _bufs[slot] = buf
    return buf


def upload(src: torch.Tensor, device: torch.device) -> torch.Tensor:
    """``src.to(device)`` for a large pageable CPU tensor, staged through two pinned buffers."""
    nbytes = src.numel() * src.element_size()
    if src.is_cuda or src.is_pinned() or nbytes < (8 << 20) or not src.is_contiguous():
        return src.to(device, non_blocking=src.is_pinned() if not src.is_cuda else False)
    dst = torch.empty(src.shape, dtype=src.dtype, device=device)
    flat_src, flat_dst = src.view(-1), dst.view(-1)
    step = _CHUNK // src.element_size()
    stream = torch.cuda.current_stream(device)
    with _lock:
        for i, a in enumerate(range(0, flat_src.numel(), step)):
            b = min(flat_src.numel(), a + step)
            stage = _slot(i & 1, _CHUNK)[: (b - a) * src.element_size()].view(src.dtype)
            stage.copy_(flat_src[a:b])                       # host memcpy; the previous chunk's DMA is in flight
            flat_dst[a:b].copy_(stage, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            _events[i & 1] = ev
    return dst


def download_numpy(src: torch.Tensor) -> np.ndarray:
    """``src.cpu().numpy()`` through a pinned staging buffer (a fresh array: the buffer is reused by the next call)."""
    nbytes = src.numel() * src.element_size()
    if not src.is_cuda or nbytes < (1 << 20):
        return src.cpu().numpy()
    src = src.contiguous()
    with _lock:
        stage = _slot(2, nbytes)[:nbytes].view(src.dtype).view(src.shape)
        stage.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(src.device))
        _events[2] = ev
        ev.synchronize()
        return stage.numpy().copy()
