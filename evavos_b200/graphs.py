"""CUDA-graph replay of the fixed-shape conv passes around the memory read.

With the kernels fused and the BatchNorms folded, a 32-frame 480p video is ~1 500 launches for ~49 ms of device time
but ~59 ms of host time: PyTorch's per-op dispatch (autocast, module calls) is the bound.  The key encoder, the decoder
and the value encoder see the same few shapes for a whole video (and for every video of a dataset), so each is
captured once per input signature and replayed: one launch per pass.  The memory read itself is NOT captured - its
arguments (bank length, workspace) change with every append.

A captured pass owns static input / output tensors: ``GraphedPass.__call__`` copies the arguments in, replays and
returns the static outputs, which the next replay of the same signature overwrites - callers that keep a result
(the key-feature cache) ask for clones.
"""
from __future__ import annotations

import torch


class GraphedPass:
    """``fn(*tensors) -> tensor | tuple of tensors``, one CUDA graph per input signature (shapes, dtypes, layouts)."""

    def __init__(self, fn, pool=None):
        self.fn = fn
        self.pool = pool
        self.entries: dict = {}

    @staticmethod
    def _sig(inputs):
        return tuple((tuple(t.shape), t.dtype, t.stride()) for t in inputs)

    def static_inputs(self, *like):
        """The static input tensors for arguments shaped like ``like`` (captured on first use): a producer may write
        into them directly - e.g. the memory read into the decoder's input - and then call ``replay``."""
        return self._entry(like)[1]

    def _entry(self, inputs):
        sig = self._sig(inputs)
        ent = self.entries.get(sig)
        if ent is None:
            ent = self.entries[sig] = self._capture(inputs)
        return ent

    def _capture(self, inputs):
        static_in = [torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device) for t in inputs]
        for s, t in zip(static_in, inputs):
            s.copy_(t)
        side = torch.cuda.Stream(device=inputs[0].device)
        side.wait_stream(torch.cuda.current_stream(inputs[0].device))
        with torch.cuda.stream(side):          # cuDNN autotuning, lazy workspaces and allocator growth happen here
            for _ in range(2):
                self.fn(*static_in)
        torch.cuda.current_stream(inputs[0].device).wait_stream(side)
        torch.cuda.synchronize(inputs[0].device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self.pool):
            out = self.fn(*static_in)
        return graph, static_in, out

    def replay(self, *like):
        """Replay the pass captured for this signature on whatever its static inputs hold now."""
        graph, _, out = self._entry(like)
        graph.replay()
        return out

    def __call__(self, *inputs, clone=False):
        graph, static_in, out = self._entry(inputs)
        for s, t in zip(static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t)
        graph.replay()
        if clone:
            return out.clone() if isinstance(out, torch.Tensor) else tuple(o.clone() for o in out)
        return out
