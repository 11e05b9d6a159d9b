"""Memory-axis (THW) sharded read for long videos: one process per GPU, NCCL over NVLink.

SURVEY.md 8e / BASELINE.json configs[3].  The bank is distributed by frame, round-robin
(frame f lives on rank f % world), so appends stay local and balanced.  One read is

  1. local fused top-k on every rank (scores + local positions; no readout),
  2. all-gather of the (score, GLOBAL position) candidates: top_k * 8 bytes per query and rank,
  3. ``evavos_topk_merge``: global top-k, softmax weights with the global maximum and denominator,
     and the local positions of the winners this rank owns,
  4. local sparse readout of the owned winners (a partial sum),
  5. all-reduce (sum) of the partial readouts: K*CV*HW*4 bytes.

The exchange is latency-bound, so queries of several frames are batched into
one call (``qk`` may be (1,CK,F,H,W)).  The compute steps are injectable (``ops``) so the host-side
plumbing is testable on CPU with the gloo backend; the default ops are the CUDA kernels.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .memory_bank import MemoryBank
from .memory_reader import memory_read


class CudaShardOps:
    """The three compute steps of the sharded read on the C ABI (no CPU path)."""

    def local_topk(self, bank: MemoryBank, qk: torch.Tensor, top_k: int):
        """-> (local positions int32 (nq,k_loc), scores f32 (nq,k_loc)), k_loc = min(top_k, local positions)."""
        k_loc = min(top_k, bank.n_pos)
        _, aff = memory_read(bank, qk, k_loc, want_readout=False, want_topk=True)
        return aff.idx, aff.score

    def merge(self, cand_idx, cand_score, top_k, rank, world, pos_per_frame):
        lib = _lib.load()
        nq, n_cand = cand_idx.shape
        dev = cand_idx.device
        out_idx = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
        weight = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        local_idx = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.evavos_topk_merge(cand_idx.data_ptr(), cand_score.data_ptr(), nq, n_cand, top_k, rank, world,
                                             pos_per_frame, out_idx.data_ptr(), weight.data_ptr(), None,
                                             local_idx.data_ptr(), _lib.current_stream_ptr(dev)))
        return out_idx, weight, local_idx

    def merge_gathered(self, gathered, top_k, rank, world, pos_per_frame):
        """gathered: (world, nq, per_shard, 2) int32 = all-gather of packed (local position, score bits)."""
        lib = _lib.load()
        _, nq, per_shard, _ = gathered.shape
        dev = gathered.device
        out_idx = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
        weight = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        local_idx = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.evavos_topk_merge_gathered(gathered.data_ptr(), nq, per_shard, top_k, rank, world,
                                                      pos_per_frame, out_idx.data_ptr(), weight.data_ptr(), None,
                                                      local_idx.data_ptr(), _lib.current_stream_ptr(dev)))
        return out_idx, weight, local_idx

    def readout(self, bank: MemoryBank, local_idx, weight):
        lib = _lib.load()
        nq, k = local_idx.shape
        out = torch.empty((bank.K, bank.CV, nq), dtype=torch.float32, device=bank.device)
        sh = bank.shadow()
        with torch.cuda.device(bank.device):
            _lib.check(lib.evavos_readout(ctypes.byref(sh), local_idx.data_ptr(), weight.data_ptr(), nq, k,
                                          out.data_ptr(), 0, 0, _lib.current_stream_ptr(bank.device)))
        return out


def local_to_global(idx_local: torch.Tensor, rank: int, world: int, pos_per_frame: int) -> torch.Tensor:
    """Local bank position -> global position under the round-robin frame distribution (-1 stays -1)."""
    frame = torch.div(idx_local, pos_per_frame, rounding_mode="floor")
    g = (frame * world + rank) * pos_per_frame + (idx_local - frame * pos_per_frame)
    return torch.where(idx_local >= 0, g, idx_local).to(torch.int32)


class ShardedMemoryBank:
    """A bank whose frames are spread round-robin over the ranks of ``group``."""

    def __init__(self, num_objects, key_dim, value_dim, height, width, capacity_frames, device, group=None,
                 bank_factory=MemoryBank, ops=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.HW = height * width
        self.K, self.CV, self.H, self.W = num_objects, value_dim, height, width
        local_cap = (capacity_frames + self.world - 1) // self.world
        self.local = bank_factory(num_objects, key_dim, value_dim, height, width, local_cap, device)
        self.n_frames = 0          # global frame count
        self.ops = ops if ops is not None else CudaShardOps()

    def owner_of(self, frame: int) -> int:
        return frame % self.world

    def append(self, key_frame, value_frame) -> int:
        """Called on every rank with the same frame; only the owner stores it.  Returns the global slot."""
        slot = self.n_frames
        if self.owner_of(slot) == self.rank:
            self.local.append(key_frame, value_frame)
        self.n_frames += 1
        return slot

    @property
    def n_pos(self) -> int:
        return self.n_frames * self.HW

    def read(self, qk: torch.Tensor, top_k: int = 50, return_topk: bool = False):
        """Replicated (K,CV,[F,]H,W) readout of ``qk`` against the whole distributed bank."""
        if self.n_pos < top_k:
            raise RuntimeError(f"selected index k out of range (THW={self.n_pos} < top_k={top_k})")
        spatial = tuple(qk.shape[2:])
        ops, rank, world = self.ops, self.rank, self.world
        if world == 1 and isinstance(ops, CudaShardOps) and not return_topk:
            out, _ = memory_read(self.local, qk, top_k)      # nothing to exchange: the ordinary fused read
            return out
        if self.local.n_pos > 0:
            idx_loc, score = ops.local_topk(self.local, qk, top_k)
            nq, k_loc = idx_loc.shape
            if k_loc < top_k:     # a shard with fewer than top_k positions contributes all of them
                pad_i = torch.full((nq, top_k - k_loc), -1, dtype=idx_loc.dtype, device=idx_loc.device)
                pad_s = torch.full((nq, top_k - k_loc), float("-inf"), dtype=score.dtype, device=score.device)
                idx_loc, score = torch.cat([idx_loc, pad_i], 1), torch.cat([score, pad_s], 1)
        else:
            nq = int(torch.tensor(qk.shape[2:]).prod())
            idx_loc = torch.full((nq, top_k), -1, dtype=torch.int32, device=qk.device)
            score = torch.full((nq, top_k), float("-inf"), dtype=torch.float32, device=qk.device)
        if hasattr(ops, "merge_gathered"):
            # one collective: packed (local position, score bits) pairs; the kernel maps them to global positions
            packed = torch.stack([idx_loc, score.view(torch.int32)], -1).contiguous()
            if world > 1:
                gathered = torch.empty((world,) + tuple(packed.shape), dtype=torch.int32, device=packed.device)
                dist.all_gather_into_tensor(gathered, packed, group=self.group)
            else:
                gathered = packed.unsqueeze(0)
            glob_idx, weight, local_idx = ops.merge_gathered(gathered, top_k, rank, world, self.HW)
            return self._finish(qk, spatial, nq, glob_idx, weight, local_idx, return_topk)
        idx_glob = local_to_global(idx_loc, rank, world, self.HW).contiguous()
        score = score.contiguous()
        if world > 1:
            gi = [torch.empty_like(idx_glob) for _ in range(world)]
            gs = [torch.empty_like(score) for _ in range(world)]
            dist.all_gather(gi, idx_glob, group=self.group)
            dist.all_gather(gs, score, group=self.group)
            cand_idx = torch.cat(gi, 1).contiguous()       # (nq, world * top_k), rank-major
            cand_score = torch.cat(gs, 1).contiguous()
        else:
            cand_idx, cand_score = idx_glob, score
        glob_idx, weight, local_idx = ops.merge(cand_idx, cand_score, top_k, rank, world, self.HW)
        return self._finish(qk, spatial, nq, glob_idx, weight, local_idx, return_topk)

    def _finish(self, qk, spatial, nq, glob_idx, weight, local_idx, return_topk):
        ops, world = self.ops, self.world
        if self.local.n_pos > 0:
            part = ops.readout(self.local, local_idx, weight)
        else:
            part = torch.zeros((self.K, self.CV, nq), dtype=torch.float32, device=qk.device)
        if world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        out = part.view(self.K, self.CV, *spatial)
        return (out, glob_idx, weight) if return_topk else out
