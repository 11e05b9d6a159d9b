"""Memory-axis (THW) sharded read for long videos: one process per GPU over NVLink.

SURVEY.md 8e / BASELINE.json configs[3].  The bank is distributed by frame, round-robin (frame f lives on rank
f % world), so appends stay local and balanced.  One read is

  1. local fused top-k on every rank (scores + local positions; no readout),
  2. exchange of the per-query lists: every rank ends up with all ranks' (local position, score) candidates,
     top_k * 8 bytes per query and rank,
  3. ``evavos_topk_merge_gathered``: global top-k, softmax weights with the global maximum and denominator, and the
     local positions of the winners this rank owns,
  4. local sparse readout of the owned winners (a partial sum, query-major),
  5. sum of the partial readouts, scattered by query slice: rank r ends with queries [r*nq/G, (r+1)*nq/G) of the
     readout in the reference layout (``scatter=True``) - the decoder consumes it where it is - or, after an
     all-gather, with all of it (``scatter=False``, the default).

Two exchange engines:

* ``exchange="peer"`` - device-initiated over NVLink peer memory.  Every rank maps every rank's exchange buffer
  (CUDA IPC, handles passed once over the process group).  Step 2 happens INSIDE the finalizer kernel: as each
  query's list is ready its 400 bytes are stored into every rank's gather region.  Step 5 is a kernel that loads the
  peers' partials straight out of their memory and writes the owned slice.  Two device-side barriers
  (``evavos_peer_barrier``) order the ranks; no host-side collective is on the data path.
* ``exchange="nccl"`` - ``all_gather_into_tensor`` + ``reduce_scatter_tensor`` (the library baseline, and the
  engine of the CPU/gloo tests through injected ``ops``).

The exchange is latency-bound, so queries of several frames are batched into one call
(``qk`` may be (1,CK,F,H,W)).  The compute steps are injectable (``ops``) so the host-side plumbing is testable on
CPU with the gloo backend; the default ops are the CUDA kernels.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from .memory_bank import MemoryBank
from .memory_reader import memory_read


class CudaShardOps:
    """The compute steps of the sharded read on the C ABI (no CPU path)."""

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

    def readout(self, bank: MemoryBank, local_idx, weight, out=None):
        """Partial readout of the owned winners, QUERY-major: (nq, K, CV) fp32 (a query slice is a contiguous chunk)."""
        lib = _lib.load()
        nq, k = local_idx.shape
        if out is None:
            out = torch.empty((nq, bank.K, bank.CV), dtype=torch.float32, device=bank.device)
        sh = bank.shadow()
        with torch.cuda.device(bank.device):
            _lib.check(lib.evavos_readout_qmajor(ctypes.byref(sh), local_idx.data_ptr(), weight.data_ptr(), nq, k,
                                                 out.data_ptr(), _lib.current_stream_ptr(bank.device)))
        return out


def local_to_global(idx_local: torch.Tensor, rank: int, world: int, pos_per_frame: int) -> torch.Tensor:
    """Local bank position -> global position under the round-robin frame distribution (-1 stays -1)."""
    frame = torch.div(idx_local, pos_per_frame, rounding_mode="floor")
    g = (frame * world + rank) * pos_per_frame + (idx_local - frame * pos_per_frame)
    return torch.where(idx_local >= 0, g, idx_local).to(torch.int32)


def query_slice(nq: int, rank: int, world: int):
    """Queries [q0, q1) whose readout rank `rank` owns: equal chunks of ceil(nq / world), the last one short."""
    chunk = (nq + world - 1) // world
    return min(nq, rank * chunk), min(nq, (rank + 1) * chunk)


class _RawDeviceMemory:
    """``__cuda_array_interface__`` view of a device allocation the library owns (lets torch see it as uint8)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def _wrap_device_memory(ptr, nbytes, device):
    with torch.cuda.device(device):
        return torch.as_tensor(_RawDeviceMemory(ptr, nbytes), device=device)


class PeerExchange:
    """The exchange buffers of all ranks, mapped into this process (CUDA IPC), and the device-side barrier epoch.

    Layout of every rank's buffer: [flags: 32 x uint32 | pad to 256] [gather: world x nq x top_k x 2 int32]
    [partial: nq x rows fp32].  Sized for ``max_queries`` queries; grown (collectively) on demand.
    """

    FLAG_BYTES = 256

    def __init__(self, device, group, world, rank, top_k, rows, max_queries):
        lib = _lib.load()
        self.device, self.group, self.world, self.rank = torch.device(device), group, world, rank
        self.top_k, self.rows, self.max_queries = int(top_k), int(rows), int(max_queries)
        self.gather_off = self.FLAG_BYTES
        self.gather_bytes = world * self.max_queries * self.top_k * 8
        self.partial_off = (self.gather_off + self.gather_bytes + 255) // 256 * 256
        self.total = self.partial_off + self.max_queries * self.rows * 4
        self._opened = []
        with torch.cuda.device(self.device):
            # the library allocates (a whole cudaMalloc allocation, zero-filled) and exports the local buffer ...
            ptr, handle = _lib._c_vp(), ctypes.create_string_buffer(64)
            _lib.check(lib.evavos_peer_buffer_alloc(self.total, ctypes.byref(ptr), handle))
            self._local_ptr = ptr.value
            handles = [None] * world
            dist.all_gather_object(handles, (bytes(handle.raw), os.getpid()), group=group)
            # ... and maps every other rank's buffer with THIS GPU current, so that this GPU's kernels can address it
            self.ptrs = []
            for r, (h, _pid) in enumerate(handles):
                if r == rank:
                    self.ptrs.append(self._local_ptr)
                    continue
                p = _lib._c_vp()
                _lib.check(lib.evavos_peer_buffer_open(h, ctypes.byref(p)))
                self._opened.append(p.value)
                self.ptrs.append(p.value)
        self.local = _wrap_device_memory(self._local_ptr, self.total, self.device)
        self.peers = _lib.Peers()
        self.peers.n_ranks, self.peers.rank = world, rank
        for r, p in enumerate(self.ptrs):
            self.peers.base[r] = p
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)      # every rank has mapped every buffer before anybody writes

    def close(self):
        """Unmap the peers' buffers and free the local one (collective: no rank may still be using them)."""
        if getattr(self, "_local_ptr", None) is None:
            return
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.evavos_peer_buffer_close(p)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            lib.evavos_peer_buffer_free(self._local_ptr)
        self._opened, self._local_ptr, self.local = [], None, None

    def barrier(self):
        """Device-side barrier among the ranks, ordered on the current stream."""
        lib = _lib.load()
        self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(lib.evavos_peer_barrier(ctypes.byref(self.peers), 0, self.epoch,
                                               _lib.current_stream_ptr(self.device)))

    def ok(self) -> bool:
        """False once a barrier gave up waiting for a peer (flag slot 31 poisoned)."""
        return int(self.local[31 * 4:32 * 4].view(torch.int32).item()) == 0

    def gather_view(self, nq):
        n = self.world * nq * self.top_k * 8
        return self.local[self.gather_off:self.gather_off + n].view(torch.int32).view(self.world, nq, self.top_k, 2)

    def partial_view(self, nq, k, cv):
        n = nq * k * cv * 4
        return self.local[self.partial_off:self.partial_off + n].view(torch.float32).view(nq, k, cv)


class ShardedMemoryBank:
    """A bank whose frames are spread round-robin over the ranks of ``group``."""

    def __init__(self, num_objects, key_dim, value_dim, height, width, capacity_frames, device, group=None,
                 bank_factory=MemoryBank, ops=None, exchange=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.HW = height * width
        self.K, self.CV, self.H, self.W = num_objects, value_dim, height, width
        local_cap = (capacity_frames + self.world - 1) // self.world
        self.local = bank_factory(num_objects, key_dim, value_dim, height, width, local_cap, device)
        self.device = torch.device(device)
        self.n_frames = 0          # global frame count
        self.ops = ops if ops is not None else CudaShardOps()
        if exchange is None:
            exchange = os.environ.get("EVAVOS_SHARD_EXCHANGE") or None     # (A/B switch for benchmarks)
        if exchange is None:
            exchange = "peer" if (ops is None and self.world > 1 and self.device.type == "cuda") else "nccl"
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        self._peer: PeerExchange | None = None

    @property
    def exchange_desc(self) -> str:
        if self.exchange == "peer":
            return ("device-initiated over NVLink peer memory: lists pushed by the finalizer kernel's epilogue, partial "
                    "readouts pulled by the reduce-scatter kernel, 2 device-side barriers; no NCCL on the data path")
        return "NCCL all_gather_into_tensor(top-k lists) + reduce_scatter_tensor(partial readouts)"

    def close(self):
        """Release the peer exchange buffers (collective; call it on every rank before the process group goes)."""
        if self._peer is not None:
            self._peer.close()
            self._peer = None

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

    # ------------------------------------------------------------------ the read
    def read(self, qk: torch.Tensor, top_k: int = 50, return_topk: bool = False, scatter: bool = False,
             timing: dict | None = None):
        """Readout of ``qk`` against the whole distributed bank.

        scatter=False: replicated (K,CV,[F,]H,W).  scatter=True: (K,CV,q1-q0), the query slice this rank owns
        (``query_slice``).  return_topk adds the merged (global positions, weights), replicated.
        ``timing``: a dict that receives CUDA events around the stages (see ``profile``).
        """
        if self.n_pos < top_k:
            raise RuntimeError(f"selected index k out of range (THW={self.n_pos} < top_k={top_k})")
        spatial = tuple(qk.shape[2:])
        nq = 1
        for s in spatial:
            nq *= int(s)
        if self.world == 1 and isinstance(self.ops, CudaShardOps) and not return_topk:
            out, _ = memory_read(self.local, qk, top_k)      # nothing to exchange: the ordinary fused read
            return out.reshape(self.K, self.CV, nq) if scatter else out
        if self.exchange == "peer" and self.world > 1:
            return self._read_peer(qk, top_k, spatial, nq, return_topk, scatter, timing)
        return self._read_collective(qk, top_k, spatial, nq, return_topk, scatter, timing)

    def _mark(self, timing, name):
        if timing is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.device))
            timing.setdefault("events", []).append((name, e))

    def _local_lists(self, qk, top_k, nq):
        ops = self.ops
        if self.local.n_pos > 0:
            idx_loc, score = ops.local_topk(self.local, qk, top_k)
            k_loc = idx_loc.shape[1]
            if k_loc < top_k:     # a shard with fewer than top_k positions contributes all of them
                pad_i = torch.full((nq, top_k - k_loc), -1, dtype=idx_loc.dtype, device=idx_loc.device)
                pad_s = torch.full((nq, top_k - k_loc), float("-inf"), dtype=score.dtype, device=score.device)
                idx_loc, score = torch.cat([idx_loc, pad_i], 1), torch.cat([score, pad_s], 1)
        else:
            idx_loc = torch.full((nq, top_k), -1, dtype=torch.int32, device=qk.device)
            score = torch.full((nq, top_k), float("-inf"), dtype=torch.float32, device=qk.device)
        return idx_loc, score

    def _read_collective(self, qk, top_k, spatial, nq, return_topk, scatter, timing):
        ops, rank, world = self.ops, self.rank, self.world
        self._mark(timing, "start")
        idx_loc, score = self._local_lists(qk, top_k, nq)
        self._mark(timing, "local_topk")
        if hasattr(ops, "merge_gathered"):
            # one collective: packed (local position, score bits) pairs; the kernel maps them to global positions
            packed = torch.stack([idx_loc, score.view(torch.int32)], -1).contiguous()
            if world > 1:
                gathered = torch.empty((world,) + tuple(packed.shape), dtype=torch.int32, device=packed.device)
                dist.all_gather_into_tensor(gathered, packed, group=self.group)
            else:
                gathered = packed.unsqueeze(0)
            self._mark(timing, "gather")
            glob_idx, weight, local_idx = ops.merge_gathered(gathered, top_k, rank, world, self.HW)
        else:
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
            self._mark(timing, "gather")
            glob_idx, weight, local_idx = ops.merge(cand_idx, cand_score, top_k, rank, world, self.HW)
        self._mark(timing, "merge")
        # partial readout, query-major and padded so that every rank's slice is one equal chunk
        chunk = (nq + world - 1) // world
        part = torch.zeros((chunk * world, self.K, self.CV), dtype=torch.float32, device=qk.device)
        if self.local.n_pos > 0:     # ops.readout -> (nq, K, CV), query-major
            got = ops.readout(self.local, local_idx, weight, out=part[:nq]) if _accepts_out(ops) else \
                ops.readout(self.local, local_idx, weight)
            if got.data_ptr() != part.data_ptr():
                part[:nq] = got.reshape(nq, self.K, self.CV)
        self._mark(timing, "readout")
        q0, q1 = query_slice(nq, rank, world)
        if world > 1 and part.is_cuda:
            mine = torch.empty((chunk, self.K, self.CV), dtype=torch.float32, device=qk.device)
            dist.reduce_scatter_tensor(mine, part, op=dist.ReduceOp.SUM, group=self.group)
        elif world > 1:      # gloo (CPU tests of the plumbing) has no reduce-scatter: all-reduce and keep the slice
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
            mine = part[rank * chunk:(rank + 1) * chunk]
        else:
            mine = part
        self._mark(timing, "reduce")
        if scatter:
            out = mine[:q1 - q0].permute(1, 2, 0).contiguous()                     # (K, CV, q1 - q0)
        else:
            if world > 1:
                full = torch.empty((chunk * world, self.K, self.CV), dtype=torch.float32, device=qk.device)
                dist.all_gather_into_tensor(full, mine, group=self.group)
            else:
                full = mine
            out = full[:nq].permute(1, 2, 0).contiguous().view(self.K, self.CV, *spatial)
        return (out, glob_idx, weight) if return_topk else out

    def _read_peer(self, qk, top_k, spatial, nq, return_topk, scatter, timing):
        lib = _lib.load()
        rank, world, dev = self.rank, self.world, self.device
        rows = self.K * self.CV
        px = self._peer
        if px is None or px.top_k != top_k or px.max_queries < nq or px.rows != rows:
            cap = max(nq, px.max_queries if px is not None else 0)
            if px is not None:
                px.close()
            self._peer = px = PeerExchange(dev, self.group, world, rank, top_k, rows, cap)
        stream = _lib.current_stream_ptr(dev)
        self._mark(timing, "start")
        # 1 + 2: local selection; the finalizer's epilogue stores every query's list into every rank's gather region
        if self.local.n_pos >= top_k:
            memory_read(self.local, qk, top_k, want_readout=False, want_topk=False, peers=px.peers,
                        peer_gather_offset=px.gather_off)
        else:   # (a shard with fewer than top_k positions: selection on the host path, then a plain push)
            idx_loc, score = self._local_lists(qk, top_k, nq)
            packed = torch.stack([idx_loc, score.view(torch.int32)], -1).contiguous()
            for r in range(world):      # plain copies through the mappings (device-to-device, peer or local)
                t = _wrap_device_memory(px.ptrs[r] + px.gather_off + rank * nq * top_k * 8, nq * top_k * 8, dev)
                t.view(torch.int32).view(nq, top_k, 2).copy_(packed)
        self._mark(timing, "local_topk+push")
        px.barrier()
        self._mark(timing, "barrier1")
        # 3: merge what the peers pushed
        glob_idx, weight, local_idx = self.ops.merge_gathered(px.gather_view(nq), top_k, rank, world, self.HW)
        self._mark(timing, "merge")
        # 4: partial readout into the exchange buffer (query-major)
        part = px.partial_view(nq, self.K, self.CV)
        if self.local.n_pos > 0:
            self.ops.readout(self.local, local_idx, weight, out=part)
        else:
            part.zero_()
        self._mark(timing, "readout")
        px.barrier()
        self._mark(timing, "barrier2")
        # 5: sum of the ranks' partials for the owned query slice, loaded straight from the peers' memory
        q0, q1 = query_slice(nq, rank, world)
        mine = torch.empty((self.K, self.CV, max(q1 - q0, 0)), dtype=torch.float32, device=dev)
        if q1 > q0:
            with torch.cuda.device(dev):
                _lib.check(lib.evavos_peer_reduce_scatter(ctypes.byref(px.peers), px.partial_off, rows, q0, q1,
                                                          mine.data_ptr(), q1 - q0, stream))
        self._mark(timing, "reduce")
        if scatter:
            out = mine
        else:
            chunk = (nq + world - 1) // world
            padded = torch.zeros((self.K, self.CV, chunk), dtype=torch.float32, device=dev)
            padded[:, :, :q1 - q0] = mine
            full = torch.empty((world, self.K, self.CV, chunk), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(full, padded, group=self.group)
            out = full.permute(1, 2, 0, 3).reshape(self.K, self.CV, world * chunk)[:, :, :nq].contiguous() \
                .view(self.K, self.CV, *spatial)
        return (out, glob_idx, weight) if return_topk else out

    # ------------------------------------------------------------------ diagnostics
    def profile(self, qk, top_k=50, reps=10):
        """Mean microseconds per stage of ``read(scatter=True)`` on this rank (CUDA events between the stages)."""
        acc, n = {}, 0
        for _ in range(reps):
            t = {}
            self.read(qk, top_k, scatter=True, timing=t)
            torch.cuda.synchronize(self.device)
            ev = t.get("events", [])
            for (_, a), (name, b) in zip(ev[:-1], ev[1:]):
                acc[name] = acc.get(name, 0.0) + a.elapsed_time(b) * 1e3
            n += 1
        return {k: v / max(n, 1) for k, v in acc.items()}


_SUBGROUPS: dict = {}     # (parent group, memory_shards) -> the sub-groups of HybridShardedBank


class HybridShardedBank:
    """world = query_groups x memory_shards: the bank is sharded along the memory axis over the ``memory_shards`` ranks
    of a group (a ShardedMemoryBank on a sub-group) and replicated across the groups; every group answers its own
    contiguous slice of the queries.

    Why: what a rank pays per QUERY in the sharded read - thresholds, candidate lists, the finalizer, the merge, the
    barriers - does not shrink with the memory shard, so G-way memory sharding of a fixed read
    stops scaling early.  Splitting the queries as well divides exactly those costs, at the price of 1/memory_shards
    instead of 1/world of the bank per GPU.  ``memory_shards = world`` is the plain memory-axis sharded read,
    ``memory_shards = 1`` pure query parallelism over replicated banks (no exchange at all).

    ``read`` returns (K, CV, q1 - q0): the readout of queries [q0, q1) = ``owned_slice(nq)`` of the flattened query
    axis, in the reference layout - the scattered form the decoder consumes where it is.
    """

    def __init__(self, num_objects, key_dim, value_dim, height, width, capacity_frames, device, memory_shards,
                 group=None, **kw):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        m = int(memory_shards)
        if m < 1 or world % m != 0:
            raise ValueError(f"memory_shards={memory_shards} must divide the world size {world}")
        self.world, self.rank, self.memory_shards, self.query_groups = world, rank, m, world // m
        self.qgroup, self.mshard = rank // m, rank % m
        self.CK = key_dim
        sub = None
        if world > 1:
            # (keyed on the live default group as well: sub-groups die with the process group that made them)
            key = (id(dist.distributed_c10d._get_default_group()), id(group) if group is not None else 0, m)
            subs = _SUBGROUPS.get(key)
            if subs is None:                         # created once per (parent group, M) and shared by all banks
                parent = dist.get_process_group_ranks(group) if group is not None else list(range(world))
                # new_group is collective: every rank creates every sub-group, in the same order
                subs = _SUBGROUPS[key] = [dist.new_group([parent[g * m + j] for j in range(m)])
                                          for g in range(self.query_groups)]
            sub = subs[self.qgroup]
        self.subgroup = sub
        self.bank = ShardedMemoryBank(num_objects, key_dim, value_dim, height, width, capacity_frames, device,
                                      group=sub, **kw)

    def append(self, key_frame, value_frame) -> int:
        """Called on every rank with the same frame: every group stores it, on its rank frame % memory_shards."""
        return self.bank.append(key_frame, value_frame)

    @property
    def n_pos(self) -> int:
        return self.bank.n_pos

    @property
    def exchange_desc(self) -> str:
        return self.bank.exchange_desc if self.memory_shards > 1 else "none (replicated banks, queries split)"

    def query_range(self, nq: int):
        """Queries [a, b) this rank's group answers."""
        return query_slice(nq, self.qgroup, self.query_groups)

    def owned_slice(self, nq: int):
        """Queries [q0, q1) whose readout this rank ends up with."""
        a, b = self.query_range(nq)
        q0, q1 = query_slice(b - a, self.mshard, self.memory_shards)
        return a + q0, a + q1

    def read(self, qk: torch.Tensor, top_k: int = 50, return_topk: bool = False, timing: dict | None = None):
        flat = qk.reshape(1, qk.shape[1], -1)
        a, b = self.query_range(flat.shape[2])
        return self.bank.read(flat[:, :, a:b], top_k, return_topk=return_topk, scatter=True, timing=timing)

    def profile(self, qk, top_k=50, reps=10):
        flat = qk.reshape(1, qk.shape[1], -1)
        a, b = self.query_range(flat.shape[2])
        if self.memory_shards == 1:
            return {}
        return self.bank.profile(flat[:, :, a:b], top_k, reps)

    def close(self):
        self.bank.close()


def _accepts_out(ops) -> bool:
    import inspect
    try:
        return "out" in inspect.signature(ops.readout).parameters
    except (TypeError, ValueError):
        return False
