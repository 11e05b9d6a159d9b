"""CPU, world_size 2 and 3, gloo: host-side plumbing of the memory-axis sharded read.

The CUDA steps are replaced by oracle-backed test doubles (injected through ``ops`` / ``bank_factory``);
what is under test is the frame distribution, local->global index mapping, the all-gather layout, ownership
and the partial-sum all-reduce.  The merged result must equal the single-bank oracle read.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import memread_np as onp


class _CpuBank:
    """Minimal stand-in for MemoryBank on the CPU (test double)."""

    def __init__(self, K, CK, CV, H, W, capacity_frames, device):
        self.K, self.CK, self.CV, self.H, self.W, self.HW = K, CK, CV, H, W, H * W
        self.device = torch.device("cpu")
        self.keys = np.zeros((CK, capacity_frames * self.HW), np.float32)
        self.vals = np.zeros((K, CV, capacity_frames * self.HW), np.float32)
        self.n_frames, self.capacity_frames = 0, capacity_frames

    @property
    def n_pos(self):
        return self.n_frames * self.HW

    def append(self, key_frame, value_frame):
        assert self.n_frames < self.capacity_frames
        s = slice(self.n_pos, self.n_pos + self.HW)
        self.keys[:, s] = key_frame.reshape(self.CK, self.HW).numpy()
        self.vals[:, :, s] = value_frame.reshape(self.K, self.CV, self.HW).numpy()
        self.n_frames += 1


class _OracleOps:
    def local_topk(self, bank, qk, top_k):
        k_loc = min(top_k, bank.n_pos)
        s = onp.affinity_scores(bank.keys[:, :bank.n_pos], qk.reshape(bank.CK, -1).numpy())
        tk = onp.topk_softmax(s, k_loc)
        return torch.from_numpy(tk.idx.astype(np.int32)), torch.from_numpy(tk.score.astype(np.float32))

    def merge(self, cand_idx, cand_score, top_k, rank, world, ppf):
        ci, cs = cand_idx.numpy().astype(np.int64), cand_score.numpy().astype(np.float64)
        cs = np.where(ci >= 0, cs, -np.inf)
        order = np.lexsort((ci, -cs), axis=1)[:, :top_k]
        gi, gs = np.take_along_axis(ci, order, 1), np.take_along_axis(cs, order, 1)
        e = np.exp(gs - gs[:, :1])
        w = e / e.sum(1, keepdims=True)
        frame, r = gi // ppf, gi % ppf
        loc = np.where(frame % world == rank, (frame // world) * ppf + r, -1)
        return (torch.from_numpy(gi.astype(np.int32)), torch.from_numpy(w.astype(np.float32)),
                torch.from_numpy(loc.astype(np.int32)))

    def readout(self, bank, local_idx, weight):
        li, w = local_idx.numpy().astype(np.int64), weight.numpy().astype(np.float64)
        w = np.where(li >= 0, w, 0.0)
        out = onp.readout(np.maximum(li, 0), w, bank.vals[:, :, :max(bank.n_pos, 1)])       # (K, CV, nq)
        return torch.from_numpy(np.ascontiguousarray(out.transpose(2, 0, 1)).astype(np.float32))  # query-major


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, frames, top_k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from evavos_b200.sharded import ShardedMemoryBank, local_to_global
        K, CK, CV, H, W = 2, 64, 12, 5, 6
        g = torch.Generator().manual_seed(99)
        mk = torch.randn(1, CK, frames, H, W, generator=g)
        mv = torch.randn(K, CV, frames, H, W, generator=g)
        qk = torch.randn(1, CK, 2, H, W, generator=g)          # two query frames in one exchange
        bank = ShardedMemoryBank(K, CK, CV, H, W, frames, "cpu", bank_factory=_CpuBank, ops=_OracleOps())
        for f in range(frames):
            assert bank.append(mk[:, :, f], mv[:, :, f:f + 1]) == f
        assert bank.local.n_frames == len(range(rank, frames, world))
        out, gidx, w = bank.read(qk, top_k, return_topk=True)
        tk, ro = onp.memory_read(mk[0].reshape(CK, -1).numpy(), qk[0].reshape(CK, -1).numpy(),
                                 mv.reshape(K, CV, -1).numpy(), top_k)
        assert out.shape == (K, CV, 2, H, W)
        assert (gidx.numpy() == tk.idx).all(), "merged global top-k differs from the single-bank oracle"
        assert np.abs(w.numpy() - tk.weight).max() < 1e-6
        assert np.abs(out.reshape(K, CV, -1).numpy() - ro).max() < 1e-5
        # scattered form: every rank keeps the query slice it owns
        from evavos_b200.sharded import query_slice
        mine = bank.read(qk, top_k, scatter=True)
        q0, q1 = query_slice(2 * H * W, rank, world)
        assert mine.shape == (K, CV, q1 - q0)
        assert np.abs(mine.numpy() - ro[:, :, q0:q1]).max() < 1e-5
        # index mapping round trip
        loc = torch.arange(bank.local.n_pos, dtype=torch.int32)
        glob = local_to_global(loc, rank, world, H * W)
        assert ((glob // (H * W)) % world == rank).all()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,frames,top_k", [(2, 7, 50), (3, 4, 50), (2, 3, 20)])
def test_sharded_read_matches_single_bank(world, frames, top_k):
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, top_k, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))


def _hybrid_worker(rank, world, port, memory_shards, frames, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from evavos_b200.sharded import HybridShardedBank
        K, CK, CV, H, W, top_k = 2, 64, 12, 5, 6, 20
        g = torch.Generator().manual_seed(7)
        mk = torch.randn(1, CK, frames, H, W, generator=g)
        mv = torch.randn(K, CV, frames, H, W, generator=g)
        qk = torch.randn(1, CK, 3, H, W, generator=g)
        bank = HybridShardedBank(K, CK, CV, H, W, frames, "cpu", memory_shards, bank_factory=_CpuBank, ops=_OracleOps())
        assert (bank.query_groups, bank.qgroup, bank.mshard) == (world // memory_shards, rank // memory_shards,
                                                                 rank % memory_shards)
        for f in range(frames):
            bank.append(mk[:, :, f], mv[:, :, f:f + 1])
        # every group holds the whole bank, split over its memory shards
        assert bank.bank.local.n_frames == len(range(bank.mshard, frames, memory_shards))
        nq = 3 * H * W
        out, gidx, w = bank.read(qk, top_k, return_topk=True)
        tk, ro = onp.memory_read(mk[0].reshape(CK, -1).numpy(), qk[0].reshape(CK, -1).numpy(),
                                 mv.reshape(K, CV, -1).numpy(), top_k)
        a, b = bank.query_range(nq)
        q0, q1 = bank.owned_slice(nq)
        assert a <= q0 <= q1 <= b and out.shape == (K, CV, q1 - q0)
        assert (gidx.numpy() == tk.idx[a:b]).all()
        assert np.abs(out.numpy() - ro[:, :, q0:q1]).max() < 1e-5
        # the owned slices of all ranks tile the query axis exactly once
        spans = [None] * world
        dist.all_gather_object(spans, (q0, q1))
        covered = sorted(spans)
        assert covered[0][0] == 0 and covered[-1][1] == nq and all(x[1] == y[0] for x, y in zip(covered, covered[1:]))
        bank.close()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,memory_shards,frames", [(2, 1, 5), (2, 2, 5), (4, 2, 6)])
def test_hybrid_query_groups_times_memory_shards(world, memory_shards, frames):
    """world = query groups x memory shards (HybridShardedBank): every group answers its slice of the queries against
    a bank sharded over its own ranks; together the ranks hold every query's readout exactly once."""
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, world, port, memory_shards, frames, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))
