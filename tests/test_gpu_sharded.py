"""GPU: the merge step of the memory-axis sharded read (CUDA kernels), on one device and over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import memread_np as onp
from tests.helpers import synth

pytestmark = pytest.mark.gpu


def test_two_shards_on_one_device_equal_single_bank():
    """Both shards computed on cuda:0: local top-k -> merge kernel -> partial readouts sum to the full read."""
    import evavos_b200 as ev
    from evavos_b200.sharded import CudaShardOps, local_to_global
    dev = torch.device("cuda:0")
    K, CK, CV, T, H, W, top_k, world = 2, 64, 64, 7, 8, 10, 50, 2
    mk, qk, mv = synth(5, CK, CV, T, H, W, K)
    full = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    ref, aff = ev.memory_read(full, qk.to(dev), top_k, want_topk=True)
    ops = CudaShardOps()
    shards = []
    for r in range(world):
        frames = list(range(r, T, world))
        b = ev.MemoryBank(K, CK, CV, H, W, len(frames), dev)
        for f in frames:
            b.append(mk[:, :, f].to(dev), mv[:, :, f:f + 1].to(dev))
        shards.append(b)
    cands = []
    for r, b in enumerate(shards):
        idx, sc = ops.local_topk(b, qk.to(dev), top_k)
        cands.append((local_to_global(idx, r, world, H * W), sc))
    cand_idx = torch.cat([c[0] for c in cands], 1).contiguous()
    cand_sc = torch.cat([c[1] for c in cands], 1).contiguous()
    total = None
    for r, b in enumerate(shards):
        gidx, w, loc = ops.merge(cand_idx, cand_sc, top_k, r, world, H * W)
        part = ops.readout(b, loc, w)
        total = part if total is None else total + part
        assert (gidx == aff.idx).all()
        assert (w - aff.weight).abs().max() < 1e-6
        owned = (loc >= 0).sum(1)
    total = total.permute(1, 2, 0).contiguous()          # the partial readouts are query-major (nq, K, CV)
    assert (total.view_as(ref) - ref).abs().max() < 1e-5
    assert int(owned.max()) <= top_k
    # the packed all-gather layout [shard][query][k][2] gives the same merge
    packed = torch.stack([torch.stack([ops.local_topk(b, qk.to(dev), top_k)[0],
                                       ops.local_topk(b, qk.to(dev), top_k)[1].view(torch.int32)], -1) for b in shards], 0)
    total2 = None
    for r, b in enumerate(shards):
        gidx2, w2, loc2 = ops.merge_gathered(packed.contiguous(), top_k, r, world, H * W)
        assert (gidx2 == aff.idx).all() and (w2 - aff.weight).abs().max() < 1e-6
        part = ops.readout(b, loc2, w2)
        total2 = part if total2 is None else total2 + part
    assert torch.equal(total2.permute(1, 2, 0).contiguous(), total)
    # against the oracle too
    tk, ro = onp.memory_read(mk[0].reshape(CK, -1).numpy(), qk[0].reshape(CK, -1).numpy(), mv.reshape(K, CV, -1).numpy(), top_k)
    assert onp.rel_l2(total.cpu().numpy().reshape(ro.shape), ro) < 1e-5


def _sharded_worker(rank, world, port, ret, backend, same_device, degenerate=False):
    """One rank of a sharded read.  backend "nccl": one GPU per rank, both exchange engines.  backend "gloo" with
    same_device: all ranks share cuda:0 (CUDA IPC works between processes on one device) - the peer-memory engine
    needs no NCCL at all, only a process group to pass the IPC handles, so it can be checked on a 1-GPU box."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dev = torch.device("cuda", 0 if same_device else rank)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from evavos_b200.sharded import ShardedMemoryBank, query_slice
        K, CK, CV, T, H, W = 2, 64, 512, 9, 12, 16
        mk, _, mv = synth(17, CK, CV, T, H, W, K)
        qk = torch.randn(1, CK, 3, H, W, generator=torch.Generator().manual_seed(18))     # 3 query frames, 576 queries
        if degenerate:
            # near-constant keys: every local candidate list overflows (> 1 024 of a shard's ~5 800 positions inside the
            # filter's margin), so the lists the peers receive come from the finalizer's own exact path on the first
            # read and from overflow_exact_kernel's finalizer afterwards (the overflow hint of api.cu)
            T, H, W = 12, 30, 32
            g = torch.Generator().manual_seed(19)
            base = torch.randn(1, CK, 1, 1, 1, generator=g)
            mk = base + 1e-3 * torch.randn(1, CK, T, H, W, generator=g)
            mv = torch.randn(K, CV, T, H, W, generator=g)
            qk = (0.9 * base + 1e-3 * torch.randn(1, CK, 1, H, W, generator=g))[:, :, :, :8].contiguous()      # 256 queries
        n_queries = qk.shape[2] * qk.shape[3] * qk.shape[4]
        tk, ro = onp.memory_read(mk[0].reshape(CK, -1).numpy(), qk[0].reshape(CK, -1).numpy(), mv.reshape(K, CV, -1).numpy(), 50)
        engines = ["peer", "nccl"] if backend == "nccl" else ["peer"]
        for engine in engines:
            bank = ShardedMemoryBank(K, CK, CV, H, W, T, dev, exchange=engine)
            for f in range(T):
                bank.append(mk[:, :, f].to(dev), mv[:, :, f:f + 1].to(dev))
            for rep in range(3):      # several reads through the same exchange buffers (barrier epochs, buffer reuse)
                mine, gidx, w = bank.read(qk.to(dev), 50, return_topk=True, scatter=True)
                torch.cuda.synchronize()
                q0, q1 = query_slice(n_queries, rank, world)
                assert mine.shape == (K, CV, q1 - q0)
                if degenerate:       # near-ties: compare the selection tie-aware, the readout against the oracle's
                    s64 = onp.affinity_scores(mk[0].reshape(CK, -1).numpy(), qk[0].reshape(CK, -1).numpy())
                    from tests.helpers import TIE_TOL
                    assert onp.compare_topk(gidx.cpu().numpy(), s64, 50, TIE_TOL)[2] == 0
                    own = onp.readout(gidx.cpu().numpy().astype(np.int64), w.cpu().numpy().astype(np.float64),
                                      mv.reshape(K, CV, -1).numpy())
                    assert onp.rel_l2(mine.cpu().numpy(), own[:, :, q0:q1]) < 1e-5
                    continue
                assert (gidx.cpu().numpy() == tk.idx).all(), engine
                assert np.abs(w.cpu().numpy() - tk.weight).max() < 1e-6
                err = onp.rel_l2(mine.cpu().numpy(), ro[:, :, q0:q1])
                assert err < 1e-5, (engine, rep, err)
            if bank._peer is not None:
                assert bank._peer.ok(), "a peer barrier timed out"
            if backend == "nccl":      # replicated form (all-gather of the slices)
                out = bank.read(qk.to(dev), 50)
                assert onp.rel_l2(out.cpu().numpy().reshape(ro.shape), ro) < 1e-5
            bank.close()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def _run_ranks(world, backend, same_device, degenerate=False):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, ret, backend, same_device, degenerate))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))


def test_peer_memory_exchange_two_ranks_on_one_gpu():
    """The device-initiated exchange (finalizer push over IPC-mapped peer buffers, device-side barriers, reduce-scatter
    by peer loads) with two processes sharing cuda:0: no NCCL involved, so it runs on the driver's 1-GPU box."""
    _run_ranks(2, "gloo", True)


def test_peer_memory_exchange_with_overflowed_lists():
    """Near-constant keys: the pushed lists come from the exact paths (finalizer warp, then the tiled overflow pass)."""
    _run_ranks(2, "gloo", True, degenerate=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_read_over_nvlink_two_gpus():
    """Both exchange engines (peer memory, NCCL) with one GPU per rank."""
    _run_ranks(2, "nccl", False)


def _hybrid_worker(rank, world, port, ret, memory_shards):
    """Four (or two) processes on cuda:0: world / M query groups x M memory shards, peer-memory exchange inside each
    group (gloo only carries the IPC handles)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from evavos_b200.sharded import HybridShardedBank
        K, CK, CV, T, H, W = 2, 64, 512, 9, 12, 16
        mk, _, mv = synth(27, CK, CV, T, H, W, K)
        qk = torch.randn(1, CK, 3, H, W, generator=torch.Generator().manual_seed(28))
        nq = 3 * H * W
        tk, ro = onp.memory_read(mk[0].reshape(CK, -1).numpy(), qk[0].reshape(CK, -1).numpy(), mv.reshape(K, CV, -1).numpy(), 50)
        bank = HybridShardedBank(K, CK, CV, H, W, T, dev, memory_shards)
        for f in range(T):
            bank.append(mk[:, :, f].to(dev), mv[:, :, f:f + 1].to(dev))
        a, b = bank.query_range(nq)
        q0, q1 = bank.owned_slice(nq)
        for rep in range(2):
            mine, gidx, w = bank.read(qk.to(dev), 50, return_topk=True)
            torch.cuda.synchronize()
            assert mine.shape == (K, CV, q1 - q0)
            assert (gidx.cpu().numpy() == tk.idx[a:b]).all()
            assert np.abs(w.cpu().numpy() - tk.weight[a:b]).max() < 1e-6
            assert onp.rel_l2(mine.cpu().numpy(), ro[:, :, q0:q1]) < 1e-5
            plain = bank.read(qk.to(dev), 50)     # (M = 1: the ordinary fused read; M > 1: the same exchange again)
            assert onp.rel_l2(plain.cpu().numpy(), ro[:, :, q0:q1]) < 1e-5
        bank.close()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,memory_shards", [(4, 2), (2, 1)])
def test_hybrid_query_groups_and_memory_shards_on_one_gpu(world, memory_shards):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, world, port, ret, memory_shards)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))
