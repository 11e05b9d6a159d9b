"""GPU parity added in round 2: full-size cfg5 in bf16, the host-buffer entry
point, in-place rewrites of a bank frame, the sampled threshold pass (sample_stride), the finalizer's exact path,
readouts into a strided destination, and the shadow cache of EvalMemoryReader.  Everything goes through the C ABI;
the oracle (oracle/memread_np.py, fp64) is the checker.
"""
import numpy as np
import pytest
import torch

from oracle import memread_np as onp
from tests.helpers import TIE_TOL, load, synth
from tests.test_gpu_memread import _check

pytestmark = pytest.mark.gpu


def _sample_queries(hw, n, seed=0):
    """n random queries plus the whole last (partial) 128-row query tile and the first row."""
    rng = np.random.default_rng(seed)
    last_tile = np.arange((hw - 1) // 128 * 128, hw)
    pick = np.concatenate([rng.choice(hw, n, replace=False), last_tile[:: max(1, len(last_tile) // 24)], [0, hw - 1]])
    return np.unique(pick)


def test_cfg5_full_size_bf16():
    """BASELINE.json configs[4]: 68x120 feature map, 5 objects, 50-frame bank (N = 408 000, 8 160 queries, 64 query
    tiles x 2 memory chunks), bf16 value shadow.  Inputs are bf16-representable on both sides (the reference would
    evaluate them in fp32).  >= 256 sampled queries (incl. the last partial query tile) against the fp64 oracle over
    ALL positions; every query against the exact CUDA-core selection; sampled readouts against a weighted gather."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    ck, cv, t, h, w, k = 64, 512, 50, 68, 120, 5
    hw, n = h * w, t * h * w
    g = torch.Generator().manual_seed(1239)
    mk = torch.randn(1, ck, t, h, w, generator=g).to(torch.bfloat16).float()
    qk = torch.randn(1, ck, h, w, generator=g).to(torch.bfloat16).float()
    gd = torch.Generator(device=dev).manual_seed(1239)
    bank = ev.MemoryBank(k, ck, cv, h, w, t, dev, value_dtype=torch.bfloat16, keep_reference_layout=False)
    for f in range(t):   # 4.2 GB of fp32 values: synthesised on the device, frame by frame
        vf = torch.randn(k, cv, 1, h, w, generator=gd, device=dev).to(torch.bfloat16).float()
        bank.append(mk[:, :, f].to(dev), vf)
    out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_TENSOR)
    _, aff_x = ev.memory_read(bank, qk.to(dev), 50, want_readout=False, want_topk=True, path=_lib.PATH_SIMT)
    torch.cuda.synchronize()
    assert torch.equal(aff.idx, aff_x.idx), "tcgen05 filter and exact selection disagree"
    assert torch.equal(aff.weight, aff_x.weight)
    idx, wgt, sc = aff.idx.cpu().numpy(), aff.weight.cpu().numpy(), aff.score.cpu().numpy()
    assert idx.min() >= 0 and idx.max() < n
    assert (np.sort(idx, 1)[:, 1:] != np.sort(idx, 1)[:, :-1]).all(), "duplicate positions"
    assert (np.diff(sc, axis=1) <= 0).all()
    assert np.abs(wgt.sum(1) - 1).max() < 1e-5
    sample = _sample_queries(hw, 256)
    assert len(sample) >= 256
    s64 = onp.affinity_scores(mk[0].reshape(ck, n).numpy(), qk[0].reshape(ck, hw).numpy()[:, sample])
    exact, tie, bad, bad_q = onp.compare_topk(idx[sample], s64, 50, TIE_TOL)
    assert bad == 0, bad_q
    assert np.abs(np.take_along_axis(s64.T, idx[sample].astype(np.int64), 1) - sc[sample]).max() < 1e-4
    # readout (north_star: 1e-2 relative in bf16): weighted gather of the selected rows of the bf16 shadow in fp64
    out = out.view(k, cv, hw)
    for q in sample[::32]:
        rows = bank.val_pm[:, torch.from_numpy(idx[q].astype(np.int64)).to(dev)].double()       # (k, 50, cv)
        ref = (rows * torch.from_numpy(wgt[q].astype(np.float64)).to(dev)[None, :, None]).sum(1)  # (k, cv)
        err = ((out[:, :, q].double() - ref).norm() / ref.norm()).item()
        assert err < 1e-2, err
        assert err < 1e-5, err   # representable inputs: only the fp32 accumulation order differs


def test_memread_host_matches_golden_and_oracle():
    """evavos_memread_host (HOST buffers in the reference layouts) - the one C-ABI entry that had no checker."""
    from evavos_b200 import _lib
    from evavos_b200.host_api import memory_read_host
    g = load("memread_small_a.npz")
    mk, qk, mv = (torch.from_numpy(g[k]) for k in ("mk", "qk", "mv"))
    top_k = int(g["top_k"])
    ck = mk.shape[1]
    s64 = onp.affinity_scores(mk[0].reshape(ck, -1).numpy(), qk[0].reshape(ck, -1).numpy())
    for path in (_lib.PATH_AUTO, _lib.PATH_SIMT):
        h2d, d2h, out, idx, wgt = memory_read_host(mk, qk, mv, top_k, path=path, want_topk=True)
        assert h2d == 4 * (mk.numel() + qk.numel() + mv.numel())
        assert d2h == 4 * out.numel() + 8 * idx.numel()
        exact, tie, bad, _ = onp.compare_topk(idx.numpy(), s64, top_k, TIE_TOL)
        assert bad == 0
        assert onp.rel_l2(out.numpy().reshape(g["readout"].shape), g["readout"]) < (1e-5 if tie == 0 else 1e-3)
        assert np.abs(wgt.numpy().sum(1) - 1).max() < 1e-5
    # pinned inputs and a caller-provided pinned output (what bench.py's stateless e2e leg does)
    out2 = torch.empty_like(out).pin_memory()
    memory_read_host(mk.pin_memory(), qk.pin_memory(), mv.pin_memory(), top_k, out=out2)
    assert torch.equal(out2, out)


def test_rewrite_of_a_middle_frame():
    """rewriting frame f of a bank whose frame size is not a multiple of the 128-row key tile
    must leave the first positions of frame f + 1 alone.  Tensor path == exact path == oracle on the edited bank."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    ck, cv, t, h, w, k = 64, 64, 5, 30, 54, 2          # 1620 positions per frame = 12.66 tiles
    mk, qk, mv = synth(31, ck, cv, t, h, w, k)
    bank = ev.MemoryBank(k, ck, cv, h, w, t, dev)
    bank.write_frames(0, mk.to(dev), mv.to(dev))
    g = torch.Generator().manual_seed(32)
    for f in (1, 3, 0):
        nk = torch.randn(1, ck, 1, h, w, generator=g) * 1.5      # strong keys: the new frame must show up in the top-k
        nv = torch.randn(k, cv, 1, h, w, generator=g)
        mk[:, :, f:f + 1], mv[:, :, f:f + 1] = nk, nv
        bank.write_frames(f, nk.to(dev), nv.to(dev))
        assert bank.n_frames == t
        assert torch.equal(bank.keys_view().cpu(), mk) and torch.equal(bank.values_view().cpu(), mv)
        res = {}
        for path in (_lib.PATH_TENSOR, _lib.PATH_SIMT):
            out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=path)
            # top-k sets vs the fp64 oracle (near-ties classified), weights, scores, readout of the selected rows
            _check(mk, qk, mv, 50, out.cpu().numpy(), aff.idx.cpu().numpy(), aff.weight.cpu().numpy(),
                   aff.score.cpu().numpy(), None, f"rewrite frame {f} path {path}")
            res[path] = aff.idx
        assert torch.equal(res[_lib.PATH_TENSOR], res[_lib.PATH_SIMT])


@pytest.mark.parametrize("t,frames", [(16, 1), (30, 2)])
def test_sample_stride_does_not_change_the_result(t, frames):
    """The threshold pass may look at every R-th key tile only (1.5 .. 1.25 sweeps instead of 2): a tuning knob, the
    selected set is the exact top-k for every R (and R = 1 is the round-1 two-sweep algorithm)."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    mk, _, mv = synth(50 + t, 64, 128, t, 30, 54, 1)
    qk = torch.randn(1, 64, frames, 30, 54, generator=torch.Generator().manual_seed(51))
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    ref_out, ref = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_SIMT)
    for r in (1, 2, 3, 4, 8):
        out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_TENSOR, sample_stride=r)
        assert torch.equal(aff.idx, ref.idx), r
        assert torch.equal(aff.weight, ref.weight), r
        assert torch.equal(out, ref_out), r


@pytest.mark.parametrize("n_copies", [3000, 600])
def test_big_tie_clusters(n_copies):
    """Many copies of one key (a static background seen in many memory frames).  3000: every copy is a candidate,
    the list overflows (> 1024 entries) and the finalizer's warp redoes the query exactly over all positions.
    600: the list holds them all but more than 256 survive the finalizer's cut - they are rescored in batches.
    Either way the lowest positions win the ties, deterministically."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(61)
    t, h, w = 4, 30, 54
    n = t * h * w
    mk = torch.randn(64, n, generator=g)
    hot = torch.randn(64, 1, generator=g) * 1.2
    copies = torch.randperm(n, generator=g)[:n_copies].sort().values
    mk[:, copies] = hot
    qk = torch.randn(1, 64, h, w, generator=g)
    qk[0, :, :4] = (hot * 1.1).view(64, 1, 1)        # 4 x 54 queries for which all copies tie at the top
    mv = torch.randn(1, 32, t, h, w, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.view(1, 64, t, h, w).to(dev), mv.to(dev))
    out_x, aff_x = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_SIMT)
    out_t, aff_t = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_TENSOR)
    assert torch.equal(aff_t.idx, aff_x.idx)
    assert torch.equal(out_t, out_x)
    assert torch.equal(aff_t.idx[:4 * 54].cpu(), copies[:50].to(torch.int32).expand(4 * 54, 50))
    s64 = onp.affinity_scores(mk.numpy(), qk[0].reshape(64, -1).numpy())
    exact, tie, bad, _ = onp.compare_topk(aff_t.idx.cpu().numpy(), s64, 50, TIE_TOL)
    assert bad == 0


def test_readout_into_a_strided_destination():
    """the readout lands directly in the first CV channels of the decoder's (K, 2*CV, H, W) input -
    no torch.cat (prop_net.py:189-190)."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    mk, qk, mv = synth(71, 64, 512, 3, 12, 17, 3)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    ref, _ = ev.memory_read(bank, qk.to(dev), 50)
    m4 = torch.full((3, 1024, 12, 17), -7.0, device=dev)
    got, _ = ev.memory_read(bank, qk.to(dev), 50, out=m4)
    assert got.data_ptr() == m4.data_ptr() and got.shape == (3, 512, 12, 17)
    assert torch.equal(m4[:, :512], ref)
    assert bool((m4[:, 512:] == -7.0).all())
    with pytest.raises(ValueError):
        ev.memory_read(bank, qk.to(dev), 50, out=torch.empty((3, 100, 12, 17), device=dev))


def test_reader_shadow_cache_follows_the_tensors():
    """ADVICE r1 (memory_reader.py): a reference-style caller hands growing T-slices of ONE allocation to every read
    (inference_core.py:150-177).  The shadow follows appended frames incrementally, never aliases a later
    allocation that happens to get the same address, and notices an in-place edit when append_only is off."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    ck, cv, t, h, w, k = 64, 64, 6, 9, 14, 2
    mk, qk, mv = synth(81, ck, cv, t, h, w, k)
    qk = qk.to(dev)

    def fresh(keys, vals):
        out, _ = ev.memory_read(ev.MemoryBank.from_tensors(keys, vals), qk, 50)
        return out

    reader = ev.EvalMemoryReader(50, None)
    for _pass in range(2):                       # two passes: the second allocation usually reuses the address
        keys = torch.empty((1, ck, t, h, w), device=dev)
        vals = torch.empty((k, cv, t, h, w), device=dev)
        src_k = mk.to(dev) * (1.0 + _pass)
        src_v = mv.to(dev) + _pass
        keys[:, :, :1], vals[:, :, :1] = src_k[:, :, :1], src_v[:, :, :1]
        for m in range(1, t + 1):
            if m > 1:
                keys[:, :, m - 1], vals[:, :, m - 1] = src_k[:, :, m - 1], src_v[:, :, m - 1]
            got = reader.read(keys[:, :, :m], qk, vals[:, :, :m])
            assert torch.equal(got, fresh(keys[:, :, :m].contiguous(), vals[:, :, :m].contiguous())), (_pass, m)
        assert torch.equal(reader.read(keys[:, :, :3], qk, vals[:, :, :3]),
                           fresh(keys[:, :, :3].contiguous(), vals[:, :, :3].contiguous()))
        del keys, vals
    strict = ev.EvalMemoryReader(50, None, append_only=False)
    keys, vals = mk.to(dev).clone(), mv.to(dev).clone()
    a = strict.read(keys, qk, vals)
    keys[:, :, 2] *= 3.0                         # in-place edit of an old frame
    b = strict.read(keys, qk, vals)
    assert torch.equal(b, fresh(keys, vals)) and not torch.equal(a, b)


def test_argmax_nan_and_many_objects_attention():
    """NaN probabilities follow torch.argmax; the attention read takes more than 32 mask rows."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    prob = torch.rand(3, 2, 1, 32, 48, device=dev)
    prob[1, 0, 0, 5, 7] = float("nan")
    prob[2, 1, 0, 9, 9] = float("nan")
    masks, unp = ev.argmax_unpad(prob, (0, 0, 0, 0), 32, 48)
    assert torch.equal(masks[:, 0].long(), prob[:, :, 0].argmax(0))
    g = torch.Generator().manual_seed(5)
    mk = torch.randn(1, 64, 6, 8, generator=g).to(dev)
    qk = torch.randn(1, 64, 6, 8, generator=g).to(dev)
    vec = torch.rand(40, 48, generator=g).to(dev)
    got = ev.attention_readout(mk, qk, vec)
    m, q = mk.reshape(64, -1).double(), qk.reshape(64, -1).double()
    s = (-(m * m).sum(0)[:, None] + 2 * m.t() @ q - (q * q).sum(0)[None]) / 8.0
    ref = vec.double() @ torch.softmax(s, 0)
    assert (got.double() - ref).abs().max().item() < 2e-6


@pytest.mark.parametrize("t,frames,noise,top_k", [(6, 1, 2e-3, 50), (3, 2, 1e-3, 50), (7, 3, 5e-4, 50), (1, 1, 1e-3, 50),
                                                  (3, 1, 1e-3, 100)])
def test_near_constant_keys_take_the_exact_tiled_pass(t, frames, noise, top_k):
    """Keys that differ by less than the filter's bf16 error margin (what networks with random weights produce for
    every frame of a video): every list overflows, the finalizer hands the queries to overflow_exact_kernel, which
    scores them exactly in 32 x 128 tiles.  Same arithmetic as the SIMT path - identical indices - and the oracle's
    top-k.  3 - 7 frames (>= 16 blocks of 128 positions): the threshold pass samples every 4th block; 1 frame (13
    blocks): every 2nd; top_k = 100 (> 64): all of them."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    from evavos_b200.memory_reader import last_overflow_count
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(97 + t)
    h, w = 30, 54
    base = torch.randn(64, 1, generator=g)
    mk = (base + noise * torch.randn(64, t * h * w - 37, generator=g))          # bank length not a multiple of 128
    mk = torch.cat([mk, base + noise * torch.randn(64, 37, generator=g)], 1).view(1, 64, t, h, w)
    shape = (1, 64, frames, h, w) if frames > 1 else (1, 64, h, w)
    qk = base.view(1, 64, *([1] * (len(shape) - 2))) * 0.9 + noise * torch.randn(shape, generator=g)
    mv = torch.randn(2, 32, t, h, w, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    out_t, aff_t = ev.memory_read(bank, qk.to(dev), top_k, want_topk=True, path=_lib.PATH_TENSOR_DENSE)
    n_over = last_overflow_count()
    out_x, aff_x = ev.memory_read(bank, qk.to(dev), top_k, want_topk=True, path=_lib.PATH_SIMT)
    assert last_overflow_count() == 0
    nq = frames * h * w
    assert n_over > 0.9 * nq, f"only {n_over} of {nq} queries overflowed: the test does not reach the tiled pass"
    assert torch.equal(aff_t.idx, aff_x.idx)
    assert torch.equal(aff_t.weight, aff_x.weight)
    assert torch.equal(out_t, out_x)
    # PATH_TENSOR launches the tiled pass only after a read has reported an overflow (the hint of api.cu); before
    # that a warp of the finalizer redoes each overflowed query - the result is the same either way
    for _ in range(2):
        out_h, aff_h = ev.memory_read(bank, qk.to(dev), top_k, want_topk=True, path=_lib.PATH_TENSOR)
        assert last_overflow_count() == n_over
        assert torch.equal(aff_h.idx, aff_x.idx) and torch.equal(out_h, out_x)
    s64 = onp.affinity_scores(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy())
    exact, tie, bad, bad_q = onp.compare_topk(aff_t.idx.cpu().numpy(), s64, top_k, TIE_TOL)
    assert bad == 0, bad_q[:5]


def test_overflow_count_is_zero_on_separable_scores():
    import evavos_b200 as ev
    from evavos_b200.memory_reader import last_overflow_count
    from tests.helpers import synth
    dev = torch.device("cuda:0")
    mk, qk, mv = synth(5, 64, 16, 4, 30, 54, 1)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    ev.memory_read(bank, qk.to(dev), 50)
    assert last_overflow_count() == 0


def test_more_query_tiles_than_sms():
    """20 800 queries = 163 query tiles on 148 SMs: the filter runs one full wave (148 tiles, one CTA per tile over
    the whole bank) and a partial wave whose 15 tiles split the bank into 9 chunks each - class maxima, thresholds
    and lists of both waves must give the exact selection."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(123)
    t, h, w, frames = 2, 40, 40, 13
    mk = torch.randn(1, 64, t, h, w, generator=g)
    mv = torch.randn(1, 32, t, h, w, generator=g)
    qk = torch.randn(1, 64, frames, h, w, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    out_t, aff_t = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_TENSOR)
    out_x, aff_x = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=_lib.PATH_SIMT)
    assert torch.equal(aff_t.idx, aff_x.idx)
    assert torch.equal(aff_t.weight, aff_x.weight)
    assert torch.equal(out_t, out_x)
    # a sample of queries from both waves against the fp64 oracle
    nq = frames * h * w
    pick = np.r_[0:64, 148 * 128 - 32:148 * 128 + 96, nq - 64:nq]
    s64 = onp.affinity_scores(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1)[:, pick].numpy())
    exact, tie, bad, bad_q = onp.compare_topk(aff_t.idx.cpu().numpy()[pick], s64, 50, TIE_TOL)
    assert bad == 0, bad_q[:5]


@pytest.mark.parametrize("bf16", [False, True])
def test_readout_into_a_channels_last_destination(bf16):
    """The decoder input of an NHWC engine: (F, K, 2*CV, H, W) / (K, 2*CV, H, W) blocks in channels_last memory
    format receive the readout channel-contiguously (readout_ch_stride = 1) - the same numbers as the plain layout."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    K, CV, T, H, W, F = 2, 512, 3, 9, 14, 3
    mk, _, mv = synth(77, 64, CV, T, H, W, K)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev), value_dtype=torch.bfloat16 if bf16 else torch.float32)
    g = torch.Generator().manual_seed(78)
    qk5 = torch.randn(1, 64, F, H, W, generator=g).to(dev)
    ref5, _ = ev.memory_read(bank, qk5, 50)                                   # (K, CV, F, H, W)
    m4 = torch.full((F * K, 2 * CV, H, W), -7.0, device=dev).contiguous(memory_format=torch.channels_last).view(F, K, 2 * CV, H, W)
    got5, _ = ev.memory_read(bank, qk5, 50, out=m4)
    assert got5.data_ptr() == m4.data_ptr() and m4.stride(2) == 1
    assert torch.equal(m4[:, :, :CV], ref5.permute(2, 0, 1, 3, 4))
    assert (m4[:, :, CV:] == -7.0).all()                                      # the other half is not touched
    qk4 = qk5[:, :, 1].contiguous()
    ref4, _ = ev.memory_read(bank, qk4, 50)
    o4 = torch.full((K, 2 * CV, H, W), -7.0, device=dev).contiguous(memory_format=torch.channels_last)
    ev.memory_read(bank, qk4, 50, out=o4)
    assert torch.equal(o4[:, :CV], ref4) and (o4[:, CV:] == -7.0).all()
