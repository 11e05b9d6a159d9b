"""GPU parity of the fused memory read (through the C ABI) against golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from oracle import memread_np as onp
from tests.helpers import TIE_TOL, load, synth

pytestmark = pytest.mark.gpu

SMALL = ["small_a", "small_b", "small_exact_k", "small_k8", "small_ck32"]


def _paths_for(ck):
    from evavos_b200 import _lib
    return [_lib.PATH_SIMT] + ([_lib.PATH_TENSOR] if ck == 64 else [])


def _run(mk, qk, mv, top_k, path):
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    out, aff = ev.memory_read(bank, qk.to(dev), top_k, want_topk=True, path=path)
    torch.cuda.synchronize()
    return out.cpu().numpy(), aff.idx.cpu().numpy(), aff.weight.cpu().numpy(), aff.score.cpu().numpy()


def _check(mk, qk, mv, top_k, out, idx, w, sc, ref_readout=None, tag=""):
    ck = mk.shape[1]
    s64 = onp.affinity_scores(mk[0].reshape(ck, -1).numpy(), qk[0].reshape(ck, -1).numpy())
    exact, tie, bad, bad_q = onp.compare_topk(idx, s64, top_k, TIE_TOL)
    assert bad == 0, f"{tag}: {bad} queries with a wrong top-k set, e.g. {bad_q[:5]}"
    # weights: softmax over the selected scores (fp64) within fp32 rounding
    sel = np.take_along_axis(s64.T, idx.astype(np.int64), axis=1)
    e = np.exp(sel - sel.max(1, keepdims=True))
    w64 = e / e.sum(1, keepdims=True)
    assert np.abs(w - w64).max() < 2e-5, f"{tag}: weight error {np.abs(w - w64).max()}"
    assert np.abs(w.sum(1) - 1).max() < 1e-5
    # scores are the reference's affinity values
    assert np.abs(sc - sel).max() < 1e-4 * max(1.0, np.abs(sel).max())
    # best-first order (prop_net.py:53 relies on sorted topk)
    assert (np.diff(sc, axis=1) <= 1e-6).all()
    # readout: 1e-3 relative in fp32 (BASELINE.json north_star); we hold 1e-5 when no tie flipped
    ro = onp.readout(idx.astype(np.int64), w64, mv.reshape(mv.shape[0], mv.shape[1], -1).numpy())
    err_self = onp.rel_l2(out.reshape(ro.shape), ro)
    assert err_self < 1e-5, f"{tag}: readout vs own selection {err_self}"
    if ref_readout is not None:
        err = onp.rel_l2(out.reshape(ref_readout.shape), ref_readout)
        assert err < 1e-3, f"{tag}: readout vs reference {err}"
        if tie == 0:
            assert err < 1e-5, f"{tag}: readout vs reference {err} with identical top-k sets"
    return exact, tie


@pytest.mark.parametrize("name", SMALL)
def test_small_golden(name):
    g = load(f"memread_{name}.npz")
    mk, qk, mv = (torch.from_numpy(g[k]) for k in ("mk", "qk", "mv"))
    top_k = int(g["top_k"])
    for path in _paths_for(mk.shape[1]):
        out, idx, w, sc = _run(mk, qk, mv, top_k, path)
        exact, tie = _check(mk, qk, mv, top_k, out, idx, w, sc, g["readout"], f"{name}/path{path}")
        # the reference's own index sets (from its dense output) must match set-wise when nothing is tied
        if tie == 0:
            assert (np.sort(idx, 1) == np.sort(g["idx"], 1)).all()
            # dense form reproduces get_affinity's return value
            order = np.argsort(idx, 1)
            ref_order = np.argsort(g["idx"], 1)
            assert np.abs(np.take_along_axis(w, order, 1) - np.take_along_axis(g["weight"], ref_order, 1)).max() < 2e-6


def test_ties_golden():
    g = load("memread_ties.npz")
    mk, qk, mv = (torch.from_numpy(g[k]) for k in ("mk", "qk", "mv"))
    for path in _paths_for(64):
        out, idx, w, sc = _run(mk, qk, mv, 50, path)
        _check(mk, qk, mv, 50, out, idx, w, sc, None, f"ties/path{path}")
        # duplicated keys carry different values, so the readout depends on which twin is taken;
        # the weight mass per distinct score must still match the reference
        assert np.abs(np.sort(w, 1) - np.sort(g["weight"], 1)).max() < 2e-6


def test_too_short_raises():
    import evavos_b200 as ev
    mk, qk, mv = synth(22, 64, 8, 1, 6, 8, 1)
    dev = torch.device("cuda:0")
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    with pytest.raises(ev.EvavosError, match="out of range"):
        ev.memory_read(bank, qk.to(dev), 50)


def test_cfg1_golden():
    g = load("memread_cfg1.npz")
    ck, cv, t, h, w_, k = (int(x) for x in g["shape"])
    mk, qk, mv = synth(int(g["seed"]), ck, cv, t, h, w_, k)
    for path in _paths_for(64):
        out, idx, w, sc = _run(mk, qk, mv, 50, path)
        exact, tie = _check(mk, qk, mv, 50, out, idx, w, sc, None, f"cfg1/path{path}")
        assert (np.sort(idx, 1) == np.sort(g["idx"], 1)).mean() > 0.9999
        ref = g["readout"]
        mine = out[:, ::16]
        assert onp.rel_l2(mine, ref) < 1e-3
        assert np.abs(out.astype(np.float64).sum((2, 3)) - g["readout_sum"]).max() < 1e-2


def test_strided_bank_golden():
    """Bank pre-allocated for 7 frames, read the first 4 (inference_core.py:150-168)."""
    import evavos_b200 as ev
    g = load("memread_strided_bank.npz")
    dev = torch.device("cuda:0")
    keys, vals, qk = (torch.from_numpy(g[k]).to(dev) for k in ("keys", "values", "qk"))
    m = int(g["m_front"])
    reader = ev.EvalMemoryReader(50, None)
    aff = reader.get_affinity(keys[:, :, :m], qk)
    outs = torch.cat([reader.readout(aff, vals[i:i + 1, :, :m]) for i in range(vals.shape[0])], 0)
    fused = reader.read(keys[:, :, :m], qk, vals[:, :, :m])
    torch.cuda.synchronize()
    assert onp.rel_l2(outs.cpu().numpy(), g["readout"]) < 1e-5
    assert onp.rel_l2(fused.cpu().numpy(), g["readout"]) < 1e-5
    dense = aff.to_dense().cpu().numpy()[0]
    assert dense.shape == (m * 54, 54)
    assert np.abs(dense.sum(0) - 1).max() < 1e-5
    assert (np.count_nonzero(dense, 0) == 50).all()
    # append path: bank built frame by frame equals the import of the slice
    bank = ev.MemoryBank(vals.shape[0], 64, vals.shape[1], 6, 9, 7, dev)
    for f in range(m):
        bank.append(keys[:, :, f], vals[:, :, f:f + 1])
    out2, _ = ev.memory_read(bank, qk, 50)
    assert onp.rel_l2(out2.cpu().numpy(), g["readout"]) < 1e-5
    assert torch.equal(bank.keys_view(), keys[:, :, :m])
    assert torch.equal(bank.values_view(), vals[:, :, :m])


def test_massive_ties_overflow_path():
    """Every key identical: all scores tie, the tcgen05 filter overflows and the exact path takes over."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    one = torch.randn(1, 64, 1, 1, 1, generator=g)
    mk = one.expand(1, 64, 3, 10, 12).contiguous()
    mk[:, :, 1, 3, 4] += 0.5          # one distinct key
    qk = torch.randn(1, 64, 10, 12, generator=g)
    mv = torch.randn(1, 32, 3, 10, 12, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    res = {}
    for path in (_lib.PATH_SIMT, _lib.PATH_TENSOR):
        out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=path)
        res[path] = (out.cpu().numpy(), aff.idx.cpu().numpy(), aff.weight.cpu().numpy())
        idx = res[path][1]
        assert all(len(set(r.tolist())) == 50 for r in idx)
    # deterministic tie rule (lowest positions first) makes both paths agree exactly
    assert (res[_lib.PATH_SIMT][1] == res[_lib.PATH_TENSOR][1]).all()
    assert np.abs(res[_lib.PATH_SIMT][0] - res[_lib.PATH_TENSOR][0]).max() < 1e-6


def test_multi_frame_query_batch():
    """mem_freq query frames read in one launch equal per-frame reads."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    mk, qk, mv = synth(77, 64, 64, 4, 8, 11, 2)
    g = torch.Generator().manual_seed(78)
    qk3 = torch.randn(1, 64, 3, 8, 11, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    batched, _ = ev.memory_read(bank, qk3.to(dev), 50)
    for f in range(3):
        single, _ = ev.memory_read(bank, qk3[:, :, f].to(dev), 50)
        assert torch.equal(batched[:, :, f], single)


@pytest.mark.parametrize("cv", [512, 256, 96])
def test_bf16_value_bank(cv):
    """bf16 mode (BASELINE.json configs[4]): both sides get bf16-representable inputs, the reference evaluates them
    in fp32; the bf16 value shadow + fp32 accumulation must stay within 1e-2 relative (in fact ~1e-6: the rounding
    is lossless on representable inputs)."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    mk, qk, mv = synth(41, 64, cv, 4, 9, 12, 2)
    mk, qk, mv = (t.to(torch.bfloat16).float() for t in (mk, qk, mv))
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev), value_dtype=torch.bfloat16)
    assert bank.val_pm.dtype == torch.bfloat16
    out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True)
    tk, ro = onp.memory_read(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy(), mv.reshape(2, cv, -1).numpy(), 50)
    exact, tie, bad, _ = onp.compare_topk(aff.idx.cpu().numpy(), onp.affinity_scores(mk[0].reshape(64, -1).numpy(),
                                                                                     qk[0].reshape(64, -1).numpy()), 50, TIE_TOL)
    assert bad == 0
    err = onp.rel_l2(out.cpu().numpy().reshape(ro.shape), ro)
    assert err < 1e-2, err
    if tie == 0:
        assert err < 1e-5, err
    # arbitrary fp32 values stored in a bf16 shadow: rounding error of the values only, well inside 1e-2
    mk2, qk2, mv2 = synth(42, 64, cv, 4, 9, 12, 2)
    bank2 = ev.MemoryBank.from_tensors(mk2.to(dev), mv2.to(dev), value_dtype=torch.bfloat16)
    out2, _ = ev.memory_read(bank2, qk2.to(dev), 50)
    _, ro2 = onp.memory_read(mk2[0].reshape(64, -1).numpy(), qk2[0].reshape(64, -1).numpy(), mv2.reshape(2, cv, -1).numpy(), 50)
    assert onp.rel_l2(out2.cpu().numpy().reshape(ro2.shape), ro2) < 1e-2


@pytest.mark.parametrize("n_extra", [0, 1, 77, 78, 79])
def test_small_banks_around_tile_edges(n_extra):
    """N = 50 (== top_k), 51, 127, 128, 129: single / partial / two key tiles through the tcgen05 path."""
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    n = 50 + n_extra
    g = torch.Generator().manual_seed(100 + n)
    mk = torch.randn(1, 64, 1, 1, n, generator=g)          # one "frame" of n positions
    qk = torch.randn(1, 64, 3, 1, 45, generator=g)          # 3 query frames of 45 positions
    mv = torch.randn(2, 40, 1, 1, n, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    s64 = onp.affinity_scores(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy())
    tk = onp.topk_softmax(s64, 50)
    ro = onp.readout(tk.idx, tk.weight, mv.reshape(2, 40, -1).numpy())
    for path in (_lib.PATH_TENSOR, _lib.PATH_SIMT):
        out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True, path=path)
        exact, tie, bad, _ = onp.compare_topk(aff.idx.cpu().numpy(), s64, 50, TIE_TOL)
        assert bad == 0
        assert onp.rel_l2(out.cpu().numpy().reshape(ro.shape), ro) < (1e-5 if tie == 0 else 1e-3)


@pytest.mark.parametrize("top_k", [1, 7, 128])
def test_other_top_k(top_k):
    import evavos_b200 as ev
    from evavos_b200 import _lib
    dev = torch.device("cuda:0")
    mk, qk, mv = synth(300 + top_k, 64, 128, 3, 10, 13, 1)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    s64 = onp.affinity_scores(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy())
    tk = onp.topk_softmax(s64, top_k)
    ro = onp.readout(tk.idx, tk.weight, mv.reshape(1, 128, -1).numpy())
    for path in (_lib.PATH_TENSOR, _lib.PATH_SIMT):
        out, aff = ev.memory_read(bank, qk.to(dev), top_k, want_topk=True, path=path)
        exact, tie, bad, _ = onp.compare_topk(aff.idx.cpu().numpy(), s64, top_k, TIE_TOL)
        assert bad == 0
        assert onp.rel_l2(out.cpu().numpy().reshape(ro.shape), ro) < (1e-5 if tie == 0 else 1e-3)


def test_more_query_tiles_than_sms():
    """n_query > 128 * SM count: the cooperative filter runs in several waves of query tiles."""
    import evavos_b200 as ev
    dev = torch.device("cuda:0")
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    nq = 128 * n_sm + 200
    g = torch.Generator().manual_seed(9)
    mk = torch.randn(1, 64, 2, 10, 10, generator=g)
    mv = torch.randn(1, 16, 2, 10, 10, generator=g)
    qk = torch.randn(1, 64, 1, nq, generator=g)
    bank = ev.MemoryBank.from_tensors(mk.to(dev), mv.to(dev))
    out, aff = ev.memory_read(bank, qk.to(dev), 50, want_topk=True)
    s64 = onp.affinity_scores(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy())
    idx = aff.idx.cpu().numpy()
    rng = np.random.default_rng(1)
    sample = np.concatenate([rng.choice(nq, 300, replace=False), np.arange(nq - 200, nq)])
    exact, tie, bad, _ = onp.compare_topk(idx[sample], s64[:, sample], 50, TIE_TOL)
    assert bad == 0
    tk = onp.topk_softmax(s64[:, sample], 50)
    ro = onp.readout(tk.idx, tk.weight, mv.reshape(1, 16, -1).numpy())
    assert onp.rel_l2(out.cpu().numpy().reshape(1, 16, nq)[:, :, sample], ro) < 1e-3
