"""CPU: the oracle (numpy fp64 checker and torch fp32 port) against the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import memread_np as onp
from oracle import torch_port as port
from tests.helpers import TIE_TOL, load, synth

SMALL = ["small_a", "small_b", "small_exact_k", "small_k8", "small_ck32"]


@pytest.mark.parametrize("name", SMALL)
def test_port_bit_exact_and_oracle_close(name):
    g = load(f"memread_{name}.npz")
    mk, qk, mv = (torch.from_numpy(g[k]) for k in ("mk", "qk", "mv"))
    top_k = int(g["top_k"])
    aff = port.dense_topk_affinity(mk.clone(), qk, top_k)
    assert np.array_equal(aff.numpy(), g["affinity"])
    out = port.memory_read(mk, qk, mv, top_k)
    assert np.array_equal(out.numpy(), g["readout"])
    ck = mk.shape[1]
    tk, ro = onp.memory_read(mk[0].reshape(ck, -1).numpy(), qk[0].reshape(ck, -1).numpy(),
                             mv.reshape(mv.shape[0], mv.shape[1], -1).numpy(), top_k)
    assert (np.sort(tk.idx, 1) == np.sort(g["idx"].astype(np.int64), 1)).all()
    assert onp.rel_l2(ro.reshape(g["readout"].shape), g["readout"]) < 2e-6
    dense = onp.dense_affinity(tk.idx, tk.weight, mk[0].reshape(ck, -1).shape[1])
    assert np.abs(dense - g["affinity"][0]).max() < 2e-6


def test_oracle_cfg1():
    g = load("memread_cfg1.npz")
    ck, cv, t, h, w, k = (int(x) for x in g["shape"])
    mk, qk, mv = synth(int(g["seed"]), ck, cv, t, h, w, k)
    s64 = onp.affinity_scores(mk[0].reshape(ck, -1).numpy(), qk[0].reshape(ck, -1).numpy())
    exact, tie, bad, _ = onp.compare_topk(g["idx"], s64, 50, TIE_TOL)
    assert bad == 0 and exact + tie == h * w
    tk = onp.topk_softmax(s64, 50)
    ro = onp.readout(tk.idx, tk.weight, mv.reshape(k, cv, -1).numpy())
    assert onp.rel_l2(ro[:, ::16].reshape(g["readout"].shape), g["readout"]) < 2e-6
    assert np.abs(g["colsum"] - 1).max() < 1e-5


def test_oracle_ties_and_range():
    g = load("memread_ties.npz")
    mk, qk = torch.from_numpy(g["mk"]), torch.from_numpy(g["qk"])
    s64 = onp.affinity_scores(mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy())
    exact, tie, bad, _ = onp.compare_topk(g["idx"], s64, 50, TIE_TOL)
    assert bad == 0
    with pytest.raises(RuntimeError, match="out of range"):
        onp.topk_softmax(np.zeros((48, 4)), 50)
    with open(__import__("os").path.join(__import__("tests.helpers", fromlist=["GOLDEN"]).GOLDEN,
                                         "memread_too_short.txt")) as f:
        assert "out of range" in f.read()


def test_oracle_strided_bank():
    g = load("memread_strided_bank.npz")
    m = int(g["m_front"])
    keys, vals, qk = g["keys"], g["values"], g["qk"]
    tk, ro = onp.memory_read(keys[0, :, :m].reshape(64, -1), qk[0].reshape(64, -1),
                             vals[:, :, :m].reshape(vals.shape[0], vals.shape[1], -1), 50)
    assert onp.rel_l2(ro.reshape(g["readout"].shape), g["readout"]) < 2e-6
    # bank append restatement writes the same slots
    k2 = np.zeros_like(keys)
    v2 = np.zeros_like(vals)
    for f in range(keys.shape[2]):
        onp.bank_append(k2, v2, f, keys[:, :, f], vals[:, :, f])
    assert np.array_equal(k2, keys) and np.array_equal(v2, vals)


def test_oracle_aggregate():
    g = load("aggregate_wbg.npz")
    for name in ("k1", "k3", "k5_odd"):
        p = g[f"{name}_prob"]
        for keep_bg in (False, True):
            ref = g[f"{name}_bg{int(keep_bg)}_hard0"]
            assert np.abs(onp.aggregate_wbg(p, keep_bg=keep_bg) - ref).max() < 2e-6
            t = port.aggregate_wbg(torch.from_numpy(p), keep_bg=keep_bg).numpy()
            assert np.array_equal(t, ref)
            th = port.aggregate_wbg(torch.from_numpy(p), keep_bg=keep_bg, hard=True).numpy()
            assert np.array_equal(th, g[f"{name}_bg{int(keep_bg)}_hard1"])


def test_oracle_pad():
    g = load("pad_divide_by.npz")
    for key in g.files:
        h, w = (int(x) for x in key.split("x"))
        lw, uw, lh, uh = onp.pad_amounts(h, w)
        assert [lw, uw, lh, uh, h + lh + uh, w + lw + uw] == g[key].tolist()


def test_oracle_sharded_equals_single():
    mk, qk, mv = synth(91, 64, 16, 6, 5, 7, 2)
    mkf, qkf, mvf = mk[0].reshape(64, -1).numpy(), qk[0].reshape(64, -1).numpy(), mv.reshape(2, 16, -1).numpy()
    tk, ro = onp.memory_read(mkf, qkf, mvf, 50)
    hw = 35
    for shards in (2, 3):
        owners = [np.concatenate([np.arange(f * hw, (f + 1) * hw) for f in range(r, 6, shards)]) for r in range(shards)]
        tk2, ro2 = onp.sharded_memory_read(mkf, qkf, mvf, 50, owners)
        assert (tk2.idx == tk.idx).all()
        assert np.abs(ro2 - ro).max() < 1e-12


@pytest.mark.parametrize("name", ["b2", "b4_peaky", "b1_flat"])
def test_oracle_attention(name):
    """Fusion-path attention (prop_net.py:117-138, 198-211): torch port bit-exact, fp64 oracle within fp32 rounding."""
    g = load("attention.npz")
    mk, qk = torch.from_numpy(g[f"{name}_mk"]), torch.from_numpy(g[f"{name}_qk"])
    pos, neg = torch.from_numpy(g[f"{name}_pos"]), torch.from_numpy(g[f"{name}_neg"])
    assert torch.equal(port.get_attention(mk, pos, neg, qk), torch.from_numpy(g[f"{name}_attn"]))
    low = port.attention_lowres(mk, pos, neg, qk)
    assert torch.equal(low, torch.from_numpy(g[f"{name}_lowres"]))
    b, _, nh, nw = low.shape
    vec = torch.stack([torch.nn.functional.interpolate(pos, size=(nh, nw), mode="area").view(b, -1),
                       torch.nn.functional.interpolate(neg, size=(nh, nw), mode="area").view(b, -1)], 1)
    o64 = onp.attention_readout(mk.reshape(64, -1).numpy(), qk.reshape(64, -1).numpy(), vec.reshape(2 * b, -1).numpy())
    assert np.abs(o64.reshape(b, 2, nh, nw) - low.numpy()).max() < 2e-6


def test_jf_restatement_reproduces_reference_metrics():
    """oracle/jf_np.py against tests/golden/jf.npz (written by the reference's own interactions/metrics.py)."""
    from oracle import jf_np
    g = load("jf.npz")
    for name in ("small", "odd"):
        shape = tuple(int(x) for x in g[f"{name}_shape"])
        pred = np.unpackbits(g[f"{name}_pred"], axis=-1)[..., :shape[2]].astype(bool)
        gt = np.unpackbits(g[f"{name}_gt"], axis=-1)[..., :shape[2]].astype(bool)
        for f in range(shape[0]):
            assert jf_np.compute_iou(pred[f], gt[f]) == g[f"{name}_j"][f]
            if gt[f].any():
                assert jf_np.j_and_f(pred[f], gt[f]) == g[f"{name}_jf"][f]
                assert jf_np.f_measure(pred[f], gt[f]) == g[f"{name}_f"][f]
            else:
                assert g[f"{name}_jf"][f] == 20


@pytest.mark.parametrize("tag", ["e2e_k1", "e2e_k2"])
def test_stock_engine_matches_reference(tag):
    """oracle/stock_engine.py (bench.py's cfg3 `gpu_baseline`) replays the interactions the live reference InferenceCore
    recorded into tests/golden/e2e_*.npz - both on CPU, same op sequence: probabilities to 1e-5, masks equal except
    where the reference itself is undecided."""
    import evavos_b200 as ev
    from evavos_b200.networks import seeded_init
    from oracle.stock_engine import StockEngine
    g = load(f"{tag}.npz")
    with torch.no_grad():
        prop, fuse = ev.PropagationNetwork().eval(), ev.FusionNet().eval()
        seeded_init(prop, 1001)
        seeded_init(fuse, 1002)
        eng = StockEngine(prop, fuse, torch.from_numpy(g["images"]), int(g["num_objects"]), mem_freq=int(g["mem_freq"]),
                          device="cpu")
        assert tuple(eng.pad) == tuple(int(x) for x in g["pad"])
        for n in range(int(g["n_interactions"])):
            out = eng.interact(torch.from_numpy(g[f"mask_{n}"]), int(g[f"frame_{n}"]), scribble=bool(g[f"scribble_{n}"]))
            ref_prob, ref_masks = g[f"prob_{n}"], g[f"np_masks_{n}"]
            assert np.abs(eng.prob.numpy() - ref_prob).max() < 1e-5
            srt = np.sort(ref_prob, 0)
            lw, uw, lh, uh = (int(x) for x in g["pad"])
            undecided = (srt[-1] - srt[-2] < 1e-5)[:, 0]
            undecided = undecided[:, lh:undecided.shape[1] - uh or None, lw:undecided.shape[2] - uw or None]
            assert out.shape == ref_masks.shape and not ((out != ref_masks) & ~undecided).any()
