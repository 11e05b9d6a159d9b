"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

Usage (in the build container, where /root/reference exists):
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so
parity is pinned on outputs of the reference code itself, executed here on CPU
(fp32, torch CPU build).  The fixtures are small; large configurations are
regenerated from seeds at test time and checked through the oracle.
While generating, this script also asserts that oracle/torch_port.py reproduces the
reference bit-for-bit and that oracle/memread_np.py agrees within fp32 rounding.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("EVAVOS_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mivos.model.propagation.prop_net import EvalMemoryReader  # noqa: E402  (reference)
from mivos.model.aggregate import aggregate_wbg as ref_aggregate_wbg  # noqa: E402  (reference)
from mivos.tensor_util import pad_divide_by as ref_pad_divide_by  # noqa: E402  (reference)

from oracle import memread_np as onp  # noqa: E402
from oracle import torch_port as port  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth(seed, ck, cv, t, h, w, k, scale=1.0):
    """Seeded synthetic keys/values of SURVEY.md 8d: all ~N(0,1), fp32, CPU generator."""
    g = torch.Generator().manual_seed(seed)
    mk = torch.randn(1, ck, t, h, w, generator=g) * scale
    qk = torch.randn(1, ck, h, w, generator=g) * scale
    mv = torch.randn(k, cv, t, h, w, generator=g)
    return mk, qk, mv


def run_reference(mk, qk, mv, top_k):
    reader = EvalMemoryReader(top_k, km=None)
    aff = reader.get_affinity(mk, qk)                       # (1, N, HW) dense
    out = torch.cat([reader.readout(aff, mv[i:i + 1]) for i in range(mv.shape[0])], 0)
    return aff, out


def sparse_from_dense(aff, top_k):
    """(idx, weight) best-first from the reference's dense scatter output."""
    a = aff[0].T.contiguous()                               # (HW, N)
    w, idx = torch.topk(a, k=top_k, dim=1)                  # weights are >0 on exactly top_k entries
    return idx.numpy().astype(np.int32), w.numpy()


def check_port_and_oracle(mk, qk, mv, top_k, aff, out, tag):
    aff_p = port.dense_topk_affinity(mk.clone(), qk, top_k)
    assert torch.equal(aff_p, aff), f"{tag}: torch_port affinity differs from reference"
    out_p = port.memory_read(mk, qk, mv, top_k)
    assert torch.equal(out_p, out), f"{tag}: torch_port readout differs from reference"
    ck = mk.shape[1]
    s64 = onp.affinity_scores(mk[0].reshape(ck, -1).numpy(), qk[0].reshape(ck, -1).numpy())
    idx, w = sparse_from_dense(aff, top_k)
    exact, tie, bad, _ = onp.compare_topk(idx, s64, top_k, tie_tol=2e-5)
    assert bad == 0, f"{tag}: reference top-k disagrees with fp64 oracle beyond ties ({bad})"
    tk = onp.topk_softmax(s64, top_k)
    ro = onp.readout(tk.idx, tk.weight, mv.reshape(mv.shape[0], mv.shape[1], -1).numpy())
    err = onp.rel_l2(out.reshape(ro.shape).numpy(), ro)
    print(f"  {tag}: port==ref bitwise; oracle sets exact={exact} tie={tie}; readout rel-L2 vs fp64 = {err:.2e}")
    assert err < 2e-5 or tie > 0, f"{tag}: oracle readout differs: {err}"
    return s64


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(min(8, os.cpu_count() or 1))

    # ---- 1. small dense cases: inputs + full dense affinity + readout committed ----
    small = [
        # tag, seed, ck, cv, t, h, w, K, top_k, scale
        ("small_a", 11, 64, 32, 3, 6, 9, 2, 50, 1.0),
        ("small_b", 12, 64, 512, 2, 5, 10, 1, 50, 1.0),     # N = 100
        ("small_exact_k", 13, 64, 16, 1, 5, 10, 3, 50, 1.0),  # N == top_k
        ("small_k8", 14, 64, 24, 4, 4, 7, 2, 8, 0.5),
        ("small_ck32", 15, 32, 16, 3, 6, 6, 1, 20, 1.0),
    ]
    for tag, seed, ck, cv, t, h, w, k, top_k, scale in small:
        mk, qk, mv = synth(seed, ck, cv, t, h, w, k, scale)
        aff, out = run_reference(mk.clone(), qk, mv, top_k)
        check_port_and_oracle(mk, qk, mv, top_k, aff, out, tag)
        idx, wt = sparse_from_dense(aff, top_k)
        np.savez_compressed(os.path.join(OUT, f"memread_{tag}.npz"),
                            mk=mk.numpy(), qk=qk.numpy(), mv=mv.numpy(), top_k=top_k,
                            affinity=aff.numpy(), idx=idx, weight=wt, readout=out.numpy())

    # ---- 2. duplicated keys: exact ties in the affinity ----
    mk, qk, mv = synth(21, 64, 16, 2, 6, 8, 1, 1.0)
    mkf = mk.view(1, 64, -1)
    mkf[:, :, 48:96] = mkf[:, :, 0:48]          # second frame == first frame -> every score appears twice
    aff, out = run_reference(mk.clone(), qk, mv, 50)
    idx, wt = sparse_from_dense(aff, 50)
    np.savez_compressed(os.path.join(OUT, "memread_ties.npz"), mk=mk.numpy(), qk=qk.numpy(), mv=mv.numpy(),
                        top_k=50, idx=idx, weight=wt, readout=out.numpy())
    print("  ties: written")

    # ---- 3. THW < top_k raises in the reference (prop_net.py:53) ----
    mk, qk, mv = synth(22, 64, 8, 1, 6, 8, 1, 1.0)            # N = 48 < 50
    try:
        run_reference(mk.clone(), qk, mv, 50)
        raised = ""
    except RuntimeError as e:  # noqa: PERF203
        raised = str(e)
    assert "out of range" in raised
    with open(os.path.join(OUT, "memread_too_short.txt"), "w") as f:
        f.write(raised + "\n")

    # ---- 4. cfg1 (BASELINE.json configs[0]): 30x54, T=5, K=1; inputs from the seed ----
    mk, qk, mv = synth(1234 + 1, 64, 512, 5, 30, 54, 1)
    aff, out = run_reference(mk.clone(), qk, mv, 50)
    s64 = check_port_and_oracle(mk, qk, mv, 50, aff, out, "cfg1")
    idx, wt = sparse_from_dense(aff, 50)
    np.savez_compressed(os.path.join(OUT, "memread_cfg1.npz"), seed=1235, shape=np.array([64, 512, 5, 30, 54, 1]),
                        top_k=50, idx=idx, weight=wt,
                        readout_ch=np.arange(0, 512, 16), readout=out.numpy()[:, ::16],
                        readout_sum=out.numpy().astype(np.float64).sum((2, 3)),
                        colsum=aff[0].sum(0).numpy())
    del aff, s64

    # ---- 5. strided bank views (inference_core.py:150-168): bank larger than m_front ----
    g = torch.Generator().manual_seed(31)
    keys = torch.randn(1, 64, 7, 6, 9, generator=g)
    vals = torch.randn(2, 48, 7, 6, 9, generator=g)
    qk = torch.randn(1, 64, 6, 9, generator=g)
    m_front = 4
    aff, out = run_reference(keys[:, :, :m_front], qk, vals[:, :, :m_front], 50)
    idx, wt = sparse_from_dense(aff, 50)
    np.savez_compressed(os.path.join(OUT, "memread_strided_bank.npz"), keys=keys.numpy(), values=vals.numpy(),
                        qk=qk.numpy(), m_front=m_front, top_k=50, idx=idx, weight=wt, readout=out.numpy())
    print("  strided bank: written")

    # ---- 6. aggregate_wbg ----
    g = torch.Generator().manual_seed(4321)
    agg = {}
    for name, k, h, w in (("k1", 1, 32, 48), ("k3", 3, 32, 48), ("k5_odd", 5, 17, 23)):
        p = torch.rand(k, 1, h, w, generator=g)
        p.view(-1)[:6] = torch.tensor([0.0, 1.0, 1e-8, 1 - 1e-8, 0.5, 1e-7])  # clamp edges
        agg[f"{name}_prob"] = p.numpy()
        for keep_bg in (False, True):
            for hard in (False, True):
                r = ref_aggregate_wbg(p.clone(), keep_bg=keep_bg, hard=hard)
                rp = port.aggregate_wbg(p.clone(), keep_bg=keep_bg, hard=hard)
                assert torch.equal(r, rp)
                o = onp.aggregate_wbg(p.numpy(), keep_bg=keep_bg, hard=hard)
                if not hard:
                    assert np.abs(o - r.numpy()).max() < 2e-6, np.abs(o - r.numpy()).max()
                agg[f"{name}_bg{int(keep_bg)}_hard{int(hard)}"] = r.numpy()
    np.savez_compressed(os.path.join(OUT, "aggregate_wbg.npz"), **agg)
    print("  aggregate: written")

    # ---- 7. pad_divide_by ----
    pads = {}
    for h, w in ((480, 854), (480, 864), (1080, 1920), (17, 33), (16, 16), (96, 128)):
        x = torch.zeros(1, 1, h, w)
        y, pad = ref_pad_divide_by(x, 16)
        assert tuple(pad) == onp.pad_amounts(h, w, 16)
        pads[f"{h}x{w}"] = np.array(list(pad) + list(y.shape[-2:]))
    np.savez_compressed(os.path.join(OUT, "pad_divide_by.npz"), **pads)
    print("  pad: written")
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
