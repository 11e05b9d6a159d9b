"""Host model of the tensor-core form of the attention read (csrc/attention.cu): the fp16 hi / lo split with
power-of-two normalisation, and the channel -> MMA k-slot assignment shared by the A (query) and B (key) fragments.
The kernel itself is checked on the GPU (tests/test_gpu_attention.py); this pins its arithmetic claims on the CPU."""
import numpy as np
import pytest


def range_factors(mx):
    """csrc/attention.cu:range_factors - exact powers of two that move max |x| into [2^14, 2^15) and back."""
    bits = np.float32(mx).view(np.uint32)
    e = int((bits >> 23) & 0xFF) - 127 - 14
    e = max(e, -100)
    return np.float32(2.0) ** np.float32(-e), np.float32(2.0) ** np.float32(e)


def split(x):
    """split_pair: hi = fp16(x), lo = fp16(x - hi), the subtraction in fp32."""
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def split_dot(q, k):
    """a.b ~ a_hi.b_hi + (a_lo.b_hi + a_hi.b_lo): fp16 operands, exact products, fp32-or-better accumulation."""
    dq, uq = range_factors(np.abs(q).max())
    dk, uk = range_factors(np.abs(k).max())
    qh, ql = split(q * dq)
    kh, kl = split(k * dk)
    f = lambda a: a.astype(np.float64)
    main = (f(qh) * f(kh)).sum()
    corr = (f(ql) * f(kh)).sum() + (f(qh) * f(kl)).sum()
    return (main + corr) * float(uq) * float(uk)


@pytest.mark.parametrize("scale_q,scale_k", [(1.0, 1.0), (3.0, 0.2), (2.0e4, 2.0e4), (1.0e-3, 1.0e-3), (1.0e-30, 5.0e3),
                                             (6.0e4, 1.0e-6)])
def test_split_dot_is_fp32_grade(scale_q, scale_k):
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(200):
        q = (rng.standard_normal(64) * scale_q).astype(np.float32)
        k = (rng.standard_normal(64) * scale_k).astype(np.float32)
        if rng.random() < 0.3:                      # a few elements far below the vector's largest one
            k[rng.integers(0, 64, 8)] *= np.float32(1e-6)
        exact = float((q.astype(np.float64) * k.astype(np.float64)).sum())
        got = split_dot(q, k)
        assert np.isfinite(got)
        scale = float(np.abs(q).max()) * float(np.abs(k).max()) * 8.0          # ~ sqrt(64) terms of the largest size
        worst = max(worst, abs(got - exact) / scale)
    # what is dropped is ~2^-22 per factor; an fp32 FMA chain over the same 64 terms carries ~2^-24 * sqrt(64)
    assert worst < 2.0 ** -20, worst


def test_normalised_vectors_sit_inside_fp16_range():
    for mx in (1e-38, 1e-30, 1e-3, 0.9999, 1.0, 1.5, 65504.0, 7e4, 3e38):
        down, up = range_factors(mx)
        assert down * up == 1.0
        if mx > 1e-16:                               # (below 2^(14-100) * 2^-14 the clamp of the exponent takes over)
            assert 2.0 ** 14 <= np.float32(mx) * down < 2.0 ** 15
        assert np.isfinite(np.float16(np.float32(mx) * down))
    down, up = range_factors(0.0)
    assert np.isfinite(down) and np.isfinite(up)


def test_channel_to_kslot_assignment_is_one_permutation_for_both_operands():
    """mma.m16n8k16: a register holds k-slots (2t, 2t+1) [a0/a1, b0] or (2t+8, 2t+9) [a2/a3, b1] of step s.  The kernel puts
    channels (16s + t, 16s + t + 4) into the first pair and (16s + t + 8, 16s + t + 12) into the second - for the query
    registers directly, for the keys through shared-memory pair-row p = 8s + 4 half + t.  Both must enumerate every
    channel exactly once, identically, and a B-fragment load (4 pair-rows x 8 positions) must touch 32 distinct banks."""
    a_slots, b_slots = {}, {}
    for s in range(4):
        for t in range(4):
            for half in range(2):
                # A side: xq[4s + j] = channel 16s + 4j + t; pairs (j = 0, 1) -> a0/a1, (j = 2, 3) -> a2/a3
                ca = [16 * s + 4 * j + t for j in (2 * half, 2 * half + 1)]
                # B side: pair-row p holds channels (c, c + 4) with c = 16 (p >> 3) + 8 ((p >> 2) & 1) + (p & 3)
                p = 8 * s + 4 * half + t
                c = 16 * (p >> 3) + 8 * ((p >> 2) & 1) + (p & 3)
                cb = [c, c + 4]
                for lane_half, (x, y) in enumerate(zip(ca, cb)):
                    slot = (s, 2 * t + lane_half + 8 * half)
                    a_slots[slot], b_slots[slot] = x, y
    assert a_slots == b_slots
    assert sorted(a_slots.values()) == list(range(64)) and len(a_slots) == 64
    ld = 72                                              # kTcLd, words per pair-row
    for s in range(4):
        for half in range(2):
            banks = {((8 * s + 4 * half + t) * ld + g) % 32 for t in range(4) for g in range(8)}
            assert len(banks) == 32
