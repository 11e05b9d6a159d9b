"""The admission thresholds of the tcgen05 filter and of the finalizer, restated in numpy (host-only property tests).

`score_select_kernel` (evavos_b200/csrc/score_tc.cu) partitions the memory positions of a SAMPLE of the key tiles
(every R-th tile; R = 1: all of them) into column classes, keeps the maximum approximate score of every class
(threshold pass) and admits every position of the bank whose approximate score reaches
tau = (k-th largest class maximum) - 2 * eps  (candidate pass).  `finalize_kernel` then cuts the list with
theta - 2 * eps, theta = (a lower bound of) the k-th largest approximate score IN the list.  If |approx - exact| <= eps
for every position, both cuts keep the exact top-k.  The tests check that argument - for the class layouts the kernel
uses, for sampled threshold passes and for arbitrary partitions - and that the lists stay far below the 1024-entry cap.
"""
import numpy as np
import pytest


def class_of(n, layout):
    tile, col = n // 128, n % 128
    if layout == "lockstep":        # (tile parity, 32-column group, column mod 16)
        return (tile % 2) * 64 + (col // 32) * 16 + col % 16
    if layout == "two_groups":      # (tile parity, 64-column half, 32-column block, column mod 16)
        return (tile % 2) * 64 + (col // 64) * 32 + ((col % 64) // 32) * 16 + col % 16
    if layout == "three_groups":    # (tile mod 3, half, pairs of (16-column block, column mod 8) classes merged)
        own = ((col % 64) // 16) * 8 + col % 8
        return (tile % 3) * 32 + (col // 64) * 16 + own // 2
    raise ValueError(layout)


@pytest.mark.parametrize("layout", ["lockstep", "two_groups", "three_groups"])
@pytest.mark.parametrize("n_pos,k", [(8100, 50), (32400, 50), (300, 50), (129, 50), (5000, 96)])
def test_threshold_admits_exact_topk(layout, n_pos, k):
    rng = np.random.default_rng(n_pos + k)
    n_cls = 96 if layout == "three_groups" else 128
    cls = np.array([class_of(n, layout) for n in range(n_pos)])
    assert cls.max() < n_cls
    for trial in range(8):
        exact = rng.normal(size=n_pos) * 3.0
        eps = 0.05
        approx = exact + rng.uniform(-eps, eps, size=n_pos)           # the bf16 contraction, error bounded by eps
        cmax = np.full(n_cls, -1e30)
        np.maximum.at(cmax, cls, approx)
        kth_class = np.sort(cmax)[::-1][k - 1] if n_cls >= k else -1e30
        tau = kth_class - 2 * eps
        admitted = np.nonzero(approx >= tau)[0]
        topk = np.argsort(-exact)[:k]
        assert set(topk) <= set(admitted)                             # nothing of the exact top-k is lost
        if kth_class > -1e29 and n_pos >= 5000 and k <= n_cls // 2:
            assert len(admitted) < 256                                # far from the candidate cap for i.i.d. scores
        # (k close to the number of classes degrades the bound: the list overflows and the exact fallback re-does
        #  the query - correct, slow; only the 96-class build option gets there below k = 128)


def test_any_partition_is_a_valid_bound():
    """The k-th largest class maximum never exceeds the k-th best score, whatever the partition."""
    rng = np.random.default_rng(3)
    for trial in range(50):
        n, k, c = 2000, 50, int(rng.integers(50, 300))
        s = rng.normal(size=n)
        cls = rng.integers(0, c, size=n)
        cmax = np.full(c, -np.inf)
        np.maximum.at(cmax, cls, s)
        assert np.sort(cmax)[::-1][k - 1] <= np.sort(s)[::-1][k - 1] + 1e-12


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 8])
def test_sampled_threshold_and_list_cut_keep_exact_topk(stride):
    """Threshold pass over every `stride`-th 128-position tile, candidate pass over everything, then the finalizer's
    cut with the k-th largest listed score truncated to 20 leading bits of its order-preserving key."""
    rng = np.random.default_rng(100 + stride)
    n_pos, k, eps = 40000, 50, 0.05
    tile = np.arange(n_pos) // 128
    cls = np.array([class_of(n, "two_groups") for n in range(n_pos)])
    sampled = tile % stride == 0
    for trial in range(6):
        exact = rng.normal(size=n_pos) * 3.0
        approx = (exact + rng.uniform(-eps, eps, size=n_pos)).astype(np.float32)
        cmax = np.full(128, -1e30)
        np.maximum.at(cmax, cls[sampled], approx[sampled])
        tau = np.sort(cmax)[::-1][k - 1] - 2 * eps
        listed = np.nonzero(approx >= tau)[0]
        topk = set(np.argsort(-exact)[:k])
        assert topk <= set(listed)
        assert len(listed) >= k
        if stride <= 4:      # (a sparser sample may overflow the 1024-entry list: correct, the exact path takes over)
            assert len(listed) < 1024
        # finalizer: theta = k-th largest listed score, truncated like kth_largest_bound (12 low key bits dropped)
        kth = np.sort(approx[listed])[::-1][k - 1]
        u = np.float32(kth).view(np.uint32)
        key = (~u) if (u & 0x80000000) else (u | np.uint32(0x80000000))
        key = np.uint32(key) & np.uint32(0xFFFFF000)
        back = (key & np.uint32(0x7FFFFFFF)) if (key & 0x80000000) else ~key
        theta = np.uint32(back).view(np.float32)
        assert theta <= kth
        survivors = listed[approx[listed] >= theta - 2 * eps]
        assert topk <= set(survivors)
        assert len(survivors) < 256
