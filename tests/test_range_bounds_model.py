"""Model check (numpy only) of the bound DESIGN.md §7 plans to use for the whole-range order statistic: with the range cut into nw
equal windows, the k-th largest key of the range lies between the smallest and the largest of the windows' own order statistics
   lower: min_w (m-th largest of window w),   m  = ceil(k / nw)
   upper: max_w (m'-th largest of window w),  m' = floor((k - 1) / nw) + 1
(pigeonhole: every window holds >= m keys >= its m-th largest; no window holds m' or more keys above its m'-th largest).  The same
holds with a short last window when its keys all count as candidates.  Keys are compared as the kernels do (IEEE totalOrder keys)."""
import numpy as np

from oracle import pyoracle as O


def _kth_largest(keys, k):
    return np.sort(keys)[len(keys) - k]


def test_window_statistics_bound_the_range_statistic():
    rng = np.random.default_rng(23)
    for trial in range(400):
        nw = int(rng.integers(1, 40))
        wp = int(rng.integers(1, 200))
        n = nw * wp
        shape = rng.integers(0, 4)
        if shape == 0:
            x = rng.normal(50.0, 5.0, n)
        elif shape == 1:
            x = np.floor(rng.uniform(30, 40, n))                       # heavy ties
        elif shape == 2:
            x = 60 + 20 * np.sin(np.arange(n) / max(1, n) * 9.0) + rng.normal(0, 0.5, n)      # drifting gauge
        else:
            x = rng.choice(np.array([0.0, -0.0, 1.0, np.inf, -np.inf, 5e-324, -7.5]), n)      # signed zeros, infinities, denormals
        keys = O.total_order_key(x)
        k = int(rng.integers(1, n + 1))
        m, m2 = -(-k // nw), (k - 1) // nw + 1
        assert 1 <= m <= wp and 1 <= m2 <= wp and m2 in (m, m + 1)
        win = keys.reshape(nw, wp)
        srt = np.sort(win, axis=1)
        lo = srt[:, wp - m].min()
        hi = srt[:, wp - m2].max()
        ans = _kth_largest(keys, k)
        assert lo <= ans <= hi, (trial, nw, wp, k)
        # what the second pass needs: rank inside the interval = k - #(keys above hi); the interval always contains the answer
        above = int((keys > hi).sum())
        inside = keys[(keys >= lo) & (keys <= hi)]
        assert above < k <= above + len(inside)
        assert _kth_largest(inside, k - above) == ans


def test_bound_with_a_short_last_window():
    rng = np.random.default_rng(29)
    for trial in range(300):
        nw = int(rng.integers(1, 20))
        wp = int(rng.integers(2, 100))
        s = int(rng.integers(1, wp))                                     # the partial window's length
        n = nw * wp + s
        keys = O.total_order_key(np.floor(rng.normal(70, 3, n) * 4) / 4)
        k = int(rng.integers(1, n + 1))
        full, short = keys[: nw * wp].reshape(nw, wp), keys[nw * wp:]
        srt = np.sort(full, axis=1)
        ans = _kth_largest(keys, k)
        m = -(-k // nw)                                                   # lower bound from the full windows alone, when they can carry rank k
        if m <= wp:
            assert srt[:, wp - m].min() <= ans
        if k - 1 - s >= 0:                                               # upper bound: the short window may sit entirely above
            m2 = (k - 1 - s) // nw + 1
            if m2 <= wp:
                T = srt[:, wp - m2].max()
                assert int((keys > T).sum()) < k and ans <= T               # at most nw (m2 - 1) + s <= k - 1 keys lie above T
