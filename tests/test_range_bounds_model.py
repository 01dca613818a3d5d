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


def _emulate_sel2(keys, nw, wp, k):
    """line-by-line emulation of k_sel2_bounds / k_sel2_collect / k_sel2_final (csrc/select.cu) with Python integers"""
    win = keys.reshape(nw, wp)
    m = -(-k // nw)
    stat = np.sort(win, axis=1)[:, wp - m]                       # what the ranked window pass reports
    t, T = int(stat.min()), int(stat.max())
    above, lst = 0, []
    for key in map(int, keys):
        if key > T:
            above += 1
        elif key >= t:
            lst.append(key)
    diff = t ^ T
    cp = 64 if diff == 0 else 64 - diff.bit_length()             # __clzll
    nbits = cp
    prefix = 0 if cp == 0 else (T if cp == 64 else T >> (64 - cp))
    kk = k - above
    assert kk >= 1
    while nbits < 64:
        d = min(11, 64 - nbits)
        shift, nb = 64 - nbits - d, 1 << d
        hist = [0] * nb
        for key in lst:
            if nbits == 0 or (key >> (64 - nbits)) == prefix:
                hist[(key >> shift) & (nb - 1)] += 1
        acc, b = 0, nb - 1
        while b > 0 and acc + hist[b] < kk:
            acc += hist[b]
            b -= 1
        prefix = ((prefix << d) if nbits else 0) | b
        nbits += d
        kk -= acc
    return prefix, len(lst)


def test_emulated_bounded_select_equals_sort():
    rng = np.random.default_rng(31)
    for trial in range(120):
        nw, wp = int(rng.integers(1, 12)), int(rng.integers(1, 64))
        n = nw * wp
        shape = rng.integers(0, 4)
        if shape == 0:
            x = rng.normal(0.0, 1.0, n)                          # both signs: keys span the sign bit (no common prefix)
        elif shape == 1:
            x = np.floor(rng.uniform(30, 34, n))
        elif shape == 2:
            x = np.full(n, 7.25)
        else:
            x = rng.uniform(1e-300, 1e300, n) * rng.choice([-1.0, 1.0], n)
        keys = O.total_order_key(x)
        k = int(rng.integers(1, n + 1))
        got, n_list = _emulate_sel2(keys, nw, wp, k)
        assert got == int(np.sort(keys)[n - k]), (trial, nw, wp, k)
        assert 1 <= n_list <= n
