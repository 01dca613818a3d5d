"""Model check (numpy only) of the positional shortcut of k_window_reduce (gpud_b200/csrc/ring.cu): for a window x of m
samples and a rank k, IF every sample is non-negative (sign bit clear), every sample before t = m - k lies in [x[0], T] and every sample
from t on lies in [T, x[m-1]] (T = x[m-k], compared as doubles), THEN in IEEE totalOrder - the order the kernels select in -
   the minimum is x[0], the maximum is x[m-1] and the k-th largest is x[m-k], bit for bit.
The condition is evaluated here exactly as the kernel does (double compares that NaN fails, sign test on the raw bits); the claim is
checked against a sort of the totalOrder keys.  The sign condition is necessary: a -0.0 among +0.0 passes every compare."""
import numpy as np

from oracle import pyoracle as O


def _condition(x, k):
    m = len(x)
    if np.any(x.view(np.uint64) >> np.uint64(63)):           # any sign bit (negative values, -0.0, negative NaN)
        return False
    tk = m - k
    e0, tv, el = x[0], x[tk], x[m - 1]
    with np.errstate(invalid="ignore"):
        lo_ok = (e0 <= x[:tk]) & (x[:tk] <= tv)
        hi_ok = (tv <= x[tk:]) & (x[tk:] <= el)
    return bool(lo_ok.all() and hi_ok.all())


def _claim(x, k):
    keys = O.total_order_key(x)
    srt = np.sort(keys)
    m = len(x)
    return keys[0] == srt[0] and keys[m - 1] == srt[m - 1] and keys[m - k] == srt[m - k]


def test_condition_implies_positions():
    rng = np.random.default_rng(77)
    taken = 0
    for trial in range(3000):
        m = int(rng.integers(2, 200))
        k = int(rng.integers(1, m + 1))
        shape = rng.integers(0, 6)
        if shape == 0:
            x = 2.0 ** 39 + np.cumsum(rng.integers(0, 2001, m)).astype(np.float64)              # a counter
        elif shape == 1:
            x = np.sort(rng.integers(0, 5, m)).astype(np.float64)                                # long runs of equal values
        elif shape == 2:
            x = np.sort(rng.random(m))
            i, j = rng.integers(0, m, 2)
            x[i], x[j] = x[j], x[i]                                                              # one swap: sometimes still positional
        elif shape == 3:
            x = np.sort(rng.choice(np.array([0.0, 5e-324, 1.0, 2.0 ** 52, np.inf]), m))
        elif shape == 4:
            x = np.sort(rng.random(m))
            x[rng.integers(0, m)] = np.nan                                                       # NaN fails every compare
        else:
            x = np.sort(rng.random(m))
            lo = x[: m - k].copy()
            rng.shuffle(lo)                                                                      # any order below sample m - k ...
            x[: m - k] = lo
            x[0] = x[: max(1, m - k)].min() if m - k > 0 else x[0]                               # ... as long as the first stays the minimum
        x = np.ascontiguousarray(x, dtype=np.float64)
        if _condition(x, k):
            taken += 1
            assert _claim(x, k), (trial, shape, m, k)
    assert taken > 800


def test_the_sign_condition_is_needed():
    x = np.array([0.0, -0.0, 0.0, 0.0])          # every double compare holds, but -0.0 is the smallest totalOrder key and sits at t = 1
    k = 1
    m = len(x)
    with np.errstate(invalid="ignore"):
        compares = bool(((x[0] <= x[: m - k]) & (x[: m - k] <= x[m - k])).all() and ((x[m - k] <= x[m - k:]) & (x[m - k:] <= x[m - 1])).all())
    assert compares and not _claim(x, k) and not _condition(x, k)
