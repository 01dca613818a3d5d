"""Synthetic counter streams generated ON THE DEVICE (torch), for the bench-sized ring (512 fields x 1 Mi samples = 4 GiB).

TEST / BENCH INFRASTRUCTURE.  The streams are the ones SURVEY.md 8(d) defines:

  survey   field k is a gauge  base_k + A_k sin(2 pi t / P_k) + sigma_k N(0,1)  with base in {30..90 C, 50..1000 W, 0..100 %}
           (utilisation clamped to [0, 100]), 0.1 % spikes above the field's threshold so that n_over != 0, and every 8th
           field a monotone u64 counter (cumulative sum of integer increments) cast to float64
  uniform  white noise in [30, 90)                      (no ties, every window takes the fast path)
  mw       integer readings uniform in [30000, 90000)   (a power gauge in mW: few ties)
  temp     integer degrees uniform in [30, 90)          (about 16 copies of every value per window)
  const    a flat gauge (utilisation pinned at 100 %)
  walk     a slow random walk in integer steps          (long runs of equal values, trends)

`chunks(shape, F, cap, chunk, seed, dev)` yields ([chunk, F] float64 device tensor, rows) in chronological order; `thresholds(shape,
F, cap, seed)` gives the per-field thresholds that go with it.  Deterministic for a given (shape, F, cap, seed).
"""
import numpy as np

SHAPES = ("survey", "uniform", "mw", "temp", "const", "walk")


def survey_params(F: int, cap: int, seed: int):
    rng = np.random.default_rng(seed)
    kind = np.arange(F) % 3                       # 0: temperature C, 1: power W, 2: utilisation %
    u = rng.random((6, F))
    base = np.where(kind == 0, 30 + 60 * u[0], np.where(kind == 1, 50 + 950 * u[0], 100 * u[0]))
    amp = np.where(kind == 0, 1 + 4 * u[1], np.where(kind == 1, base * (0.02 + 0.08 * u[1]), 2 + 13 * u[1]))
    sigma = np.where(kind == 0, 0.1 + 0.9 * u[2], np.where(kind == 1, base * (0.002 + 0.018 * u[2]), 0.5 + 3.5 * u[2]))
    period = 10.0 ** (3 + 3 * u[3])
    thr = base + amp + 5 * sigma
    counter = (np.arange(F) % 8) == 7
    cbase = np.floor(1e9 * (1 + np.arange(F)))   # where each monotone counter starts
    thr = np.where(counter, cbase + 0.999 * 1000.0 * cap, thr)   # increments average 1000: the last 0.1 % of the range is over
    return {"kind": kind, "base": base, "amp": amp, "sigma": sigma, "period": period, "thr": thr, "counter": counter, "cbase": cbase}


def thresholds(shape: str, F: int, cap: int, seed: int) -> np.ndarray:
    if shape == "survey":
        return survey_params(F, cap, seed)["thr"].astype(np.float64)
    return np.full(F, {"uniform": 88.0, "mw": 88000.0, "temp": 88.0, "const": 99.0, "walk": 75.0}[shape])


def chunks(shape: str, F: int, cap: int, chunk: int, seed: int, dev):
    import torch
    gen = torch.Generator(device=dev).manual_seed(seed)
    f64 = torch.float64
    if shape == "survey":
        p = survey_params(F, cap, seed)
        T = {k: torch.tensor(p[k], dtype=f64, device=dev)[None, :] for k in ("base", "amp", "sigma", "period", "thr", "cbase")}
        is_util = torch.tensor(p["kind"] == 2, device=dev)[None, :]
        is_counter = torch.tensor(p["counter"], device=dev)[None, :]
        carry = torch.zeros((1, F), dtype=f64, device=dev)
    walk = torch.zeros((1, F), dtype=f64, device=dev)
    for t0 in range(0, cap, chunk):
        n = min(chunk, cap - t0)
        if shape == "survey":
            t = torch.arange(t0, t0 + n, dtype=f64, device=dev)[:, None]
            x = T["base"] + T["amp"] * torch.sin(t * (2.0 * np.pi) / T["period"]) + T["sigma"] * torch.randn((n, F), dtype=f64, device=dev, generator=gen)
            x = torch.where(is_util, x.clamp(0.0, 100.0), x)
            spike = torch.rand((n, F), device=dev, generator=gen) < 0.001
            mag = T["thr"] + T["sigma"] * (1.0 + 3.0 * torch.rand((n, F), dtype=f64, device=dev, generator=gen))
            x = torch.where(spike, mag, x)
            inc = torch.randint(0, 2001, (n, F), dtype=torch.int64, device=dev, generator=gen).to(f64)
            c = carry + torch.cumsum(inc, 0)
            carry = c[-1:].clone()
            x = torch.where(is_counter, T["cbase"] + c, x)
        elif shape == "mw":
            x = torch.randint(30000, 90000, (n, F), dtype=torch.int32, device=dev, generator=gen).to(f64)
        elif shape == "temp":
            x = torch.randint(30, 90, (n, F), dtype=torch.int32, device=dev, generator=gen).to(f64)
        elif shape == "const":
            x = torch.full((n, F), 100.0, dtype=f64, device=dev)
        elif shape == "walk":
            steps = torch.randint(-1, 2, (n, F), dtype=torch.int32, device=dev, generator=gen).to(f64)
            x = walk + torch.cumsum(steps * (torch.rand((n, F), device=dev, generator=gen) < 0.05), 0)
            walk = x[-1:].clone()
            x = x + 60.0
        elif shape == "uniform":
            x = torch.rand((n, F), dtype=f64, device=dev, generator=gen) * 60.0 + 30.0
        else:
            raise ValueError(shape)
        yield x, n


def fill_ring(ring, shape: str, F: int, cap: int, seed: int, dev, chunk: int = 1 << 16, host_copy=None):
    """Append the whole stream through the real append kernel.  host_copy: optional [F, cap] float64 numpy array that receives the
    same samples field-major (for the CPU oracle)."""
    import torch
    t0 = 0
    for x, n in chunks(shape, F, cap, chunk, seed, dev):
        x = x.contiguous()
        torch.cuda.current_stream().synchronize()      # the ring appends on ITS stream: the chunk must be complete before the append reads it
        ring.push_device(x.data_ptr(), n)
        ring.sync()
        if host_copy is not None:
            host_copy[:, t0:t0 + n] = x.T.contiguous().cpu().numpy()
        t0 += n
    torch.cuda.synchronize()
