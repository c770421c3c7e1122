"""
Synthetic inputs of the BASELINE.json configurations, shared by ``bench.py``,
``tools/bench_configs.py`` and the BASELINE-size GPU tests:

  cfg2 / cfg5  Kepler-like stars drawn with replacement from the Huber-2011 table shipped with the
               reference (notebooks/huber2011.ecsv, re-serialised in data/huber2011_stars.csv),
               jittered by the catalogue errors, flat bandpass (alpha = 1)
  cfg4         the solar kernel with (S0, w0, Q) scaled by up to +-10 % on a 3-D lattice
"""
import os
import time

import numpy as np

__all__ = ["star_table", "kepler_like_stars", "kepler_like_batch", "lattice_batch"]


def star_table():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "huber2011_stars.csv")
    return np.genfromtxt(path, delimiter=",", names=True, skip_header=1)


def kepler_like_stars(n, seed):
    """(M, R, T, L) of the cfg2 / cfg5 population: rows of the Huber-2011 table drawn with
    replacement, jittered by their sig_* columns."""
    tab = star_table()
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, len(tab), n)
    z = rng.standard_normal((n, 4))
    r = tab[rows]
    M = np.maximum(r["mass"] + z[:, 0] * r["sig_mass"], 0.3)
    R = np.maximum(r["rad"] + z[:, 1] * r["sig_rad"], 0.3)
    T = np.maximum(r["teff"] + z[:, 2] * r["sig_teff"], 3500.0)
    L = np.maximum(r["lum"] + z[:, 3] * r["sig_lum"], 0.05)
    return M, R, T, L


def kepler_like_batch(n, seed, solver=None):
    """cfg2 / cfg5 population built by the batched feeder (gadfly_b200/feeder.py): on the host, or
    with ``solver=`` on that solver's GPU (csrc/feed.cu).  Returns (KernelBatch, feeder seconds)."""
    from . import feeder
    M, R, T, L = kepler_like_stars(n, seed)
    t0 = time.perf_counter()
    kb = feeder.kernel_batch_for_stars(M, R, T, L, texp_s=60.0, bandpass='SOHO VIRGO', solver=solver)
    return kb, time.perf_counter() - t0


def lattice_batch(n, seed, lo=0, hi=None, solver=None):
    """cfg4 grid: solar hyper-parameters with S0, w0, Q of every term scaled by lattice factors
    0.9 .. 1.1 (a random subset of n points of the side^3 lattice).  ``lo, hi``: build only the grid
    points [lo, hi) of that list (a rank's shard); ``solver=``: coefficients on that solver's GPU
    (gf_feed_sho).  Returns (KernelBatch, feeder seconds)."""
    from . import feeder
    from .core import Hyperparameters
    hp = Hyperparameters.for_sun()
    S0, w0, Q = (np.array([q['hyperparameters'][k] for q in hp]) for k in ('S0', 'w0', 'Q'))
    side = int(np.ceil(n ** (1.0 / 3.0)))
    f = np.linspace(0.9, 1.1, side)
    pts = np.random.default_rng(seed).permutation(side ** 3)[:n][lo:hi]
    m = len(pts)
    i, j, k = pts // (side * side), (pts // side) % side, pts % side
    t0 = time.perf_counter()
    hpb = feeder.HyperparameterBatch((S0[None, :] * f[i][:, None]).ravel(), (w0[None, :] * f[j][:, None]).ravel(),
                                     (Q[None, :] * f[k][:, None]).ravel(),
                                     np.arange(m + 1, dtype=np.int64) * len(S0))
    kb = feeder.kernel_batch_from_sho(hpb, 6e-5, solver=solver)
    return kb, time.perf_counter() - t0
