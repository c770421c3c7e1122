#!/usr/bin/env python
"""Bisecting the fast path's deviation on a dumped stress case: grids and single terms."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gadfly_b200 import solver as S
from gadfly_b200.solver import Geometry, KernelBatch, default_solver
import oracle

d = np.load(sys.argv[1], allow_pickle=True)
scan = tuple(np.asarray(x, dtype=np.float64) for x in d["scan"]) + (float(d["ddiag"]),)
t, diag, nrm = d["t"], d["diag"], d["nrm"]
N = len(t)
solver = default_solver()


def batch_of(sc):
    ar, cr, ac, bc, cc, dc, dd = sc
    kb = object.__new__(KernelBatch)
    kb.B = 1
    kb.coef = np.ascontiguousarray(np.concatenate([np.stack([ar, 0 * ar, cr, 0 * cr], 1), np.stack([ac, bc, cc, dc], 1)]))
    kb.base = kb.coef.copy()
    kb.j_off = np.array([0, len(ar) + len(ac)], dtype=np.int64)
    kb.ddiag = np.array([dd]); kb.delta = np.array([0.0])
    return kb


def dev(sc, tt, dg, flags=S.FLAG_WIDE_KERNEL):
    x, ld, st = solver.sample(batch_of(sc), Geometry.shared_t(1, len(tt)), tt, dg, normals=nrm[:len(tt)], flags=flags)
    xr = oracle.stream(1, sc, tt, nrm[:len(tt)], diag=dg)[0]
    return np.max(np.abs(x - xr)) / np.max(np.abs(xr))


dt = 8.64e-5
grids = {"cumsum (original)": t,
         "t0 + arange * dt": t[0] + np.arange(N) * dt,
         "arange(1..N) * dt": np.arange(1, N + 1) * dt,
         "cumsum, dt = 6e-5": np.cumsum(np.full(N, 6e-5)),
         "cumsum + 1e-10 jitter": np.cumsum(np.full(N, dt) * (1 + 1e-10 * np.random.default_rng(1).standard_normal(N)))}
for name, tt in grids.items():
    print(f"{name:26s} {dev(scan, tt, diag):.2e}")
ar, cr, ac, bc, cc, dc, dd = scan
for j in (0, 10, 20, 30, 43):
    sc = (ar, cr, ac[j:j + 1], bc[j:j + 1], cc[j:j + 1], dc[j:j + 1], 0.0)
    dg = np.full(N, 1e-2 * ac[j])
    print(f"term {j}: c = {cc[j]:.3g} d = {dc[j]:.6g}  dev {dev(sc, t, dg):.2e}   (arange grid {dev(sc, t[0] + np.arange(N) * dt, dg):.2e})")
