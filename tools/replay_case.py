#!/usr/bin/env python
"""Replay one case dumped by tools/stress.py (gpurun_out/stress_case_*.npz: copy it to variants/ so that
it travels to the GPU box) through the fast and the reference-order kernels and the oracle, at several
lengths and on a few re-gridded time axes.  usage: python tools/replay_case.py variants/case.npz"""
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
kb = object.__new__(KernelBatch)
ar, cr, ac, bc, cc, dc, dd = scan
kb.B = 1
kb.coef = np.ascontiguousarray(np.concatenate([np.stack([ar, 0 * ar, cr, 0 * cr], 1), np.stack([ac, bc, cc, dc], 1)]))
kb.base = kb.coef.copy()
kb.j_off = np.array([0, len(ar) + len(ac)], dtype=np.int64)
kb.ddiag = np.array([dd]); kb.delta = np.array([0.0])
geom = Geometry.shared_t(1, N)
solver = default_solver()
x_ref = oracle.stream(1, scan, t, nrm, diag=diag)[0]
for name, fl in (("fast", 0), ("reference order", S.FLAG_REFERENCE_ORDER)):
    x, ld, st = solver.sample(kb, geom, t, diag, normals=nrm, flags=fl)
    dev = np.abs(x - x_ref) / np.max(np.abs(x_ref))
    print(f"{name:16s} status {st[0]}  max rel dev {dev.max():.2e} at n = {dev.argmax()}  "
          f"first n with dev > 1e-10: {int(np.argmax(dev > 1e-10)) if (dev > 1e-10).any() else -1}")
for n_cut in (64, 128, 512, 1024, 2048):
    if n_cut < N:
        x, ld, st = solver.sample(kb, Geometry.shared_t(1, n_cut), t[:n_cut], diag[:n_cut], normals=nrm[:n_cut])
        xr = oracle.stream(1, scan, t[:n_cut], nrm[:n_cut], diag=diag[:n_cut])[0]
        print(f"  N = {n_cut:5d}: fast max rel dev {np.max(np.abs(x - xr)) / np.max(np.abs(xr)):.2e}")
# bisect: exact producer rows (cadence jitter beyond the fast path's window) and other cadences
rng = np.random.default_rng(0)
for label, tt in (("jittered 1e-3 (exact rows)", np.cumsum(np.diff(t, prepend=0.0) * (1 + 1e-3 * rng.standard_normal(N)))),
                  ("1-min cadence", t[0] + np.arange(N) * 6e-5),
                  ("exact multiples of 2^-14", np.arange(1, N + 1) * 2.0 ** -14)):
    x, ld, st = solver.sample(kb, geom, tt, diag, normals=nrm)
    xr = oracle.stream(1, scan, tt, nrm, diag=diag)[0]
    print(f"{label:28s} fast max rel dev {np.max(np.abs(x - xr)) / np.max(np.abs(xr)):.2e}")
