#!/usr/bin/env python
"""Where does the fast kernel's factor start to deviate?  Single term of a dumped case, factor mode,
fast kernel against the reference-order kernel: d_n per step."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gadfly_b200 import solver as S
from gadfly_b200.solver import Geometry, KernelBatch, default_solver

d = np.load(sys.argv[1], allow_pickle=True)
scan = [np.asarray(x, dtype=np.float64) for x in d["scan"]]
t = d["t"]
N = len(t)
j = int(sys.argv[2]) if len(sys.argv) > 2 else 43
ac, bc, cc, dc = (scan[i][j:j + 1] for i in (2, 3, 4, 5))
kb = object.__new__(KernelBatch)
kb.B = 1
kb.coef = np.ascontiguousarray(np.stack([ac, bc, cc, dc], 1))
kb.base = kb.coef.copy()
kb.j_off = np.array([0, 1], dtype=np.int64)
kb.ddiag = np.array([0.0]); kb.delta = np.array([0.0])
dg = np.full(N, 1e-2 * ac[0])
solver = default_solver()
geom = Geometry.shared_t(1, N)
for name, tt in (("cumsum grid", t), ("arange grid", t[0] + np.arange(N) * 8.64e-5)):
    d_f, W_f, _, _, _ = solver.factor(kb, geom, tt, dg, flags=S.FLAG_WIDE_KERNEL)
    d_r, W_r, _, _, _ = solver.factor(kb, geom, tt, dg, flags=S.FLAG_REFERENCE_ORDER)
    rel = np.abs(d_f / d_r - 1)
    wrel = np.abs(W_f - W_r).reshape(N, 2).max(1) / np.abs(W_r).max()
    idx = [1, 7, 8, 9, 15, 16, 17, 63, 64, 65, 127, 128, 129, 500, 1000, 2000]
    print(name, "d dev at n:", " ".join(f"{n}:{rel[n]:.1e}" for n in idx if n < N))
    print(name, "W dev at n:", " ".join(f"{n}:{wrel[n]:.1e}" for n in idx if n < N))
    Wf, Wr = W_f.reshape(N, 2), W_r.reshape(N, 2)
    for n in (500, 1000, 2000):
        if n < N:
            amp = np.hypot(*Wf[n]) / np.hypot(*Wr[n]) - 1
            ang = np.arctan2(Wf[n][1], Wf[n][0]) - np.arctan2(Wr[n][1], Wr[n][0])
            print(f"   n={n}: W amplitude dev {amp:.2e}, rotation {ang:.2e} rad, d dev {d_f[n] / d_r[n] - 1:.2e}")
