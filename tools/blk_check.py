#!/usr/bin/env python
"""The 4-step blocked scan kernel (GF_FLAG_BLOCKED, csrc/scan_blk.cu) against the default fused scan:
log det / quadratic form / samples on a set of awkward cases, then kernel times.
usage: python tools/blk_check.py [n_points_for_timing]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import solver as S
from gadfly_b200.solver import Geometry, KernelBatch, Solver

solver = Solver(0)
solar = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
narrow = g.StellarOscillatorKernel(terms=list(solar.term.terms)[:22], delta=solar.delta)
rng = np.random.default_rng(0)
worst = 0.0


def compare(label, kernels, t, diag=None):
    global worst
    B, N = len(kernels), len(t)
    kb = KernelBatch(kernels)
    geom = Geometry.shared_t(B, N)
    k0 = np.array([np.sum(kb.coef[kb.j_off[b]:kb.j_off[b + 1], 0]) + kb.ddiag[b] for b in range(B)])
    y = (rng.standard_normal((B, N)) * np.sqrt(k0)[:, None]).ravel()
    if diag is not None:
        diag = np.tile(diag, B)
    ld0, q0, s0 = solver.loglike(kb, geom, t, y, diag=diag)
    ld1, q1, s1 = solver.loglike(kb, geom, t, y, diag=diag, flags=S.FLAG_BLOCKED)
    x0, _, _ = solver.sample(kb, geom, t, diag=diag, seed=5)
    x1, _, sx = solver.sample(kb, geom, t, diag=diag, seed=5, flags=S.FLAG_BLOCKED)
    same_status = np.array_equal(s0, s1)
    ok = s0 == 0
    rel = 0.0
    if ok.any():
        rel = max(np.max(np.abs(ld1[ok] - ld0[ok]) / np.maximum(np.abs(ld0[ok]), 1.0)),
                  np.max(np.abs(q1[ok] - q0[ok]) / np.maximum(np.abs(q0[ok]), 1.0)))
        xs = x0.reshape(B, N)[ok]
        rel = max(rel, np.max(np.abs(x1.reshape(B, N)[ok] - xs)) / np.max(np.abs(xs)))
    worst = max(worst, rel)
    print(f"{label:50s} B={B:3d} N={N:6d} status equal {same_status} {s0[:4]} {s1[:4]}  max rel {rel:.2e}", flush=True)
    return same_status and rel < 1e-9


allok = True
for N in (1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 100, 1000, 4099):
    allok &= compare("solar, uniform 1-min", [solar] * 3, np.arange(N) * 6e-5)
allok &= compare("solar + narrow mixed", [solar, narrow, solar, narrow, narrow], np.arange(3000) * 6e-5)
t = np.cumsum(rng.choice([6e-5, 6e-5, 6e-5, 1.2e-4, 3e-3, 0.5], 5000))
allok &= compare("solar, gaps up to half a day", [solar] * 4, t, diag=np.full(5000, 30.0 ** 2))
t = np.cumsum(rng.uniform(3e-5, 2e-4, 3000))
allok &= compare("narrow, jittered cadence", [narrow] * 5, t)
allok &= compare("solar, 30-min cadence (frame change every other row)", [solar] * 2, np.arange(2000) * 1.8e-3)
allok &= compare("solar, absolute time stamps", [solar] * 2, 2.1e5 + np.arange(3000) * 6e-5)
# not positive definite: a negative diagonal makes a pivot fail somewhere
bad = np.full(2000, 25.0); bad[1234:] = -4.0e5
allok &= compare("solar, negative diagonal from row 1234", [solar] * 2, np.arange(2000) * 6e-5, diag=bad)
print("all ok" if allok else "MISMATCH", "worst", worst)

# timing
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
B = 148
kb = KernelBatch([solar] * B)
geom = Geometry.shared_t(B, N)
import torch
dev = torch.device("cuda", 0)
t_d = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
y_d = torch.randn(B * N, dtype=torch.float64, device=dev) * 285.0
x_d = torch.empty(B * N, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for name, fl in (("default", 0), ("blocked", S.FLAG_BLOCKED)):
    ms = []
    for rep in range(3):
        solver.loglike(kb, geom, t_d, y_d, flags=fl)
        ms.append(solver.last_kernel_ms)
    ms2 = []
    for rep in range(3):
        solver.sample(kb, geom, t_d, seed=3, out=x_d, flags=fl)
        ms2.append(solver.last_kernel_ms)
    cyc = min(ms) * 1e-3 * 1.965e9 / N
    print(f"{name}: loglike {min(ms):.3f} ms ({cyc:.0f} cycles per step), sample {min(ms2):.3f} ms ({min(ms2) * 1e-3 * 1.965e9 / N:.0f})")
