#!/usr/bin/env python
"""Kernel time of the four sweeps (solve / matmul, lower / upper) on a stored factor of one solar light
curve of 2^18 points, in ms and cycles per step.  usage: python tools/sweep_time.py"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gadfly_b200 as g
from gadfly_b200.solver import Geometry, KernelBatch, Solver
k = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
s = Solver(0); N = 1 << 18
kb = KernelBatch([k]); geom = Geometry.shared_t(1, N)
dev = torch.device("cuda", 0)
t = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
W = torch.empty(N * k.J, dtype=torch.float64, device=dev)
d = torch.empty(N, dtype=torch.float64, device=dev)
w_off = np.zeros(1, dtype=np.int64)
s.factor(kb, geom, t, d=d, W=W, w_off=w_off)
y = torch.randn(N, dtype=torch.float64, device=dev); z = torch.empty_like(y)
for op, name in enumerate(["solve_lower", "matmul_lower", "solve_upper", "matmul_upper"]):
    for _ in range(2):
        s.sweep(op, kb, geom, w_off, t, W, y, Z=z)
    print(name, "%.1f ms  %.0f cycles/step" % (s.last_kernel_ms, s.last_kernel_ms * 1e-3 * 1.965e9 / N))
