#!/usr/bin/env python
"""Kernel time per mode (log-likelihood / Philox sample / factor without W / factor with W) of the
scan kernel: 148 solar light curves x N points. usage: python tools/run_modes.py [N]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gadfly_b200 as g
from gadfly_b200.solver import Geometry, KernelBatch, Solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
solver = Solver(0)
B = 148
kb = KernelBatch([kernel] * B)
geom = Geometry.shared_t(B, N)
dev = torch.device("cuda", 0)
t = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
y = torch.randn(B * N, dtype=torch.float64, device=dev) * 285.0
x = torch.empty(B * N, dtype=torch.float64, device=dev)
W = torch.empty(B * N * kernel.J, dtype=torch.float64, device=dev)
ld = torch.empty(B, dtype=torch.float64, device=dev)
q = torch.empty(B, dtype=torch.float64, device=dev)
st = torch.empty(B, dtype=torch.int32, device=dev)
w_off = np.arange(B, dtype=np.int64) * N * kernel.J
cyc = lambda ms: ms * 1e-3 * 1.965e9 / N
for rep in range(2):
    solver.loglike(kb, geom, t, y, logdet=ld, quad=q, status=st); a = solver.last_kernel_ms
    solver.sample(kb, geom, t, seed=1, out=x, logdet=ld, status=st); b = solver.last_kernel_ms
    solver.factor(kb, geom, t, d=x, want_W=False, logdet=ld, status=st); c = solver.last_kernel_ms
    solver.factor(kb, geom, t, d=x, W=W, w_off=w_off, logdet=ld, status=st); d = solver.last_kernel_ms
print(f"cycles/step: loglike {cyc(a):.0f}  sample {cyc(b):.0f}  factor (d only) {cyc(c):.0f}  factor (d, W) {cyc(d):.0f}")
