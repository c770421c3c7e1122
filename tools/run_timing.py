#!/usr/bin/env python
"""Run one batched log-likelihood (148 solar light curves x 16384 points) so that a library built with
-DGF_TIMING prints its per-role cycle counts.  usage: GADFLY_B200_LIB=variants/libgadfly_b200_tm.so python tools/run_timing.py"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gadfly_b200 as g
from gadfly_b200.solver import Geometry, KernelBatch, Solver
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
solver = Solver(0)
B, n = 148, 16384
kb = KernelBatch([kernel] * B); geom = Geometry.shared_t(B, n)
dev = torch.device("cuda", 0)
t = torch.arange(n, dtype=torch.float64, device=dev) * 6e-5
y = torch.randn(B * n, dtype=torch.float64, device=dev) * 285.0
logdet = torch.empty(B, dtype=torch.float64, device=dev); quad = torch.empty(B, dtype=torch.float64, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev)
solver.loglike(kb, geom, t, y, logdet=logdet, quad=quad, status=status)
torch.cuda.synchronize()
print("ms", solver.last_kernel_ms, "cyc/step", solver.last_kernel_ms * 1e-3 * 1.965e9 / n)
