#!/usr/bin/env python
"""Cost of k realisations per light curve on ONE factor (gf_sample_multi) against one fused
realisation (gf_sample_batched) and against k fused ones.  usage: python tools/multi_rhs_time.py [N] [B]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gadfly_b200 as g
from gadfly_b200.solver import Geometry, KernelBatch, Solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
solver = Solver(0)
dev = torch.device("cuda", 0)
kb = KernelBatch([kernel] * B)
geom = Geometry.shared_t(B, N)
t = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


x1 = torch.empty(B * N, dtype=torch.float64, device=dev)
one = timed(lambda: solver.sample(kb, geom, t, seed=1, out=x1))
print(f"B = {B}, N = {N}: one fused realisation per light curve {one * 1e3:.2f} ms")
for k in (1, 2, 8, 32):
    xk = torch.empty(B * N * k, dtype=torch.float64, device=dev)
    tk = timed(lambda: solver.sample_multi(kb, geom, t, k, seed=1, out=xk))
    print(f"  k = {k:3d} on one factor: {tk * 1e3:8.2f} ms = {tk / one:5.2f} x one fused realisation "
          f"(k fused ones: {k:.0f} x)")
    del xk
