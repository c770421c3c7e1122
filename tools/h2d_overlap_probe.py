#!/usr/bin/env python
"""Does a host-to-device copy overlap a running kernel on this box?  (a) a long torch kernel (FP64 matmul
loop), (b) the fused scan kernel.  Copy issued 100 ms into the kernel on a fresh stream; host clock."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda", 0)
n = 1 << 27
src = torch.empty(n, dtype=torch.float64).pin_memory()
dst = torch.empty(n, dtype=torch.float64, device=dev)
back = torch.empty(n, dtype=torch.float64).pin_memory()
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
fresh = torch.cuda.Stream(device=dev)


def run(label, launch, sync):
    for direction in ("H2D", "D2H"):
        launch(); sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        launch()
        time.sleep(0.1)
        with torch.cuda.stream(fresh):
            if direction == "H2D":
                dst.copy_(src, non_blocking=True)
            else:
                back.copy_(dst, non_blocking=True)
        fresh.synchronize()
        t1 = time.perf_counter()
        sync()
        t2 = time.perf_counter()
        print(f"{label}: {direction} of 1 GiB issued at 100 ms: copy done at {(t1 - t0) * 1e3:.0f} ms, kernel done at {(t2 - t0) * 1e3:.0f} ms")


def mm():
    for _ in range(12):
        torch.matmul(a, a)


run("torch FP64 matmul loop", mm, torch.cuda.synchronize)

import gadfly_b200 as g
from gadfly_b200 import solver as S
from gadfly_b200.solver import Geometry, KernelBatch, Solver
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
solver = Solver(0)
for B in (148, 100):
    N = 1 << 19
    kb = KernelBatch([kernel] * B)
    geom = Geometry.shared_t(B, N)
    t_dev = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
    x_dev = torch.empty(B * N, dtype=torch.float64, device=dev)
    run(f"fused scan, {B} CTAs", lambda: solver.sample(kb, geom, t_dev, seed=2, out=x_dev, flags=S.FLAG_ASYNC),
        solver.synchronize)
