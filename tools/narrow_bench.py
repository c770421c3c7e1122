#!/usr/bin/env python
"""Throughput of the one-warp-per-sequence path (J <= 32, csrc/scan_small.cu) against the wide
kernel on the same batch: granulation-only solar kernel (J = 10) and random 8- / 16-term kernels.
usage: python tools/narrow_bench.py [B] [N]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gadfly_b200 as g
from gadfly_b200 import solver as S
from gadfly_b200.solver import Geometry, KernelBatch, Solver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
solver = Solver(0)
dev = torch.device("cuda", 0)
sun = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
rng = np.random.default_rng(17)


def sho(n):
    return [g.SHOTerm(S0=float(10 ** rng.uniform(0, 3)), w0=float(10 ** rng.uniform(0.5, 3.5)),
                      Q=float(10 ** rng.uniform(-0.2, 2))) for _ in range(n)]


cases = {"granulation J=10": g.StellarOscillatorKernel(terms=list(sun.term.terms[:5]), delta=sun.delta),
         "8 terms J=16": g.StellarOscillatorKernel(terms=sho(8), delta=6e-5),
         "16 terms J=32": g.StellarOscillatorKernel(terms=sho(16), delta=6e-5)}
t = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
y = torch.randn(N, dtype=torch.float64, device=dev) * 100
for name, k in cases.items():
    kb = KernelBatch([k] * B)
    kb.ddiag = kb.ddiag + 100.0
    geom = Geometry.shared_t(B, N)
    ld = torch.empty(B, dtype=torch.float64, device=dev)
    q = torch.empty(B, dtype=torch.float64, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    res = {}
    for label, fl in (("narrow", S.FLAG_SHARED_Y), ("wide", S.FLAG_SHARED_Y | S.FLAG_WIDE_KERNEL)):
        Bx = B if label == "narrow" else min(B, 296)
        kbx = kb.take(np.arange(Bx)) if Bx != B else kb
        gx = Geometry.shared_t(Bx, N)
        for _ in range(2):
            solver.loglike(kbx, gx, t, y, logdet=ld[:Bx], quad=q[:Bx], status=st[:Bx], flags=fl)
        assert int(st[:Bx].abs().sum()) == 0
        res[label] = Bx * N / (solver.last_kernel_ms * 1e-3)
        res[label + "_ll"] = (ld[:3].cpu().numpy(), q[:3].cpu().numpy())
    rel = max(abs(res["narrow_ll"][i] / res["wide_ll"][i] - 1).max() for i in range(2))
    print(f"{name:18s} narrow {res['narrow']:.3e} steps/s   wide {res['wide']:.3e} steps/s   "
          f"x{res['narrow'] / res['wide']:.1f}   narrow vs wide rel {rel:.1e}")
