import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gadfly_b200 as g
from gadfly_b200 import solver as S
from gadfly_b200.solver import Geometry, KernelBatch, Solver
solver = Solver(0)
solar = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
N, B = 16384, 148
kb = KernelBatch([solar] * B); geom = Geometry.shared_t(B, N)
dev = torch.device("cuda", 0)
t_d = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
y_d = torch.randn(B * N, dtype=torch.float64, device=dev) * 285.0
torch.cuda.synchronize()
solver.loglike(kb, geom, t_d, y_d, flags=S.FLAG_BLOCKED)
print("ms", solver.last_kernel_ms, "cycles per row", solver.last_kernel_ms * 1e-3 * 1.965e9 / N)
