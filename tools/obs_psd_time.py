#!/usr/bin/env python
"""Time of the observed power spectrum + log binning on the device (K7):
B light curves x N points resident in HBM -> power[B, N/2] -> 15 log bins.  usage: [B] [N]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gadfly_b200 import psd
from gadfly_b200.solver import Solver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
solver = Solver(0)
dev = torch.device("cuda", 0)
flux = torch.randn(B, N, dtype=torch.float64, device=dev) * 300.0
torch.cuda.synchronize()


def timed(fn, reps=5):
    fn(); solver.synchronize(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn(); solver.synchronize(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r


t_ps, (freq, power, norm) = timed(lambda: psd.power_spectra(flux, d_days=1 / 1440, solver=solver))
t_bin, (fb, pb, eb) = timed(lambda: psd.bin_power_spectra(freq, power, bins=15, solver=solver))
ref = np.abs(np.fft.rfft(flux[0].cpu().numpy())[1:]) ** 2 * (60e-6 / np.sqrt(2 * np.pi)) / N
got = power[0].cpu().numpy()
gb = B * N * 8 / 1e9
print(f"{B} light curves x {N} points ({gb:.2f} GB of flux): power spectra {t_ps * 1e3:.2f} ms "
      f"({gb / t_ps:.0f} GB/s of input; cuFFT D2Z + normalisation kernel), binning into 15 log bins {t_bin * 1e3:.2f} ms; "
      f"max rel vs numpy rfft (first light curve) {np.max(np.abs(got[:len(ref)] / ref[:len(got)] - 1)):.1e}")
