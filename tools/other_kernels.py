#!/usr/bin/env python
"""One launch of each round-2 kernel that is not the fused scan (for an ncu capture):
device feeder (hyper / coef / coef_csr / bandpass), kernel PSD, observed PSD + binning, k right-hand
sides on one factor (prep / sweep / quad).  usage: python tools/other_kernels.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gadfly_b200 as g
from gadfly_b200 import batch, feeder, psd, scale, workloads
from gadfly_b200.solver import Solver

solver = Solver(0)
dev = torch.device("cuda", 0)
M, R, T, L = workloads.kepler_like_stars(4096, 1)
kb = feeder.kernel_batch_for_stars_device(solver, M, R, T, L, texp_s=60.0)
workloads.lattice_batch(4096, 3, solver=solver)


class Band:
    wavelength = np.linspace(0.4, 0.9, 200)
    transmittance = np.exp(-0.5 * ((np.linspace(0.4, 0.9, 200) - 0.65) / 0.1) ** 2)


wl, tr = scale.bandpass_grid(g.Filter(Band))
solver.bandpass_amplitude(T, wl, tr)
omega = torch.as_tensor(2 * np.pi * np.linspace(0.01, 8333.0, 1000000), device=dev)
out = torch.empty(256 * 1000000, dtype=torch.float64, device=dev)
solver.psd(kb.take(np.arange(256)), omega, out=out)
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
N = 1 << 17
t = np.arange(N) * 6e-5
x, st = batch.sample([kernel] * 16, t, size=8, seed=1, solver=solver)             # prep + factor + sweep
flux = torch.as_tensor(x.reshape(128, N), device=dev)
freq, power, norm = psd.power_spectra(flux, d_days=1 / 1440, solver=solver)
psd.bin_power_spectra(freq, power, bins=15, solver=solver)
solver.synchronize()
print("ok", kb.B, x.shape)
