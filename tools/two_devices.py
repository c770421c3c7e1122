#!/usr/bin/env python
"""Two solver handles on two devices in ONE process give identical numbers (the > 48 KB shared-memory
opt-in of the scan and sweep kernels is per device).  usage (2 GPUs): python tools/two_devices.py"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import batch
from gadfly_b200.solver import Solver
k = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
N = 700
t = np.arange(N) * 6e-5
y = np.random.default_rng(0).standard_normal((3, N)) * 200
res = []
for dev in (0, 1, 0):
    s = Solver(dev)
    ll = batch.log_likelihood([k] * 3, t, y, solver=s)
    gp = g.GaussianProcess(k, t=t, solver=s)
    res.append((ll, gp.log_likelihood(y[0]), gp.apply_inverse(y[1])[:3]))
    print("device", dev, ll, res[-1][1])
assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2])
print("two devices in one process: identical")
