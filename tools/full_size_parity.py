#!/usr/bin/env python
"""Parity at BASELINE's full length: one solar light curve of 2^20 points through the fused GPU
kernels against the CPU oracle on the same inputs (the oracle needs ~25 s per pass on one core)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import batch, philox
import oracle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
t = np.arange(N) * 6e-5
scan = k.scan_coefficients()
t0 = time.time()
x, status = batch.sample([k], t, seed=5, subtract_mean=False)
nrm = philox.normals(5, 0, N)
x_ref, ld_ref, st_ref = oracle.stream(1, scan, t, nrm, fast=True)
print("sample: status", status[0], st_ref, "max rel", np.max(np.abs(x[0] - x_ref)) / np.max(np.abs(x_ref)),
      "(%.1f s)" % (time.time() - t0))
ll, logdet, quad, status = batch.log_likelihood([k], t, x_ref, return_parts=True)
o_ld, o_q, _ = oracle.stream(0, scan, t, x_ref, fast=True)
print("loglike: logdet rel", abs(logdet[0] - o_ld) / abs(o_ld), "quad rel", abs(quad[0] - o_q) / abs(o_q),
      "quad/N", quad[0] / N, "sum n^2 / N", np.sum(nrm * nrm) / N)
