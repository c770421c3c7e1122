#!/usr/bin/env python
"""Latency of the reference-shaped single-light-curve API (gadfly_b200.GaussianProcess) on one GPU:
compute / log_likelihood / sample / apply_inverse / predict for the solar kernel.
usage: python tools/api_latency.py [N ...]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g


def timed(f, *a, **k):
    t0 = time.perf_counter()
    r = f(*a, **k)
    return r, time.perf_counter() - t0


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [10000, 100000, 1000000]
    kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
    g.GaussianProcess(kernel, t=np.arange(100) * 6e-5).log_likelihood(np.zeros(100))   # warm-up
    for N in sizes:
        t = np.arange(N) * 6e-5
        gp, t_compute = timed(g.GaussianProcess, kernel, t=t)
        np.random.seed(42)
        y, t_sample = timed(gp.sample)
        ll, t_ll = timed(gp.log_likelihood, y)
        _, t_inv = timed(gp.apply_inverse, y)
        line = dict(N=N, compute_s=t_compute, sample_s=t_sample, log_likelihood_s=t_ll, apply_inverse_s=t_inv,
                    logL=float(ll))
        if N <= 200000:
            tp = t[::50] + 3e-5
            _, t_pred = timed(gp.predict, y, t=tp)
            line["predict_s"] = t_pred
        print(line, flush=True)


if __name__ == "__main__":
    main()
