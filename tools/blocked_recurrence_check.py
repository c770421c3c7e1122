#!/usr/bin/env python
"""Numerical check of the k-step blocked recurrence proposed in DESIGN.md section 5b as the next scan
kernel: per block of k steps, k matrix-vector products against the same state, ONE batch of dot
products (Gram entries), a scalar k x k recursion, a rank-k update.  Plain numpy FP64 against the
oracle's step-by-step factor / solve_lower (oracle/celerite_oracle.c) on the solar kernel: d, W, z,
log det and the quadratic form for k = 1, 2, 4, 8 (CPU only; test infrastructure).
usage: python tools/blocked_recurrence_check.py [n_points]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g
import oracle
from oracle import terms_oracle as T

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
rng = np.random.default_rng(1)
t = np.cumsum(rng.choice([6e-5, 6e-5, 6e-5, 1.2e-4, 3e-3], N))          # cadence with gaps
coeffs = T.scan_coefficients(kernel.base_coefficients(), kernel.exposure)
c, a, U, V = oracle.celerite_matrices(coeffs[:6], t, diag=np.full(N, 30.0 ** 2), ddiag=coeffs[6])
y = rng.standard_normal(N) * 300.0
d_ref, W_ref = oracle.factor(t, c, a, U, V)
z_ref = oracle.solve_lower(t, c, U, W_ref, y)
J = U.shape[1]


def blocked(k):
    """S, F in the frame of the block's first step; rows moved into that frame (lazy decay)."""
    S = np.zeros((J, J))
    F = np.zeros(J)
    d = np.empty(N)
    W = np.empty((N, J))
    z = np.empty(N)
    for n0 in range(0, N, k):
        kk = min(k, N - n0)
        q = np.exp(-c[None, :] * (t[n0:n0 + kk] - t[n0])[:, None])       # decay since the block start
        Ut, Vt = U[n0:n0 + kk] * q, V[n0:n0 + kk] / q                    # u~ = u q, v~ = v / q
        G = Ut @ S                                                       # kk products, same S
        T0 = Vt - G
        M = Ut @ T0.T                                                    # Gram entries (m < i used)
        qf = np.einsum('ij,ij->i', G, Ut)
        f = Ut @ F
        cc = np.zeros((kk, kk))
        tt = np.empty((kk, J))
        dd = np.empty(kk)
        zz = np.empty(kk)
        for i in range(kk):                                              # scalar recursion
            for m in range(i):
                cc[i, m] = M[i, m] - sum(cc[m, l] / dd[l] * cc[i, l] for l in range(m))
            dd[i] = a[n0 + i] - qf[i] - sum(cc[i, m] ** 2 / dd[m] for m in range(i))
            zz[i] = y[n0 + i] - f[i] - sum(cc[i, m] / dd[m] * zz[m] for m in range(i))
            tt[i] = T0[i] - sum(cc[i, m] / dd[m] * tt[m] for m in range(i))
        ww = tt / dd[:, None]
        d[n0:n0 + kk], z[n0:n0 + kk] = dd, zz
        W[n0:n0 + kk] = ww * q                                           # W in its own step's frame, as stored
        S = S + tt.T @ ww                                                # rank-k update
        F = F + ww.T @ zz
        if n0 + kk < N:                                                  # into the next block's frame
            r = np.exp(-c * (t[n0 + kk] - t[n0]))
            S = S * np.outer(r, r)
            F = F * r
    return d, W, z


print(f"solar kernel J = {J}, N = {N}, cadence with gaps; oracle: log det {np.sum(np.log(d_ref)):.6f}, "
      f"quad {np.sum(z_ref ** 2 / d_ref):.6f}")
for k in (1, 2, 4, 8):
    d, W, z = blocked(k)
    print(f"k = {k}: max rel d {np.max(np.abs(d / d_ref - 1)):.2e}, W (scale of the row) "
          f"{np.max(np.abs(W - W_ref) / np.max(np.abs(W_ref), axis=1, keepdims=True)):.2e}, "
          f"z {np.max(np.abs(z - z_ref)) / np.max(np.abs(z_ref)):.2e}, log det "
          f"{abs(np.sum(np.log(d)) / np.sum(np.log(d_ref)) - 1):.2e}, quad "
          f"{abs(np.sum(z ** 2 / d) / np.sum(z_ref ** 2 / d_ref) - 1):.2e}")
