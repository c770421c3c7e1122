#!/usr/bin/env python
"""Ground truth from the kernel DEFINITION in extended precision -> tests/golden/def_*.npz.

TEST INFRASTRUCTURE.  Imports neither ``gadfly_b200`` nor ``oracle`` (nor their closed forms):

  inputs   (S0, w0, Q) per term from tests/golden/for_star.json (itself an independent mpmath
           evaluation of the scaling relations, tools/make_forstar_fixture.py)
  kernel   k(tau) = sum_j S0 w0 Q exp(-w0 tau / 2Q) [cos(eta w0 tau) + sin(eta w0 tau) / (2 eta Q)],
           eta = sqrt(1 - 1/(4 Q^2))      -- the SHO kernel whose PSD is the reference's
           ``_sho_psd`` (gadfly/core.py:33-41; Foreman-Mackey et al. 2017, Eq. 23)
  exposure k_D(tau) = D^-2 int_{-D}^{D} (D - |x|) k(|tau + x|) dx   (what TermConvolution stands
           for) by Gauss-Legendre quadrature (40 nodes per smooth piece, nodes from
           mpmath) in numpy longdouble (64-bit mantissa) -- NOT by the closed form
  algebra  dense K = k_D(|t_i - t_j|) + diag, Cholesky / triangular solves in longdouble

so that SHOTerm.get_coefficients (A.3), the exposure transform and its diagonal term (A.4), the
row generation (A.2) and the recurrences (A.6) are pinned END TO END by numbers none of them
produced.  Time stamps are integer multiples of a power of two, so every lag is exact in FP64.

Exposure times.  celerite2's FP64 closed form for the exposure-integrated coefficients is only
well conditioned when 1e-3 <~ |c + i d| D <~ 5 for EVERY term: below, cosh(c D) cos(d D) - 1 cancels
(relative error eps / (|c + i d| D)^2 on a'); above, sum a' and the diagonal
correction are each ~exp(c D) k(0) and cancel (30-min exposure of the solar kernel: 7.5e12 - 7.5e12
= 6e4).  An FP64 implementation that follows celerite2 inherits both, so the tight cases use exposures
inside that window for their kernel, and one case states the limit at the reference's default 1 min.

Cases
  def_solar_200s   Sun (86 terms, J = 172), 200 s exposure, 244 s cadence, N = 1024
  def_subgiant     1.25 Msun / 2.1 Rsun subgiant (J = 172), 400 s exposure, gaps, heteroscedastic errors, N = 2048
  def_giant        KIC 9333184 (62 terms, J = 124; granulation 100x slower than the Sun's), exposure
                   0.02 uHz^-1 on a 2^-5 uHz^-1 cadence, N = 1024
  def_solar_sc     Sun, 1-min exposure and cadence (the reference's default), N = 1024: the slowest
                   granulation term (c D = 7e-5) carries ~6e-9 relative FP64 cancellation error in a', so
                   no FP64 implementation of celerite2's formulas agrees with the definition better
                   than ~3e-9 on samples; the fixture quantifies exactly that.
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(HERE, "tests", "golden")
LD = np.longdouble


def gauss_legendre(n):
    mp.mp.dps = 40
    x, w = [], []
    # explicit n-point rule on [-1, 1]: roots of P_n by Newton in 40 digits
    for k in range(1, n + 1):
        z = mp.cos(mp.pi * (k - mp.mpf("0.25")) / (n + mp.mpf("0.5")))
        for _ in range(100):
            p0, p1 = mp.mpf(1), z
            for j in range(2, n + 1):
                p0, p1 = p1, ((2 * j - 1) * z * p1 - (j - 1) * p0) / j
            dp = n * (z * p1 - p0) / (z * z - 1)
            dz = p1 / dp
            z -= dz
            if abs(dz) < mp.mpf(10) ** -38:
                break
        x.append(z)
        w.append(2 / ((1 - z * z) * dp * dp))
    return (np.array([LD(mp.nstr(v, 25)) for v in x]), np.array([LD(mp.nstr(v, 25)) for v in w]))


GL_X, GL_W = None, None


def sho_kernel(tau, S0, w0, Q):
    """k(tau) summed over the terms; tau >= 0, longdouble arrays broadcast against [J]."""
    eta = np.sqrt(1 - 1 / (4 * Q * Q))
    a = S0 * w0 * Q
    arg = eta * w0 * tau
    return a * np.exp(-w0 * tau / (2 * Q)) * (np.cos(arg) + np.sin(arg) / (2 * eta * Q))


def exposure_kernel(lags, S0, w0, Q, D):
    """k_D at the given lags [M] (longdouble), summed over terms."""
    global GL_X, GL_W
    if GL_X is None:
        GL_X, GL_W = gauss_legendre(40)
    lags = np.asarray(lags, dtype=LD)
    out = np.zeros(len(lags), dtype=LD)
    S0, w0, Q = (np.asarray(v, dtype=LD)[None, :] for v in (S0, w0, Q))
    D = LD(D)
    for m, tau in enumerate(lags):
        # smooth pieces of the integrand on [-D, D]: split at x = 0 (weight) and x = -tau (|tau + x|)
        cuts = sorted({-D, LD(0), D} | ({-tau} if -D < -tau < D else set()))
        total = LD(0)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            x = (hi + lo) / 2 + (hi - lo) / 2 * GL_X
            wts = (hi - lo) / 2 * GL_W
            k = sho_kernel(np.abs(tau + x)[:, None], S0, w0, Q).sum(axis=1)
            total += np.sum(wts * (D - np.abs(x)) * k)
        out[m] = total / (D * D)
    return out


def cholesky_ld(K):
    """Lower Cholesky factor in longdouble (column version, numpy vector operations)."""
    n = K.shape[0]
    L = np.zeros_like(K)
    for j in range(n):
        v = K[j:, j] - L[j:, :j] @ L[j, :j]
        if not v[0] > 0:
            raise np.linalg.LinAlgError(f"pivot {j} not positive")
        L[j:, j] = v / np.sqrt(v[0])
    return L


def solve_lower_ld(L, b):
    n = len(b)
    z = np.zeros(n, dtype=LD)
    for i in range(n):
        z[i] = (b[i] - L[i, :i] @ z[:i]) / L[i, i]
    return z


def solve_upper_ld(L, b):
    n = len(b)
    z = np.zeros(n, dtype=LD)
    for i in range(n - 1, -1, -1):
        z[i] = (b[i] - L[i + 1:, i] @ z[i + 1:]) / L[i, i]
    return z


def make_case(name, star, delta, step_log2, index, yerr, seed):
    """index: integer time stamps (multiples of 2^step_log2 uHz^-1)."""
    with open(os.path.join(OUT, "for_star.json")) as fh:
        st = json.load(fh)["stars"][star]
    S0, w0, Q = (np.array(st[k], dtype=np.float64) for k in ("S0", "w0", "Q"))
    index = np.asarray(index, dtype=np.int64)
    n = len(index)
    dt = 2.0 ** step_log2
    t = index.astype(np.float64) * dt          # exact
    max_lag = int(index[-1] - index[0])
    table = exposure_kernel(np.arange(max_lag + 1).astype(LD) * LD(dt), S0, w0, Q, delta)
    lag = np.abs(index[:, None] - index[None, :])
    K = table[lag]
    rng = np.random.default_rng(seed)
    diag = np.zeros(n) if yerr is None else (yerr * (0.5 + rng.random(n))) ** 2
    K[np.arange(n), np.arange(n)] += diag.astype(LD)
    L = cholesky_ld(K)
    normals = rng.standard_normal(n)
    x = L @ normals.astype(LD)                                  # = L_c sqrt(D) n of celerite (unique factor)
    y = np.asarray(L @ rng.standard_normal(n).astype(LD), dtype=np.float64)   # data: a draw, rounded to FP64
    z = solve_lower_ld(L, y.astype(LD))
    quad = z @ z
    logdet = 2 * np.sum(np.log(np.diag(L)))
    alpha = solve_upper_ld(L, z)
    logl = -(quad + logdet + n * np.log(2 * LD(np.pi))) / 2
    np.savez(os.path.join(OUT, name + ".npz"), S0=S0, w0=w0, Q=Q, delta=delta, t=t, diag=diag,
             normals=normals, y=y, x=np.asarray(x, dtype=np.float64), alpha=np.asarray(alpha, dtype=np.float64),
             logdet=float(logdet), quad=float(quad), logl=float(logl), k0=float(table[0]),
             cond_estimate=float(np.max(np.diag(L)) ** 2 / np.min(np.diag(L)) ** 2))
    print(f"{name}: N = {n}, J = {2 * len(S0)}, k(0) = {float(table[0]):.6g}, logL = {float(logl):.12g}, "
          f"min pivot {float(np.min(np.diag(L)) ** 2):.4g}")


def logl_definition(S0, w0, Q, delta, index, step_log2, diag, y):
    """log L from the definition in longdouble (inputs may be longdouble arrays)."""
    n = len(index)
    dt = LD(2.0) ** step_log2
    max_lag = int(index[-1] - index[0])
    table = exposure_kernel(np.arange(max_lag + 1).astype(LD) * dt, S0, w0, Q, delta)
    K = table[np.abs(index[:, None] - index[None, :])]
    K[np.arange(n), np.arange(n)] += np.asarray(diag, dtype=LD)
    L = cholesky_ld(K)
    z = solve_lower_ld(L, np.asarray(y, dtype=LD))
    return -(z @ z + 2 * np.sum(np.log(np.diag(L))) + n * np.log(2 * LD(np.pi))) / 2


def make_gradient_case(name, seed):
    """d log L / d ln(S0_j, w0_j, Q_j) of a 6-term kernel (3 granulation-like, 3 p-mode-like terms) by
    central differences of the longdouble definition (h = 1e-7 in ln p: truncation (h Q)^2 / 6 < 1e-9 even
    for ln w0 of the Q = 650 term, rounding ~1e-19 |log L| / h ~ 1e-9 absolute)."""
    rng = np.random.default_rng(seed)
    S0 = np.array([900.0, 12.0, 0.6, 2.0e-3, 3.5e-3, 1.2e-3])
    w0 = np.array([7.0, 120.0, 900.0, 17000.0, 19500.0, 21500.0])
    Q = np.array([0.6, 0.6, 0.7, 400.0, 650.0, 300.0])
    delta, step_log2, n = 2e-4, -12, 256
    index = np.arange(n)
    t = index * 2.0 ** step_log2
    diag = np.full(n, 4.0)
    # data: a draw from the model
    table = exposure_kernel(np.arange(n).astype(LD) * LD(2.0) ** step_log2, S0, w0, Q, delta)
    K = table[np.abs(index[:, None] - index[None, :])]
    K[np.arange(n), np.arange(n)] += diag.astype(LD)
    y = np.asarray(cholesky_ld(K) @ rng.standard_normal(n).astype(LD), dtype=np.float64)
    h = LD(1e-7)
    base = [np.asarray(v, dtype=LD) for v in (S0, w0, Q)]
    grad = np.zeros((3, len(S0)))
    for i in range(3):
        for j in range(len(S0)):
            vals = []
            for sgn in (-1, 1):
                p = [v.copy() for v in base]
                p[i][j] = p[i][j] * np.exp(sgn * h)
                vals.append(logl_definition(p[0], p[1], p[2], delta, index, step_log2, diag, y))
            grad[i, j] = float((vals[1] - vals[0]) / (2 * h))
    logl = float(logl_definition(base[0], base[1], base[2], delta, index, step_log2, diag, y))
    np.savez(os.path.join(OUT, name + ".npz"), S0=S0, w0=w0, Q=Q, delta=delta, t=t, diag=diag, y=y,
             logl=logl, grad=grad)
    print(f"{name}: log L = {logl:.10g}, |grad| = {np.linalg.norm(grad):.6g}")
    print(np.array2string(grad, precision=6))


def main():
    which = sys.argv[1:] or ["def_solar_200s", "def_subgiant", "def_giant", "def_solar_sc", "def_grad"]
    if "def_grad" in which:
        make_gradient_case("def_grad", 21)
    if "def_solar_200s" in which:
        make_case("def_solar_200s", "Sun", 2e-4, -12, np.arange(1024), None, 11)
    if "def_subgiant" in which:
        rng = np.random.default_rng(5)
        idx = np.cumsum(1 + (rng.random(2048) < 0.03) * rng.integers(1, 40, 2048))     # gaps
        make_case("def_subgiant", "subgiant", 4e-4, -11, idx, 30.0, 12)
    if "def_giant" in which:
        make_case("def_giant", "KIC 9333184", 0.02, -5, np.arange(1024), 200.0, 14)
    if "def_solar_sc" in which:
        # 1-min exposure (6e-5) on a 2^-14 uHz^-1 = 61.04 s cadence
        make_case("def_solar_sc", "Sun", 6e-5, -14, np.arange(1024), None, 13)


if __name__ == "__main__":
    main()
