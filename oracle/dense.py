"""
ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Dense ground truth straight from the kernel definition: build K and use
numpy.linalg.cholesky.  This is what pins the recurrences mathematically
(SURVEY.md section 8c): K = L L^T is unique, celerite's factor is L = L_c sqrt(D).
"""
import numpy as np

from . import terms_oracle as T


def covariance_semiseparable(scan_coeffs, t, diag=None):
    """Dense K exactly as the semiseparable generators define it: off-diagonal k'(tau) from
    the (already exposure-transformed) coefficients, diagonal sum(a') + ddiag + diag.  This is
    the matrix the recurrences factor, free of the separate FP64 cancellation that
    :func:`covariance` (the |tau| < delta closed form) carries (SURVEY.md finding 0.6)."""
    *coeffs, ddiag = scan_coeffs
    t = np.asarray(t, dtype=float)
    K = T.get_value(tuple(coeffs), np.abs(t[:, None] - t[None, :]))
    idx = np.arange(len(t))
    K[idx, idx] = np.sum(coeffs[0]) + np.sum(coeffs[2]) + ddiag
    if diag is not None:
        K = K + np.diag(np.broadcast_to(diag, t.shape))
    return K


def covariance(coeffs, t, delta=None, diag=None):
    """Dense K from the kernel definition k_delta(|t_i - t_j|) (A.1/A.4 get_value)."""
    t = np.asarray(t, dtype=float)
    tau = np.abs(t[:, None] - t[None, :])
    if delta:
        K = T.get_value_convolved(coeffs, delta, tau.ravel()).reshape(tau.shape)
    else:
        K = T.get_value(coeffs, tau)
    if diag is not None:
        K = K + np.diag(np.broadcast_to(diag, t.shape))
    return K


def log_likelihood(K, y):
    L = np.linalg.cholesky(K)
    z = np.linalg.solve(L, y)   # small N only
    return float(-0.5 * z @ z - np.sum(np.log(np.diag(L))) - 0.5 * len(y) * np.log(2 * np.pi))


def dot_tril(K, n):
    return np.linalg.cholesky(K) @ n


def apply_inverse(K, y):
    return np.linalg.solve(K, y)


def exposure_integral(coeffs, delta, tau, order=64):
    """k_delta(tau) = delta^-2 int_{-delta}^{delta} (delta - |x|) k(tau + x) dx by Gauss-Legendre
    on the two smooth halves -- the definition A.4's closed forms must match."""
    xg, wg = np.polynomial.legendre.leggauss(order)
    out = 0.0
    for lo, hi in ((-delta, 0.0), (0.0, delta)):
        # split again at the kink of k(|tau + x|) if it falls inside the half interval
        cuts = [lo, hi]
        if lo < -tau < hi:
            cuts = [lo, -tau, hi]
        for a, b in zip(cuts[:-1], cuts[1:]):
            x = 0.5 * (b - a) * xg + 0.5 * (b + a)
            out += 0.5 * (b - a) * np.sum(wg * (delta - np.abs(x)) * T.get_value(coeffs, tau + x))
    return out / delta ** 2
