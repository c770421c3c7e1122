"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/__init__.py).

numpy restatement of the celerite2 term algebra gadfly relies on
(reference call sites gadfly/core.py:371-373,379,394,
gadfly/psd.py:151; closed-form SHO PSD gadfly/core.py:33-41).

Functions take and return plain arrays; "coeffs" is always the 6-tuple
(ar, cr, ac, bc, cc, dc).
"""
import numpy as np

E = np.empty(0)


def sho_coefficients(S0, w0, Q, eps=1e-5):
    """A.3: SHOTerm(S0, w0, Q) -> coeffs."""
    if Q < 0.5:
        f = np.sqrt(np.maximum(1.0 - 4.0 * Q ** 2, eps))
        return (0.5 * S0 * w0 * Q * np.array([1.0 + 1.0 / f, 1.0 - 1.0 / f]),
                0.5 * w0 / Q * np.array([1.0 - f, 1.0 + f]), E, E, E, E)
    f = np.sqrt(np.maximum(4.0 * Q ** 2 - 1.0, eps))
    a = S0 * w0 * Q
    c = 0.5 * w0 / Q
    return (E, E, np.array([a]), np.array([a / f]), np.array([c]), np.array([c * f]))


def sum_coefficients(list_of_coeffs):
    """A.1: TermSum concatenates in term order."""
    return tuple(np.concatenate([np.atleast_1d(c[i]) for c in list_of_coeffs]) for i in range(6))


def sho_sum(hyper):
    """coeffs of a sum of SHO terms given [(S0, w0, Q), ...]."""
    return sum_coefficients([sho_coefficients(*h) for h in hyper])


def convolve_coefficients(coeffs, delta):
    """A.4: TermConvolution.get_coefficients (exposure-time integration)."""
    ar, cr, a, b, c, d = coeffs
    crd = cr * delta
    real = 2 * ar * (np.cosh(crd) - 1) / crd ** 2
    cd = c * delta
    dd = d * delta
    c2 = c ** 2
    d2 = d ** 2
    factor = 2.0 / (delta * (c2 + d2)) ** 2
    cos_term = np.cosh(cd) * np.cos(dd) - 1
    sin_term = np.sinh(cd) * np.sin(dd)
    C1 = a * (c2 - d2) + 2 * b * c * d
    C2 = b * (c2 - d2) - 2 * a * c * d
    return (real, cr, factor * (C1 * cos_term - C2 * sin_term),
            factor * (C2 * cos_term + C1 * sin_term), c, d)


def convolve_diagonal(coeffs, delta):
    """A.4: the diagonal correction k_delta(0) - sum(a')."""
    ar, cr, a, b, c, d = coeffs
    cd = cr * delta
    out = 2 * np.sum(ar * (cd - np.sinh(cd)) / cd ** 2)
    cd = c * delta
    dd = d * delta
    c2 = c ** 2
    d2 = d ** 2
    c2pd2 = c2 + d2
    C1 = a * (c2 - d2) + 2 * b * c * d
    C2 = b * (c2 - d2) - 2 * a * c * d
    norm = (delta * c2pd2) ** 2
    sinh = np.sinh(cd)
    cosh = np.cosh(cd)
    out += 2 * np.sum(
        (C2 * cosh * np.sin(dd) - C1 * sinh * np.cos(dd) + (a * c + b * d) * delta * c2pd2) / norm)
    return float(out)


def scan_coefficients(coeffs, delta=None):
    """7-tuple the scan consumes: convolved coeffs + ddiag (or the plain coeffs, 0)."""
    if delta is None or delta == 0:
        return tuple(coeffs) + (0.0,)
    return convolve_coefficients(coeffs, delta) + (convolve_diagonal(coeffs, delta),)


def get_value(coeffs, tau):
    """A.1: k(tau) of the plain term set."""
    ar, cr, ac, bc, cc, dc = coeffs
    tau = np.abs(np.asarray(tau, dtype=float))[..., None]
    k = np.sum(ar * np.exp(-cr * tau), axis=-1)
    arg = dc * tau
    return k + np.sum(np.exp(-cc * tau) * (ac * np.cos(arg) + bc * np.sin(arg)), axis=-1)


def get_value_convolved(coeffs, delta, tau):
    """A.4: k_delta(tau), both the |tau| >= delta and |tau| < delta branches."""
    ar, cr, a, b, c, d = coeffs
    tau = np.abs(np.atleast_1d(np.asarray(tau, dtype=float)))
    tau = tau[..., None]
    dt = delta
    # real part
    crd = cr * dt
    cosh = np.cosh(crd)
    norm = 2 * ar / crd ** 2
    K_large = np.sum(norm * (cosh - 1) * np.exp(-cr * tau), axis=-1)
    crdmt = np.maximum(crd - cr * tau, 0.0)
    K_small = K_large + np.sum(norm * (crdmt - np.sinh(crdmt)), axis=-1)
    # complex part
    cd = c * dt
    dd = d * dt
    c2 = c ** 2
    d2 = d ** 2
    c2pd2 = c2 + d2
    C1 = a * (c2 - d2) + 2 * b * c * d
    C2 = b * (c2 - d2) - 2 * a * c * d
    norm = 1.0 / (dt * c2pd2) ** 2
    k0 = np.exp(-c * tau)
    cdt = np.cos(d * tau)
    sdt = np.sin(d * tau)
    cos_term = 2 * (np.cosh(cd) * np.cos(dd) - 1)
    sin_term = 2 * (np.sinh(cd) * np.sin(dd))
    factor = k0 * norm
    K_large = K_large + np.sum((C1 * cos_term - C2 * sin_term) * factor * cdt, axis=-1)
    K_large = K_large + np.sum((C2 * cos_term + C1 * sin_term) * factor * sdt, axis=-1)
    # (the overlapping-exposure form is only used for tau < delta: clip, so that exp(c (tau - delta))
    # cannot overflow where it is discarded anyway)
    dmt = np.maximum(dt - tau, 0.0)
    dpt = dt + tau
    ec_m, ec_p = np.exp(-c * dmt), np.exp(-c * dpt)
    K_small = K_small + np.sum(2 * (a * c + b * d) * c2pd2 * dmt * norm, axis=-1)
    K_small = K_small + np.sum(
        (C1 * (ec_m * np.cos(d * dmt) + ec_p * np.cos(d * dpt) - 2 * k0 * cdt)
         + C2 * (ec_m * np.sin(d * dmt) + ec_p * np.sin(d * dpt) - 2 * k0 * sdt)) * norm, axis=-1)
    # K_small started from the *real* K_large only and its complex part is self-contained (A.4)
    return np.where(tau[..., 0] >= dt, K_large, K_small)


def psd(coeffs, omega):
    """A.5: Term.get_psd on plain coefficients."""
    ar, cr, ac, bc, cc, dc = coeffs
    w2 = np.asarray(omega, dtype=float)[..., None] ** 2
    out = np.sum(ar * cr / (cr ** 2 + w2), axis=-1)
    c2, d2 = cc ** 2, dc ** 2
    w02 = c2 + d2
    acc, bdc = ac * cc, bc * dc
    out = out + np.sum(((acc + bdc) * w02 + (acc - bdc) * w2)
                       / (w2 * w2 + 2.0 * (c2 - d2) * w2 + w02 * w02), axis=-1)
    return np.sqrt(2.0 / np.pi) * out


def psd_convolved(coeffs, delta, omega):
    """A.5: TermConvolution.get_psd = base psd * sinc^2(delta omega / 2)."""
    omega = np.asarray(omega, dtype=float)
    arg = 0.5 * delta * omega
    sinc = np.ones_like(arg)
    m = np.abs(arg) > 0
    sinc[m] = np.sin(arg[m]) / arg[m]
    return psd(coeffs, omega) * sinc ** 2


def sho_psd(omega, S0, w0, Q):
    """Closed-form SHO PSD exactly as the reference writes it (gadfly/core.py:33-41)."""
    return (np.sqrt(2 / np.pi) * S0 * w0 ** 4 /
            ((omega ** 2 - w0 ** 2) ** 2 + (omega ** 2 * w0 ** 2 / Q ** 2)))
