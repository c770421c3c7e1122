"""
Host-side kernel terms: the O(J) coefficient algebra that feeds the CUDA scan.

These classes mirror the part of ``celerite2.terms`` that gadfly touches
(reference call sites: gadfly/core.py:336 ``TermConvolution`` base class,
:371-373 ``SHOTerm(**hyperparameters)``, :379 ``TermSum(*terms)``, :394
``TermConvolution.__init__(term_sum, delta)``, :426 ``.term.terms``/``.delta``,
gadfly/psd.py:151 ``kernel.get_psd(omega)``).  celerite2 itself is an external,
un-vendored dependency of the reference; the formulas below are its published
algorithm (Foreman-Mackey et al. 2017; celerite2 ``terms.py``), see SURVEY.md
Appendix A.3-A.5.

Only the O(J) work lives here (plain numpy, FP64, a fixed expression order --
the exposure-time transform cancels catastrophically for the slowest
granulation term and must be evaluated identically wherever it is evaluated).
Everything O(N*J) or larger -- rows of U/V, the O(N J^2) scan, dense PSD
grids -- runs in the CUDA library (:mod:`gadfly_b200.solver`).
"""
import numpy as np

__all__ = ["Term", "SHOTerm", "RealTerm", "ComplexTerm", "TermSum", "TermConvolution"]

_EMPTY = np.empty(0, dtype=np.float64)


class Term:
    """Base class: a sum of real and complex exponential kernel components."""

    name = None

    def get_coefficients(self):
        """-> (ar, cr, ac, bc, cc, dc), each a float64 vector."""
        raise NotImplementedError

    def __add__(self, other):
        return TermSum(self, other)

    def __radd__(self, other):
        return TermSum(other, self)

    # -- flat views used by the CUDA library ---------------------------------
    def base_coefficients(self):
        """Coefficients of the *un-convolved* kernel (what the PSD uses)."""
        return tuple(np.ascontiguousarray(x, dtype=np.float64) for x in self.get_coefficients())

    def scan_coefficients(self):
        """(ar, cr, ac, bc, cc, dc, ddiag): what the semiseparable scan uses.
        ``ddiag`` is a constant added to the user diagonal (0 for plain terms)."""
        return self.base_coefficients() + (0.0,)

    @property
    def exposure(self):
        """Exposure time folded into this kernel (0 for plain terms), in 1/uHz."""
        return 0.0

    @property
    def J(self):
        ar, _, ac, _, _, _ = self.get_coefficients()
        return len(ar) + 2 * len(ac)

    def get_value(self, tau):
        """Covariance function k(tau) (celerite2 ``Term.get_value``).  Host numpy,
        one term at a time: used for the dense cross-covariances of the predictive variance."""
        ar, cr, ac, bc, cc, dc = self.get_coefficients()
        tau = np.abs(np.asarray(tau, dtype=np.float64))
        k = np.zeros_like(tau)
        for a, c in zip(ar, cr):
            k += a * np.exp(-c * tau)
        for a, b, c, d in zip(ac, bc, cc, dc):
            arg = d * tau
            k += np.exp(-c * tau) * (a * np.cos(arg) + b * np.sin(arg))
        return k

    def get_psd(self, omega):
        """Power spectral density at angular frequencies ``omega`` [rad uHz].

        Evaluated by the CUDA PSD kernel (reference call: gadfly/psd.py:151)."""
        from . import solver
        omega = np.asarray(omega, dtype=np.float64)
        flat = np.ascontiguousarray(omega.ravel())
        out = solver.psd([self], flat)[0]
        return out.reshape(omega.shape)


class RealTerm(Term):
    r"""k(tau) = a exp(-c tau)"""

    def __init__(self, *, a, c):
        self.a, self.c = float(a), float(c)

    def get_coefficients(self):
        return (np.array([self.a]), np.array([self.c]), _EMPTY, _EMPTY, _EMPTY, _EMPTY)


class ComplexTerm(Term):
    r"""k(tau) = exp(-c tau) (a cos(d tau) + b sin(d tau))"""

    def __init__(self, *, a, b, c, d):
        self.a, self.b, self.c, self.d = float(a), float(b), float(c), float(d)

    def get_coefficients(self):
        return (_EMPTY, _EMPTY, np.array([self.a]), np.array([self.b]),
                np.array([self.c]), np.array([self.d]))


class SHOTerm(Term):
    """Stochastically-driven damped simple harmonic oscillator.

    Parameterised by ``S0, w0, Q`` exactly as the reference passes them
    (gadfly/core.py:372); ``eps`` regularises the critically damped case."""

    def __init__(self, *, S0, w0, Q, eps=1e-5, name=None):
        self.S0, self.w0, self.Q, self.eps = float(S0), float(w0), float(Q), float(eps)
        if name is not None:
            self.name = name

    def get_coefficients(self):
        S0, w0, Q = self.S0, self.w0, self.Q
        if Q < 0.5:  # overdamped: two real terms
            f = np.sqrt(np.maximum(1.0 - 4.0 * Q ** 2, self.eps))
            ar = 0.5 * S0 * w0 * Q * np.array([1.0 + 1.0 / f, 1.0 - 1.0 / f])
            cr = 0.5 * w0 / Q * np.array([1.0 - f, 1.0 + f])
            return (ar, cr, _EMPTY, _EMPTY, _EMPTY, _EMPTY)
        f = np.sqrt(np.maximum(4.0 * Q ** 2 - 1.0, self.eps))
        a = S0 * w0 * Q
        c = 0.5 * w0 / Q
        return (_EMPTY, _EMPTY, np.array([a]), np.array([a / f]), np.array([c]), np.array([c * f]))

    def __repr__(self):
        return f"SHOTerm(S0={self.S0!r}, w0={self.w0!r}, Q={self.Q!r})"


class TermSum(Term):
    """Sum of terms; coefficients are concatenated in term order."""

    def __init__(self, *terms):
        flat = []
        for t in terms:
            if isinstance(t, TermSum):
                flat.extend(t.terms)
            else:
                flat.append(t)
        if any(isinstance(t, TermConvolution) for t in flat):
            raise TypeError("exposure-integrated terms cannot be summed; "
                            "sum the terms first and convolve the sum")
        self._terms = tuple(flat)

    @property
    def terms(self):
        return self._terms

    def get_coefficients(self):
        if not self._terms:
            return (_EMPTY,) * 6
        parts = [t.get_coefficients() for t in self._terms]
        return tuple(np.concatenate([np.atleast_1d(p[i]) for p in parts]) for i in range(6))


class TermConvolution(Term):
    """A term integrated over a box exposure of length ``delta`` (same unit as t).

    Valid as a semiseparable model when all |t_i - t_j| >= delta, i != j."""

    def __init__(self, term, delta):
        self.term = term
        self.delta = float(delta)

    @property
    def exposure(self):
        return self.delta

    def base_coefficients(self):
        return self.term.base_coefficients()

    def get_coefficients(self):
        ar, cr, a, b, c, d = self.term.get_coefficients()
        dt = self.delta
        # real components
        crd = cr * dt
        ar_new = 2 * ar * (np.cosh(crd) - 1) / crd ** 2
        # complex components
        cd = c * dt
        dd = d * dt
        c2 = c ** 2
        d2 = d ** 2
        factor = 2.0 / (dt * (c2 + d2)) ** 2
        cos_term = np.cosh(cd) * np.cos(dd) - 1
        sin_term = np.sinh(cd) * np.sin(dd)
        C1 = a * (c2 - d2) + 2 * b * c * d
        C2 = b * (c2 - d2) - 2 * a * c * d
        return (ar_new, cr,
                factor * (C1 * cos_term - C2 * sin_term),
                factor * (C2 * cos_term + C1 * sin_term),
                c, d)

    def get_value(self, tau):
        """Exposure-integrated covariance k_delta(tau) = delta^-2 int (delta - |x|) k(tau + x) dx
        (celerite2 ``TermConvolution.get_value``): the semiseparable form with the
        transformed coefficients for |tau| >= delta, the closed form of the overlapping-exposure
        case below."""
        ar, cr, ac, bc, cc, dc = self.term.get_coefficients()
        dt = self.delta
        tau = np.abs(np.asarray(tau, dtype=np.float64))
        small = tau < dt
        ts = tau[small]                     # overlapping exposures (usually only tau = 0)
        dmt, dpt = dt - ts, dt + ts
        k = np.zeros_like(tau)              # |tau| >= delta form, everywhere
        ks = np.zeros_like(ts)
        for a, c in zip(ar, cr):
            cd = c * dt
            norm = 2 * a / cd ** 2
            k += norm * (np.cosh(cd) - 1) * np.exp(-c * tau)
            ks += norm * (np.cosh(cd) - 1) * np.exp(-c * ts) + norm * (c * dmt - np.sinh(c * dmt))
        for a, b, c, d in zip(ac, bc, cc, dc):
            cd, dd = c * dt, d * dt
            c2pd2 = c * c + d * d
            C1 = a * (c * c - d * d) + 2 * b * c * d
            C2 = b * (c * c - d * d) - 2 * a * c * d
            norm = 1.0 / (dt * c2pd2) ** 2
            cos_term = 2 * (np.cosh(cd) * np.cos(dd) - 1)
            sin_term = 2 * (np.sinh(cd) * np.sin(dd))
            k += ((C1 * cos_term - C2 * sin_term) * np.cos(d * tau)
                  + (C2 * cos_term + C1 * sin_term) * np.sin(d * tau)) * (np.exp(-c * tau) * norm)
            if ts.size:
                k0 = np.exp(-c * ts)
                em, ep = np.exp(-c * dmt), np.exp(-c * dpt)
                ks += (2 * (a * c + b * d) * c2pd2 * dmt
                       + C1 * (em * np.cos(d * dmt) + ep * np.cos(d * dpt) - 2 * k0 * np.cos(d * ts))
                       + C2 * (em * np.sin(d * dmt) + ep * np.sin(d * dpt) - 2 * k0 * np.sin(d * ts))) * norm
        k[small] = ks
        return k

    def diagonal_correction(self):
        """k_delta(0) - sum(a'): added to the user diagonal before the scan."""
        ar, cr, a, b, c, d = self.term.get_coefficients()
        dt = self.delta
        cd = cr * dt
        delta_diag = 2 * np.sum(ar * (cd - np.sinh(cd)) / cd ** 2)
        cd = c * dt
        dd = d * dt
        c2 = c ** 2
        d2 = d ** 2
        c2pd2 = c2 + d2
        C1 = a * (c2 - d2) + 2 * b * c * d
        C2 = b * (c2 - d2) - 2 * a * c * d
        norm = (dt * c2pd2) ** 2
        sinh = np.sinh(cd)
        cosh = np.cosh(cd)
        delta_diag += 2 * np.sum(
            (C2 * cosh * np.sin(dd) - C1 * sinh * np.cos(dd) + (a * c + b * d) * dt * c2pd2) / norm
        )
        return float(delta_diag)

    def scan_coefficients(self):
        coeffs = tuple(np.ascontiguousarray(x, dtype=np.float64) for x in self.get_coefficients())
        return coeffs + (self.diagonal_correction(),)
