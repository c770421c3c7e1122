"""
``GaussianProcess``: gadfly's unit-aware GP interface (reference gadfly/gp.py:13-395),
with the celerite2 solver underneath replaced by the gadfly_b200 CUDA library.

Method names, keyword arguments and error behaviour follow the reference class and
the ``celerite2.GaussianProcess`` it subclasses (reference gadfly/gp.py:59,202-204,
232-234,327,350,370,391):

* ``ValueError`` for unsorted or non-1D ``t``, for both ``yerr`` and ``diag``, and for
  shape mismatches; ``RuntimeError`` when used before ``compute``;
  :class:`~gadfly_b200.solver.LinAlgError` when the matrix is not positive definite
  (silenced by ``quiet=True``: ``log_likelihood`` is then ``-inf``).
* time is converted to 1/uHz and flux to ppm at this boundary (reference
  gadfly/gp.py:61-126); ``sample`` subtracts the sample mean (reference gadfly/gp.py:392).

``compute`` runs the factor kernel (K3) and keeps ``d`` and ``W`` on the device;
``log_likelihood`` / ``dot_tril`` / ``apply_inverse`` / ``sample`` are O(N J) sweeps (K4).
Batches of independent light curves go through :mod:`gadfly_b200.batch` instead, which
uses the fused kernels and materialises nothing.
"""
import numpy as np

from . import units as u
from .solver import Geometry, KernelBatch, LinAlgError, default_solver
from .units import Quantity, to_value

__all__ = ['GaussianProcess', 'ConditionalDistribution']


class ConstantMean:
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, x):
        return self.value


class GaussianProcess:
    """
    The ``gadfly`` interface to the semiseparable Gaussian Process solver.
    """

    def __init__(self, kernel, t=None, mean=0.0, light_curve=None, **kwargs):
        self._original_flux_median = None
        self.kernel = kernel
        self.mean = mean

        # placeholders filled by compute
        self._t = None
        self._size = None
        self._diag = None
        self._mean_value = None
        self._log_det = -np.inf
        self._norm = np.inf
        self._d = None          # host copy of the pivots
        self._W = None          # device tensor [N * J], blocked column order
        self._kb = None
        self._solver = kwargs.pop('solver', None)

        if t is not None:
            t = self._time_to_freq(t)

        if light_curve is not None:
            t = self._time_to_freq(light_curve.time)
            flux = light_curve.flux
            if hasattr(flux, 'unmasked'):
                median_flux = np.nanmedian(flux.unmasked)
            else:
                median_flux = np.median(flux)
            self._original_flux_median = median_flux
            kwargs['yerr'] = self._flux_to_ppm(light_curve.flux_err, is_error=True)

        if t is not None:
            self.compute(t, **kwargs)

    # ---- mean (celerite2 semantics: scalar or callable) -------------------------------
    @property
    def mean(self):
        return self._mean

    @mean.setter
    def mean(self, mean):
        self._mean = mean if callable(mean) else ConstantMean(mean)

    @property
    def mean_value(self):
        if self._mean_value is None:
            raise RuntimeError("'compute' must be called before accessing mean_value")
        return self._mean_value

    # ---- unit handling (reference gadfly/gp.py:61-165) ---------------------------------
    @staticmethod
    def _time_to_freq(time, freq_unit=u.uHz):
        if hasattr(time, 'jd') and not isinstance(time, Quantity):
            return np.asarray(time.jd, dtype=float) * (86400.0 / (1 / freq_unit).scale)
        if not hasattr(time, 'unit'):
            return time   # assume time is already in the correct units
        return to_value(time, 1 / freq_unit)

    def _flux_to_ppm(self, flux, flux_unit=u.ppm, is_error=False):
        if isinstance(flux, np.ndarray) and not hasattr(flux, 'unit'):
            return flux   # assume already in [ppm]
        if hasattr(flux, 'unit') and hasattr(flux.unit, 'is_equivalent') and \
                self._original_flux_median is not None and \
                _is_electron_rate(flux.unit):
            med = self._original_flux_median
            ratio = np.asarray(to_value(flux / med, u.dimensionless_unscaled), dtype=float)
            return 1e6 * ratio if is_error else 1e6 * (ratio - 1)
        return to_value(flux, flux_unit)

    def _ppm_to_flux(self, value_in_ppm, power=1):
        if self._original_flux_median is not None and power == 1:
            return (1e-6 * value_in_ppm + 1) * self._original_flux_median
        elif self._original_flux_median is not None and power == 2:
            med = self._original_flux_median
            unit = getattr(med, 'unit', None)
            out = (1e-6 * value_in_ppm) * med
            return out * unit if unit is not None else out
        return Quantity(value_in_ppm, u.ppm)

    # ---- solver plumbing -------------------------------------------------------------
    def _get_solver(self):
        if self._solver is None:
            self._solver = default_solver()
        return self._solver

    def _device_tensor(self, n):
        import torch
        return torch.empty(int(n), dtype=torch.float64, device=f'cuda:{self._get_solver().device}')

    def compute(self, t, yerr=None, diag=None, check_sorted=True, quiet=False):
        """
        Compute the factorization of the GP covariance matrix
        (reference gadfly/gp.py:167-204).
        """
        if hasattr(t, 'jd') or hasattr(t, 'unit'):
            t = self._time_to_freq(t)
        if yerr is not None and hasattr(yerr, 'unit'):
            yerr = self._flux_to_ppm(yerr, is_error=True)
        if diag is not None and hasattr(diag, 'unit'):
            diag = self._flux_to_ppm(diag)

        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        if check_sorted and np.any(np.diff(t) < 0.0):
            raise ValueError("The input coordinates must be sorted")
        if check_sorted and t.ndim > 1:
            raise ValueError("The input coordinates must be one dimensional")
        if t.ndim != 1:
            raise ValueError("The input coordinates must be one dimensional")

        self._t = np.ascontiguousarray(t)
        self._size = self._t.shape[0]
        self._mean_value = self._mean(self._t)
        self._diag = np.empty(self._size, dtype=np.float64)
        if yerr is None and diag is None:
            self._diag[:] = 0.0
        elif yerr is not None:
            if diag is not None:
                raise ValueError("Only one of 'yerr' and 'diag' can be provided")
            self._diag[:] = np.square(np.asarray(yerr, dtype=np.float64))
        else:
            self._diag[:] = np.asarray(diag, dtype=np.float64)

        kb = KernelBatch([self.kernel])
        J = int(kb.J[0])
        self._kb = kb
        self._geom = Geometry.shared_t(1, self._size)
        self._w_off = np.zeros(1, dtype=np.int64)
        solver = self._get_solver()
        W = self._device_tensor(max(self._size * J, 1))
        d, W, _, logdet, status = solver.factor(
            kb, self._geom, self._t, self._diag, W=W, w_off=self._w_off)
        self._d, self._W = d, W
        if status[0] != 0:
            self._log_det = -np.inf
            self._norm = np.inf
            if not quiet:
                raise LinAlgError(f"failed to factorize or solve matrix: d[{status[0] - 1}] <= 0")
        else:
            self._log_det = float(logdet[0])
            self._norm = -0.5 * (self._log_det + self._size * np.log(2 * np.pi))

    def recompute(self, *, quiet=False):
        """Re-factor after the kernel's parameters changed (celerite2 ``recompute``)."""
        if self._t is None:
            raise RuntimeError("The processes must be initialized by running 'compute'")
        self.compute(self._t, diag=self._diag, check_sorted=False, quiet=quiet)

    def _process_input(self, y, *, inplace=False, require_vector=False):
        if self._t is None:
            raise RuntimeError("'compute' must be called before this method")
        y = np.ascontiguousarray(y, dtype=np.float64) if inplace else \
            np.array(y, dtype=np.float64, order='C', copy=True)
        if y.ndim < 1 or y.shape[0] != self._size:
            raise ValueError("dimension mismatch")
        if require_vector and y.ndim != 1:
            raise ValueError("'y' must be one dimensional")
        if y.ndim > 2:
            raise ValueError("'y' can be at most two dimensional")
        return y

    def _sweep(self, op, Y):
        """Apply one O(N J) sweep to the columns of Y ([N] or [N, k])."""
        solver = self._get_solver()
        if Y.ndim == 1:
            return solver.sweep(op, self._kb, self._geom, self._w_off, self._t, self._W, Y)
        k = Y.shape[1]
        kb = self._kb.take(np.zeros(k, dtype=np.int64))
        geom = Geometry.shared_t(k, self._size)
        w_off = np.zeros(k, dtype=np.int64)
        Z = solver.sweep(op, kb, geom, w_off, self._t, self._W, np.ascontiguousarray(Y.T))
        return np.ascontiguousarray(Z.reshape(k, self._size).T)

    # ---- celerite2 methods -----------------------------------------------------------
    def log_likelihood(self, y, *, inplace=False):
        """
        Compute the marginalized likelihood of the GP model (reference gadfly/gp.py:329-350).
        """
        if hasattr(y, 'unit'):
            y = self._flux_to_ppm(y)
        y = self._process_input(y, inplace=inplace, require_vector=True)
        if not np.isfinite(self._log_det):
            return -np.inf
        z = self._sweep(0, y - self._mean_value)
        loglike = self._norm - 0.5 * np.sum(np.square(z) / self._d)
        if not np.isfinite(loglike):
            return -np.inf
        return float(loglike)

    def dot_tril(self, y, *, inplace=False):
        """
        Dot the Cholesky factor of the GP system into a vector or matrix
        (reference gadfly/gp.py:308-327).
        """
        if hasattr(y, 'unit'):
            y = self._flux_to_ppm(y)
        y = self._process_input(y, inplace=inplace)
        sqrt_d = np.sqrt(self._d)
        z = y * (sqrt_d if y.ndim == 1 else sqrt_d[:, None])
        out = self._sweep(1, z)
        if inplace:
            y[...] = out
            return y
        return out

    def apply_inverse(self, y, *, inplace=False):
        """
        Apply the inverse of the covariance matrix to a vector or matrix
        (reference gadfly/gp.py:352-370).
        """
        if hasattr(y, 'unit'):
            y = self._flux_to_ppm(y)
        y = self._process_input(y, inplace=inplace)
        z = self._sweep(0, y)
        z = z / (self._d if z.ndim == 1 else self._d[:, None])
        out = self._sweep(2, z)
        if inplace:
            y[...] = out
            return y
        return out

    def sample(self, *, size=None, include_mean=True, return_quantity=False):
        """
        Generate random samples from the prior implied by the GP system
        (reference gadfly/gp.py:372-395).  Normal draws come from NumPy's global
        generator exactly as in celerite2 (``np.random.randn``), so seeding it reproduces
        the reference's stream.
        """
        if self._t is None:
            raise RuntimeError("'compute' must be called before this method")
        if size is None:
            n = np.random.randn(self._size)
        else:
            n = np.random.randn(self._size, size)
        result = self.dot_tril(n, inplace=True).T
        if include_mean:
            result = result + self._mean_value
        result = result - result.mean(axis=0 if result.ndim == 2 else None)
        if return_quantity:
            return self._ppm_to_flux(result)
        return result

    def conditional_distribution(self, gp, y, t=None, include_mean=True, kernel=None):
        return ConditionalDistribution(self, y, t=t, include_mean=include_mean, kernel=kernel)

    def condition(self, y, t=None, include_mean=True, kernel=None, return_quantity=False):
        """
        Condition the Gaussian process given observations ``y``
        (reference gadfly/gp.py:206-241).
        """
        if t is not None and (hasattr(t, 'jd') or hasattr(t, 'unit')):
            t = self._time_to_freq(t)
        if hasattr(y, 'unit'):
            y = self._flux_to_ppm(y)
        result = self.conditional_distribution(
            self, y, t=t, include_mean=include_mean, kernel=kernel)
        if return_quantity:
            return self._ppm_to_flux(result.mean)
        return result

    def predict(self, y, t=None, return_cov=False, return_var=False, include_mean=True,
                kernel=None, return_quantity=False):
        """
        Compute the conditional distribution (reference gadfly/gp.py:243-306).
        """
        if hasattr(y, 'unit'):
            y = self._flux_to_ppm(y)
        if t is not None and (hasattr(t, 'jd') or hasattr(t, 'unit')):
            t = self._time_to_freq(t)
        cond = self.condition(y, t=t, include_mean=include_mean, kernel=kernel)
        if return_var and return_quantity:
            return self._ppm_to_flux(cond.mean), self._ppm_to_flux(cond.variance, power=2)
        elif return_cov and return_quantity:
            return self._ppm_to_flux(cond.mean), self._ppm_to_flux(cond.covariance, power=2)
        elif return_quantity:
            return self._ppm_to_flux(cond.mean)
        elif return_var:
            return cond.mean, cond.variance
        elif return_cov:
            return cond.mean, cond.covariance
        return cond.mean


def _is_electron_rate(unit):
    try:
        return unit.is_equivalent(u.electron / u.s)
    except Exception:
        return False


class ConditionalDistribution:
    """Predictive distribution of the process given observations ``y`` (celerite2
    ``ConditionalDistribution``; reference use gadfly/gp.py:206-306, docs/gadfly/synth.rst:193-201).

    * ``mean`` at the observed times is ``y - diag * K^-1 (y - mu)``; at new times ``t`` (sorted) it
      is ``K(t, t_obs) K^-1 (y - mu)``, evaluated in O((N + M) J) on the device by
      ``gf_conditional_mean`` (celerite2's general_matmul_lower / _upper with the kernel's
      semiseparable coefficients) after the two O(N J) sweeps of ``apply_inverse``;
    * ``variance`` / ``covariance`` follow celerite2: dense cross-covariances from
      ``kernel.get_value`` on the host, solved against the stored factor by multi-right-hand-side
      sweeps on the device (in column blocks, so that N x M never has to fit at once).
    ``kernel`` predicts a different process than the one that was factored (e.g. one component of
    a sum), as in celerite2."""

    _BLOCK = 64      # right-hand sides per sweep launch

    def __init__(self, gp, y, t=None, include_mean=True, kernel=None):
        self.gp = gp
        self.y = gp._process_input(y, require_vector=True)
        self.include_mean = include_mean
        self.kernel = kernel
        if t is None:
            self.t = None
        else:
            t = np.atleast_1d(np.asarray(t, dtype=np.float64))
            if t.ndim != 1:
                raise ValueError("The input coordinates must be one dimensional")
            if np.any(np.diff(t) < 0.0):
                raise ValueError("The input coordinates must be sorted")
            self.t = np.ascontiguousarray(t)
        self._mean = self._variance = self._covariance = None

    # -- helpers -----------------------------------------------------------------------
    def _k(self):
        return self.gp.kernel if self.kernel is None else self.kernel

    def _xs(self):
        return self.gp._t if self.t is None else self.t

    def _cross_block(self, j0, j1):
        """K(t_obs, xs[j0:j1]) dense, and K^-1 of it."""
        xs = self._xs()[j0:j1]
        Kc = self._k().get_value(xs[None, :] - self.gp._t[:, None])
        return Kc, self.gp.apply_inverse(Kc)

    # -- celerite2 properties ----------------------------------------------------------
    @property
    def mean(self):
        if self._mean is None:
            gp = self.gp
            alpha = gp.apply_inverse(self.y - gp._mean_value)
            if self.t is None and self.kernel is None:
                mu = self.y - gp._diag * alpha
                if not self.include_mean:
                    mu = mu - gp._mean_value
            else:
                xs = self._xs()
                coef = KernelBatch([self._k()]).coef
                mu = gp._get_solver().conditional_mean(coef, gp._t, xs, alpha)
                if self.include_mean:
                    mu = mu + gp._mean(xs)
            self._mean = mu
        return self._mean

    @property
    def variance(self):
        if self._variance is None:
            M = len(self._xs())
            var = np.empty(M, dtype=np.float64)
            k0 = float(self._k().get_value(0.0))
            for j0 in range(0, M, self._BLOCK):
                j1 = min(M, j0 + self._BLOCK)
                Kc, S = self._cross_block(j0, j1)
                var[j0:j1] = k0 - np.einsum("ij,ij->j", Kc, S)
            self._variance = var
        return self._variance

    @property
    def covariance(self):
        if self._covariance is None:
            xs = self._xs()
            M = len(xs)
            cov = self._k().get_value(xs[:, None] - xs[None, :])
            KxsT = self._k().get_value(xs[None, :] - self.gp._t[:, None])     # [N, M]
            for j0 in range(0, M, self._BLOCK):
                j1 = min(M, j0 + self._BLOCK)
                cov[:, j0:j1] -= KxsT.T @ self.gp.apply_inverse(KxsT[:, j0:j1])
            self._covariance = cov
        return self._covariance

    def sample(self, *, size=None, regularize=None):
        """Draw from the predictive distribution (dense Cholesky of the covariance, as celerite2)."""
        mu, cov = self.mean, self.covariance.copy()
        if regularize is not None:
            cov[np.diag_indices_from(cov)] += regularize
        return np.random.multivariate_normal(mu, cov, size=size)
