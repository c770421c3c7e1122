"""
Batched hyper-parameter feeder: ``Hyperparameters.for_star`` (reference gadfly/core.py:107-333)
and the kernel assembly behind it (``StellarOscillatorKernel.__init__`` gadfly/core.py:345-394,
celerite2 ``SHOTerm.get_coefficients`` / ``TermConvolution.get_coefficients``) for B stars at once,
straight into the flat coefficient arrays the CUDA library takes.

Why: the per-star path builds ~90 Python term objects per star (9 ms/star: 18 s for the 4096 stars
of BASELINE configs[1], ten times the GPU time of their log-likelihoods).  Here the same formulas
run once over ``[B, 81]`` arrays.  The arithmetic is the per-star code's, operation
by operation; only the order of two sums differs, so coefficients agree to a few ulp
(tests/test_host.py::test_batched_feeder_matches_per_star).  The parity tests stay on the per-star
path.
"""
import numpy as np

from . import scale
from .core import _sho_psd, _solar_hyperparameter_list, Filter
from .sun import _p_mode_fit_to_sho_hyperparams
from .terms import TermConvolution, TermSum, SHOTerm  # noqa: F401  (documented counterparts)

__all__ = ['HyperparameterBatch', 'for_stars', 'kernel_batch_from_sho', 'kernel_batch_for_stars',
           'kernel_batch_for_stars_device', 'solar_tables']


class HyperparameterBatch:
    """(S0, w0, Q) of the SHO terms of B stars: flat float64 arrays in star order (granulation
    terms first, then the p-modes that survive the scaling) and CSR offsets ``j_off[B + 1]``."""

    def __init__(self, S0, w0, Q, j_off, degree=None):
        self.S0 = np.ascontiguousarray(S0, dtype=np.float64)
        self.w0 = np.ascontiguousarray(w0, dtype=np.float64)
        self.Q = np.ascontiguousarray(Q, dtype=np.float64)
        self.j_off = np.ascontiguousarray(j_off, dtype=np.int64)
        self.degree = degree        # -1 granulation, else the p-mode degree (metadata)

    def __len__(self):
        return len(self.j_off) - 1

    def star(self, b):
        """The b-th star as the reference's list of ``{"hyperparameters": ...}`` dicts."""
        sl = slice(self.j_off[b], self.j_off[b + 1])
        return [dict(hyperparameters=dict(S0=float(s), w0=float(w), Q=float(q)))
                for s, w, q in zip(self.S0[sl], self.w0[sl], self.Q[sl])]


def _alpha(bandpass, alpha, T, solver=None):
    if alpha is not None:
        return np.broadcast_to(np.asarray(alpha, dtype=np.float64), T.shape).copy(), None
    filt = Filter(bandpass)
    if filt.mean_wavelength is None:     # flat bandpass: both ratios of Morris+ (2020) Eqn 11 are 1
        return np.ones_like(T), None
    if solver is not None:
        wl, tr = scale.bandpass_grid(filt)
        return solver.bandpass_amplitude(T, wl, tr), filt.mean_wavelength
    return scale.amplitude_with_wavelength_many(filt, T), filt.mean_wavelength


_SOLAR_TABLES = None


def solar_tables():
    """The star-independent half of ``for_stars`` as the two tables the device feeder takes
    (include/gadfly_b200.h gf_feed_stars): ``gran`` [5, 3] = solar (S0, w0, Q) of the granulation
    terms; ``modes`` [81, 4 + 5] = per solar p-mode (nu, Q, Gamma, unscaled height, background PSD
    of each granulation term at nu); plus the mode degrees."""
    global _SOLAR_TABLES
    if _SOLAR_TABLES is None:
        hp = _solar_hyperparameter_list()
        gran = [i['hyperparameters'] for i in hp if i['metadata']['source'] == 'granulation']
        osc = [i for i in sorted(hp, key=lambda x: x['metadata'].get('degree', -1))
               if i['metadata']['source'] == 'oscillation']
        p_mode_vec = np.transpose([[p['hyperparameters'].get(k) for k in ('S0', 'Q')] for p in osc]).ravel()
        (S0_fit, solar_w0, Q_fit), ell = _p_mode_fit_to_sho_hyperparams(p_mode_vec)
        gS0, gw0, gQ = np.transpose([[p[k] for k in ('S0', 'w0', 'Q')] for p in gran])
        solar_nu = solar_w0 / (2 * np.pi)
        bg = _sho_psd(2 * np.pi * solar_nu[:, None], gS0[None, :], gw0[None, :], gQ[None, :])
        solar_Gamma = solar_nu / Q_fit / 2
        solar_peak = _sho_psd(2 * np.pi * solar_nu, S0_fit, solar_w0, Q_fit)
        A = 2 * np.sqrt(4 * np.pi * solar_nu * solar_peak)
        unscaled_height = 2 * A ** 2 / (np.pi * solar_Gamma)
        modes = np.concatenate([np.stack([solar_nu, Q_fit, solar_Gamma, unscaled_height], axis=1), bg], axis=1)
        _SOLAR_TABLES = (np.ascontiguousarray(np.stack([gS0, gw0, gQ], axis=1)), np.ascontiguousarray(modes),
                         ell.astype(np.int64))
    return _SOLAR_TABLES


def kernel_batch_for_stars_device(solver, mass, radius, temperature, luminosity, texp_s=60.0,
                                  bandpass='SOHO VIRGO', alpha=None, return_hyperparameters=False):
    """``kernel_batch_for_stars`` on the GPU of ``solver`` (csrc/feed.cu through gf_feed_stars):
    scaling relations, term selection, SHO -> (a, b, c, d), exposure transform and diagonal
    correction in two launches; the bandpass amplitude ratio, when the bandpass is not flat, in a
    third.  Same arithmetic as the host functions of this module, device libm."""
    from .solver import KernelBatch, GF_MAX_J_WIDE
    M, R, T, L = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in
                  (mass, radius, temperature, luminosity))
    B = len(M)
    amp, mean_wl = _alpha(bandpass, alpha, T, solver=solver)
    gran, modes, ell = solar_tables()
    delta = np.broadcast_to(np.asarray(texp_s, dtype=np.float64) * 1e-6, (B,)).copy()
    j_off, sho, coef, base, ddiag = solver.feed_stars(
        M, R, T, L, delta, gran, modes, alpha=amp, wavelength_nm=550.0 if mean_wl is None else float(mean_wl),
        want_sho=return_hyperparameters)
    if B and int(np.diff(j_off).max()) * 2 > GF_MAX_J_WIDE:
        raise ValueError(f"kernel state wider than GF_MAX_J_WIDE = {GF_MAX_J_WIDE}")
    kb = object.__new__(KernelBatch)
    kb.B = B
    kb.coef = np.ascontiguousarray(coef)
    kb.base = np.ascontiguousarray(base)
    kb.j_off = j_off
    kb.ddiag = ddiag
    kb.delta = delta
    if return_hyperparameters:
        return kb, HyperparameterBatch(sho[:, 0], sho[:, 1], sho[:, 2], j_off)
    return kb


def for_stars(mass, radius, temperature, luminosity, bandpass='SOHO VIRGO', alpha=None):
    """``Hyperparameters.for_star`` for arrays of stars (M_sun, R_sun, K, L_sun floats)."""
    M, R, T, L = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in
                  (mass, radius, temperature, luminosity))
    B = len(M)
    hp = _solar_hyperparameter_list()
    gran = [i['hyperparameters'] for i in hp if i['metadata']['source'] == 'granulation']
    osc = [i for i in sorted(hp, key=lambda x: x['metadata'].get('degree', -1))
           if i['metadata']['source'] == 'oscillation']
    p_mode_vec = np.transpose([[p['hyperparameters'].get(k) for k in ('S0', 'Q')] for p in osc]).ravel()
    (S0_fit, solar_w0, Q_fit), ell = _p_mode_fit_to_sho_hyperparams(p_mode_vec)
    gS0, gw0, gQ = np.transpose([[p[k] for k in ('S0', 'w0', 'Q')] for p in gran])

    amp, mean_wl = _alpha(bandpass, alpha, T)
    solar_nu_max = scale._NUMAX_SUN
    scaled_nu_max = solar_nu_max * (M * R ** -2 * (T / scale._T_SUN) ** -0.5)
    gran_amp = scale._granulation_power_factor(M, T, L) / \
        scale._granulation_power_factor(1.0, scale._T_SUN, 1.0)
    gran_tau = scale._tau_gran(M, T, L) / scale._tau_gran(1.0, scale._T_SUN, 1.0)

    # granulation terms [B, 5]
    g_S0 = gS0[None, :] * gran_amp[:, None] * amp[:, None]
    g_w0 = gw0[None, :] / gran_tau[:, None]
    g_Q = np.broadcast_to(gQ[None, :], g_S0.shape)
    g_keep = g_w0 > 0

    # p-modes [B, 81]
    solar_nu = solar_w0 / (2 * np.pi)
    bg = _sho_psd(2 * np.pi * solar_nu[:, None], gS0[None, :], gw0[None, :], gQ[None, :])   # [81, 5]
    bg_sum = (bg[None, :, :] * amp[:, None, None]).sum(2)                                  # [B, 81]
    scale_dnu = M ** 0.5 * R ** (-3 / 2)
    scaled_nu = scaled_nu_max[:, None] + (solar_nu - solar_nu_max)[None, :] * scale_dnu[:, None]
    scaled_w0 = 2 * np.pi * scaled_nu
    positive = scaled_w0 > 0
    nu_safe = np.where(positive, scaled_nu, 1.0)

    wl = 550.0 if mean_wl is None else float(mean_wl)
    dnu = (scale._DNU_SUN * scale_dnu)[:, None]
    i_freq = scale._velocity_to_intensity(
        scale._v_osc_kiefer_scaled(nu_safe, scaled_nu_max[:, None], dnu), T[:, None], wl)
    i_numax = scale._velocity_to_intensity(
        scale._v_osc_kiefer_scaled(scaled_nu_max[:, None], scaled_nu_max[:, None], dnu), T[:, None], wl)
    c_K = (T / 5934.0) ** 0.8
    amp_huber = L ** scale._huber_s / (M ** scale._huber_t * T ** (scale._huber_r - 1) * c_K)
    amp_huber_sun = float(scale._amplitudes_huber(1.0, scale._T_SUN, 1.0))
    factor = (i_freq / i_numax) * (amp_huber / amp_huber_sun)[:, None]

    scaled_Gamma = 1.02 * np.exp((T - scale._T_SUN) / 436.0)
    solar_Gamma = solar_nu / Q_fit / 2
    scaled_Q = Q_fit[None, :] * scaled_Gamma[:, None] / solar_Gamma[None, :]
    solar_peak = _sho_psd(2 * np.pi * solar_nu, S0_fit, solar_w0, Q_fit)
    A = 2 * np.sqrt(4 * np.pi * solar_nu * solar_peak)
    unscaled_height = 2 * A ** 2 / (np.pi * solar_Gamma)
    scaled_height = unscaled_height[None, :] * factor
    scaled_A = np.sqrt(np.pi * scaled_Gamma[:, None] * scaled_height / 2)
    scaled_peak = (scaled_A / 2) ** 2 / (4 * np.pi * nu_safe)
    p_S0 = (0.5 * (np.pi / 2) ** 0.5 * scaled_peak / scaled_Q ** 2) * bg_sum
    p_keep = positive & (p_S0 > 0)

    # flatten: per star, granulation first, then the surviving modes in table order
    S0 = np.concatenate([g_S0, p_S0], axis=1)
    w0 = np.concatenate([g_w0, scaled_w0], axis=1)
    Q = np.concatenate([g_Q, scaled_Q], axis=1)
    keep = np.concatenate([g_keep, p_keep], axis=1)
    deg = np.broadcast_to(np.concatenate([-np.ones(len(gS0), dtype=np.int64), ell.astype(np.int64)])[None, :],
                          keep.shape)
    j_off = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.int64)
    assert len(j_off) == B + 1
    return HyperparameterBatch(S0[keep], w0[keep], Q[keep], j_off, deg[keep])


def kernel_batch_from_sho(hpb, delta, eps=1e-5, solver=None):
    """``KernelBatch`` of ``StellarOscillatorKernel(hyperparameters, delta=...)`` for every star of
    a :class:`HyperparameterBatch` (delta: exposure in 1/uHz, scalar or [B]).

    SHOTerm -> (a, b, c, d) (underdamped branch; gadfly only produces Q >= 0.5) and the
    exposure-time transform + diagonal correction (A.4) in the expression order of
    ``terms.TermConvolution`` -- it is cancellation-sensitive."""
    from .solver import KernelBatch, GF_MAX_J_WIDE
    S0, w0, Q, j_off = hpb.S0, hpb.w0, hpb.Q, hpb.j_off
    B = len(j_off) - 1
    if np.any(Q < 0.5):
        raise ValueError("overdamped terms (Q < 0.5): build those kernels per star")
    delta = np.broadcast_to(np.asarray(delta, dtype=np.float64), (B,)).copy()
    widths = np.diff(j_off)
    if B and int(widths.max()) * 2 > GF_MAX_J_WIDE:
        raise ValueError(f"kernel state wider than GF_MAX_J_WIDE = {GF_MAX_J_WIDE}")
    if solver is not None:          # the same arithmetic on that solver's GPU (csrc/feed.cu, gf_feed_sho)
        kb = object.__new__(KernelBatch)
        kb.B = B
        kb.coef, kb.base, kb.ddiag = solver.feed_sho(j_off, S0, w0, Q, delta)
        kb.j_off = j_off.copy()
        kb.delta = delta
        return kb
    f = np.sqrt(np.maximum(4.0 * Q ** 2 - 1.0, eps))
    a = S0 * w0 * Q
    b = a / f
    c = 0.5 * w0 / Q
    d = c * f
    dt = np.repeat(delta, widths)
    cd = c * dt
    dd = d * dt
    c2 = c ** 2
    d2 = d ** 2
    factor = 2.0 / (dt * (c2 + d2)) ** 2
    cos_term = np.cosh(cd) * np.cos(dd) - 1
    sin_term = np.sinh(cd) * np.sin(dd)
    C1 = a * (c2 - d2) + 2 * b * c * d
    C2 = b * (c2 - d2) - 2 * a * c * d
    a_new = factor * (C1 * cos_term - C2 * sin_term)
    b_new = factor * (C2 * cos_term + C1 * sin_term)
    c2pd2 = c2 + d2
    norm = (dt * c2pd2) ** 2
    dterm = (C2 * np.cosh(cd) * np.sin(dd) - C1 * np.sinh(cd) * np.cos(dd) + (a * c + b * d) * dt * c2pd2) / norm
    ddiag = np.array([2 * np.sum(dterm[j_off[i]:j_off[i + 1]]) for i in range(B)], dtype=np.float64)

    kb = object.__new__(KernelBatch)
    kb.B = B
    kb.coef = np.ascontiguousarray(np.stack([a_new, b_new, c, d], axis=1))
    kb.base = np.ascontiguousarray(np.stack([a, b, c, d], axis=1))
    kb.j_off = j_off.copy()
    kb.ddiag = ddiag
    kb.delta = delta
    return kb


def kernel_batch_for_stars(mass, radius, temperature, luminosity, texp_s=60.0, bandpass='SOHO VIRGO',
                           alpha=None, solver=None):
    """Stellar parameters -> ``KernelBatch`` in one call (``texp_s``: exposure time in seconds);
    on the host, or with ``solver=`` on that solver's GPU."""
    if solver is not None:
        return kernel_batch_for_stars_device(solver, mass, radius, temperature, luminosity, texp_s=texp_s,
                                             bandpass=bandpass, alpha=alpha)
    hpb = for_stars(mass, radius, temperature, luminosity, bandpass=bandpass, alpha=alpha)
    return kernel_batch_from_sho(hpb, np.asarray(texp_s, dtype=np.float64) * 1e-6)
