"""
Asteroseismic scaling relations used by ``Hyperparameters.for_star``.

Same public names and argument meaning as the reference's ``gadfly/scale.py``
(only the relations the hot path's feeder calls, reference gadfly/core.py:175,
181,186,190,243,267-276).  Inputs may be plain floats in solar units / kelvin /
microhertz, our :mod:`gadfly_b200.units` quantities, or astropy quantities.
Everything here is O(J) host arithmetic in FP64.
"""
import numpy as np
from scipy.special import wofz

from . import units as u
from .units import to_value

__all__ = [
    'p_mode_amplitudes', 'delta_nu', 'nu_max', 'tau_gran', 'granulation_amplitude', 'c_K',
    'p_mode_intensity', 'amplitude_with_wavelength',
]

# Solar parameters (reference gadfly/scale.py:20-29)
_solar_temperature = 5777 * u.K
_solar_mass = 1 * u.M_sun
_solar_radius = 1 * u.R_sun
_solar_luminosity = 1 * u.L_sun
_solar_nu_max = 3090 * u.uHz       # Huber et al. (2011)
_solar_delta_nu = 135.1 * u.uHz

# Huber et al. (2011) amplitude relation exponents (reference gadfly/scale.py:40-42)
_huber_r = 2
_huber_s = 0.886
_huber_t = 1.89

_T_SUN = 5777.0
_NUMAX_SUN = 3090.0
_DNU_SUN = 135.1


def _mtrl(mass=None, temperature=None, radius=None, luminosity=None):
    out = []
    if mass is not None:
        out.append(float(to_value(mass, u.M_sun)))
    if temperature is not None:
        out.append(float(to_value(temperature, u.K)))
    if radius is not None:
        out.append(float(to_value(radius, u.R_sun)))
    if luminosity is not None:
        out.append(float(to_value(luminosity, u.L_sun)))
    return out


def c_K(temperature):
    """Bolometric correction factor, Ballot et al. (2011) / Huber et al. (2011) Eqn 8
    (reference gadfly/scale.py:50-73)."""
    (T,) = _mtrl(temperature=temperature)
    return float((T / 5934.0) ** 0.8)


def _amplitudes_huber(M, T, L):
    return L ** _huber_s / (M ** _huber_t * T ** (_huber_r - 1) * c_K(T))


def p_mode_amplitudes(mass, temperature, luminosity):
    """p-mode power amplitude scaling, Huber et al. (2011) Eqn 9
    (reference gadfly/scale.py:83-107)."""
    M, T, L = _mtrl(mass, temperature, luminosity=luminosity)
    return float(_amplitudes_huber(M, T, L) / _amplitudes_huber(1.0, _T_SUN, 1.0))


def delta_nu(mass, radius):
    """Large frequency separation scaling, Huber et al. (2012) Eqn 3
    (reference gadfly/scale.py:176-198)."""
    M, R = _mtrl(mass, radius=radius)
    return float(M ** 0.5 * R ** (-3 / 2))


def nu_max(mass, temperature, radius):
    """Frequency of maximum power scaling, Huber et al. (2012) Eqn 4
    (reference gadfly/scale.py:201-226)."""
    M, T, R = _mtrl(mass, temperature, radius)
    return float(M * R ** -2 * (T / _T_SUN) ** -0.5)


def _tau_gran(M, T, L):
    return L / (M * T ** 3.5)


def tau_gran(mass, temperature, luminosity):
    """Granulation timescale scaling, Kjeldsen & Bedding (2011) Eqn 9
    (reference gadfly/scale.py:382-406)."""
    M, T, L = _mtrl(mass, temperature, luminosity=luminosity)
    return float(_tau_gran(M, T, L) / _tau_gran(1.0, _T_SUN, 1.0))


def _granulation_power_factor(M, T, L):
    return L ** 2 / (M ** 3 * T ** 5.5)


def granulation_amplitude(mass, temperature, luminosity):
    """Granulation amplitude scaling, Kjeldsen & Bedding (2011) Eqn 24
    (reference gadfly/scale.py:458-484)."""
    M, T, L = _mtrl(mass, temperature, luminosity=luminosity)
    return float(_granulation_power_factor(M, T, L) / _granulation_power_factor(1.0, _T_SUN, 1.0))


def _voigt(x, x_0, amplitude_L, fwhm_L, fwhm_G):
    """Voigt profile parameterised like astropy's ``Voigt1D`` (Lorentzian peak
    amplitude, Lorentzian and Gaussian FWHM), evaluated with the Faddeeva function."""
    sqrt_ln2 = np.sqrt(np.log(2.0))
    z = (2.0 * (np.asarray(x, dtype=float) - x_0) + 1j * fwhm_L) * sqrt_ln2 / fwhm_G
    return wofz(z).real * np.sqrt(np.log(2.0) * np.pi) / fwhm_G * fwhm_L * amplitude_L


def _v_osc_kiefer_scaled(freq, nu_max_uHz, delta_nu_uHz):
    """Velocity power envelope of Kiefer et al. (2018), widths scaled by delta_nu
    (reference gadfly/scale.py:515-539).  All arguments in uHz; returns m^2/s^2."""
    w = _DNU_SUN / delta_nu_uHz
    sigma = 181.8 / w   # stddev of Gaussian [uHz]
    gamma = 150.9 / w   # HWHM of Lorentzian [uHz]
    Sigma = 611.8 / w   # FWHM of Voigt [uHz]
    S = -0.1            # asymmetry parameter
    a = 3299 * 1e4      # height factor  [m^2 s^-2 Hz^-1]
    b = -581.0          # offset factor  [m^2 s^-2 Hz^-1]
    freq = np.asarray(freq, dtype=float)
    A = 1 / np.pi * (np.arctan(S * (freq - nu_max_uHz) / Sigma) + 0.5)
    voigt = _voigt(freq, nu_max_uHz, a, 2 * gamma, 2.355 * sigma)
    # (m^2 s^-2 Hz^-1) * uHz -> m^2 s^-2
    return A * (b + voigt) * 1e-6


def _velocity_to_intensity(velocity_power, T, wavelength_nm=550.0):
    # Kjeldsen & Bedding (1995) Eqn 5 (reference gadfly/scale.py:579-588)
    return 20.1 * (velocity_power / (wavelength_nm / 550.0) / (T / 5777.0) ** 2)


def p_mode_intensity(temperature, freq, nu_max, delta_nu, wavelength=550 * u.nm):
    """Relative p-mode intensity envelope, unity at ``nu_max``
    (reference gadfly/scale.py:591-632)."""
    (T,) = _mtrl(temperature=temperature)
    f = to_value(freq, u.uHz)
    numax = float(to_value(nu_max, u.uHz))
    dnu = float(to_value(delta_nu, u.uHz))
    wl = float(to_value(wavelength, u.nm))
    i_freq = _velocity_to_intensity(_v_osc_kiefer_scaled(f, numax, dnu), T, wl)
    i_numax = _velocity_to_intensity(_v_osc_kiefer_scaled(numax, numax, dnu), T, wl)
    return i_freq / i_numax


# Planck function B_nu(T) (what astropy's BlackBody evaluates by default), SI constants
_h = 6.62607015e-34
_c = 299792458.0
_kB = 1.380649e-23


def _planck_nu(wavelength_um, T):
    nu = _c / (np.asarray(wavelength_um, dtype=float) * 1e-6)
    with np.errstate(over='ignore'):
        return 2.0 * _h * nu ** 3 / _c ** 2 / np.expm1(_h * nu / (_kB * T))


def _trapz(y, x):
    return float(np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(x)))


def bandpass_grid(filter, n_wavelengths=10_000):
    """(wavelength grid [micron], transmittance on it) of the quadratures below."""
    wl = np.logspace(-1.5, 1.5, n_wavelengths)
    f_wl = np.asarray(to_value(filter.wavelength, u.um), dtype=float)
    f_tr = np.asarray(filter.transmittance, dtype=float)
    return wl, np.interp(wl, f_wl, f_tr, left=0, right=0)


def amplitude_with_wavelength_many(filter, temperatures, n_wavelengths=10_000, chunk=256):
    """:func:`amplitude_with_wavelength` for an array of temperatures at once (the batched feeder's
    non-flat bandpasses): the same quadrature, broadcast over ``[stars, wavelengths]`` in chunks."""
    T = np.atleast_1d(np.asarray(temperatures, dtype=np.float64))
    wl = np.logspace(-1.5, 1.5, n_wavelengths)  # micron
    if isinstance(filter, str):
        if filter.upper() == 'SOHO VIRGO':
            return np.ones_like(T)
        raise ValueError(
            f"filter must be 'SOHO VIRGO' or an object with wavelength/transmittance "
            f"arrays (tynt's filter tables are not bundled), but got: {filter}")
    f_wl = np.asarray(to_value(filter.wavelength, u.um), dtype=float)
    f_tr = np.asarray(filter.transmittance, dtype=float)
    filt1 = np.interp(wl, f_wl, f_tr, left=0, right=0)
    dx = np.diff(wl)

    def trapz(y):       # same summation as _trapz, along the wavelength axis
        return np.sum(0.5 * (y[:, 1:] + y[:, :-1]) * dx, axis=1)

    out = np.empty_like(T)
    for lo in range(0, len(T), chunk):
        Tc = T[lo:lo + chunk, None]
        I_nu = _planck_nu(wl[None, :], Tc)
        dI_dT = (_planck_nu(wl[None, :], Tc + 10.0) - _planck_nu(wl[None, :], Tc - 10.0)) / 20.0
        ratio_0 = trapz(dI_dT * wl * filt1) / trapz(dI_dT * wl)
        ratio_1 = trapz(I_nu * wl) / trapz(I_nu * wl * filt1)
        out[lo:lo + chunk] = ratio_0 * ratio_1
    return out


def amplitude_with_wavelength(filter, temperature, n_wavelengths=10_000, **kwargs):
    """Amplitude of intensity features in a bandpass relative to SOHO VIRGO/PMO6
    (bolometric), Morris et al. (2020) Eqn 11 (reference gadfly/scale.py:635-729).

    ``filter`` is ``'SOHO VIRGO'`` (flat, alpha = 1), or any object with
    ``wavelength`` [um or Quantity] and ``transmittance`` arrays (e.g. a
    :class:`gadfly_b200.core.Filter`).  Named SVO/tynt filters need tynt's
    tables, which are not shipped here: pass the transmittance curve instead."""
    (T,) = _mtrl(temperature=temperature)
    wl = np.logspace(-1.5, 1.5, n_wavelengths)  # micron
    if isinstance(filter, str):
        if filter.upper() == 'SOHO VIRGO':
            f_wl, f_tr = wl, np.ones_like(wl)
        else:
            raise ValueError(
                f"filter must be 'SOHO VIRGO' or an object with wavelength/transmittance "
                f"arrays (tynt's filter tables are not bundled), but got: {filter}")
    else:
        f_wl = np.asarray(to_value(filter.wavelength, u.um), dtype=float)
        f_tr = np.asarray(filter.transmittance, dtype=float)

    I_nu = _planck_nu(wl, T)
    dI_dT = (_planck_nu(wl, T + 10.0) - _planck_nu(wl, T - 10.0)) / 20.0
    filt0 = np.ones_like(wl)
    filt1 = np.interp(wl, f_wl, f_tr, left=0, right=0)
    ratio_0 = _trapz(dI_dT * wl * filt1, wl) / _trapz(dI_dT * wl * filt0, wl)
    ratio_1 = _trapz(I_nu * wl * filt0, wl) / _trapz(I_nu * wl * filt1, wl)
    return float(ratio_0 * ratio_1)
