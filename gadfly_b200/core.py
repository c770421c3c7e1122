"""
Hyperparameters and kernels: same class names, constructor arguments and
behaviour as the reference's ``gadfly/core.py`` -- the objects a user builds
before handing work to the GP solver.  All of this is O(J) host work; the
kernel objects only hold coefficient arrays for the CUDA library.
"""
import json
import warnings

import numpy as np

from . import scale
from . import units as u
from .units import to_value
from .sun import _p_mode_fit_to_sho_hyperparams, solar_fit
from .terms import SHOTerm, TermSum, TermConvolution

__all__ = [
    'Hyperparameters',
    'StellarOscillatorKernel',
    'SolarOscillatorKernel',
    'ShotNoiseKernel',
    'Filter',
    'GadflyUserWarning',
]

try:  # keep the reference's warning class when astropy exists
    from astropy.utils.exceptions import AstropyUserWarning as _BaseWarning
except Exception:  # pragma: no cover - astropy is absent on the GPU box
    _BaseWarning = UserWarning


class GadflyUserWarning(_BaseWarning):
    """Warning category for defaulted inputs / dropped kernel terms."""


def _sho_psd(omega, S0, w0, Q):
    """Stochastically driven, damped harmonic oscillator PSD (reference gadfly/core.py:33-41)."""
    return (
        np.sqrt(2 / np.pi) * S0 * w0**4 /
        ((omega**2 - w0**2)**2 + (omega**2 * w0**2 / Q**2))
    )


def _solar_hyperparameter_list():
    """The solar fit in the reference's JSON layout (list of
    ``{"hyperparameters": ..., "metadata": ...}``; reference data/hyperparameters.json)."""
    fit = solar_fit()
    out = []
    for S0, w0, Q in fit['granulation']:
        out.append(dict(hyperparameters=dict(S0=S0, w0=w0, Q=Q),
                        metadata=dict(source='granulation')))
    for degree, (S0, Q) in enumerate(zip(fit['p_mode_S0'], fit['p_mode_Q'])):
        out.append(dict(hyperparameters=dict(S0=S0, Q=Q),
                        metadata=dict(degree=degree, source='oscillation')))
    return out


class Hyperparameters(list):
    """
    Gaussian process hyperparameters for approximating the total stellar
    irradiance power spectrum (reference gadfly/core.py:44-333).
    """
    def __init__(self, hyperparameters, name=None, magnitude=None):
        super().__init__(hyperparameters)
        self.name = name
        self.magnitude = magnitude

    def __repr__(self):
        first = json.dumps(self[0], indent=4)
        return (
            f'<{self.__class__.__name__} ' +
            (f'"{self.name}" ' if self.name is not None else '') +
            f'(showing 1 of {len(self)}):\n[{first}...]>'
        )

    @staticmethod
    def _load_from_json(path):
        with open(path, 'r') as param_file:
            hyperparameters = json.load(param_file)
        return hyperparameters

    @classmethod
    def from_soho_virgo(cls, path=None, name='SOHO VIRGO/PMO6'):
        """Load the SOHO VIRGO/PMO6 total solar irradiance hyperparameters
        (reference gadfly/core.py:81-105)."""
        if path is None:
            hyperparameters = _solar_hyperparameter_list()
        else:
            hyperparameters = cls._load_from_json(path)
        return cls(hyperparameters, name=name)

    @classmethod
    def for_sun(cls, bandpass='SOHO VIRGO', name=None, **kwargs):
        """Exactly solar mass, radius, temperature and luminosity run through
        :meth:`for_star` -- what ``SolarOscillatorKernel`` builds
        (reference gadfly/core.py:454-458): 5 granulation + 81 p-mode terms."""
        return cls.for_star(mass=1 * u.M_sun, radius=1 * u.R_sun, temperature=5777 * u.K,
                            luminosity=1 * u.L_sun, bandpass=bandpass, name=name, **kwargs)

    @classmethod
    def for_stars(cls, mass, radius, temperature, luminosity, bandpass='SOHO VIRGO', alpha=None):
        """Arrays of stars at once (extension; gadfly_b200/feeder.py): the same scaling relations
        over ``[B, 81]`` arrays, returned as a flat :class:`~gadfly_b200.feeder.HyperparameterBatch`
        (``.star(b)`` gives the reference's list-of-dicts view of one star)."""
        from .feeder import for_stars
        return for_stars(mass, radius, temperature, luminosity, bandpass=bandpass, alpha=alpha)

    @classmethod
    def for_star(
            cls, mass, radius, temperature, luminosity,
            bandpass=None, name=None, quiet=False, magnitude=None, alpha=None
    ):
        """
        Apply scaling relations to the SOHO VIRGO/PMO6 solar hyperparameters for
        given stellar properties (reference gadfly/core.py:107-333).

        ``mass, radius, temperature, luminosity`` are quantities (or floats in
        M_sun, R_sun, K, L_sun).  ``alpha`` (extension) overrides the bandpass
        amplitude factor when the bandpass transmittance is not available.
        """
        M = float(to_value(mass, u.M_sun))
        R = float(to_value(radius, u.R_sun))
        T = float(to_value(temperature, u.K))
        L = float(to_value(luminosity, u.L_sun))

        hyperparameters = _solar_hyperparameter_list()
        granulation_hyperparams = [
            item for item in hyperparameters if item['metadata']['source'] == 'granulation'
        ]
        p_mode_hyperparams = [
            item for item in sorted(hyperparameters, key=lambda x: x['metadata'].get('degree', -1))
            if item['metadata']['source'] == 'oscillation'
        ]
        p_mode_vec = np.transpose(
            [[ps['hyperparameters'].get(p) for p in ['S0', 'Q']] for ps in p_mode_hyperparams]
        ).ravel()
        (S0_fit, solar_w0, Q_fit), ell_labels = _p_mode_fit_to_sho_hyperparams(p_mode_vec)

        solar_gran_S0, solar_gran_w0, solar_gran_Q = np.transpose(
            [[ps['hyperparameters'].get(p) for p in ['S0', 'w0', 'Q']]
             for ps in granulation_hyperparams]
        )

        # basic asteroseismic parameters [uHz]:
        solar_nu_max = scale._NUMAX_SUN
        scaled_nu_max = solar_nu_max * scale.nu_max(M, T, R)

        # amplitudes in the observing bandpass relative to SOHO VIRGO:
        if alpha is not None:
            amp_with_wavelength = float(alpha)
            mean_wavelength = None
        else:
            filt = Filter(bandpass)
            amp_with_wavelength = scale.amplitude_with_wavelength(filt, T)
            mean_wavelength = filt.mean_wavelength

        granulation_amp = scale.granulation_amplitude(M, T, L)
        granulation_timescale = scale.tau_gran(M, T, L)

        scaled_hyperparameters = []
        for item in granulation_hyperparams:
            params = item['hyperparameters']
            scale_S0 = params['S0'] * granulation_amp * amp_with_wavelength
            scaled_w0 = params['w0'] / granulation_timescale
            if scaled_w0 > 0:
                scaled_hyperparameters.append(
                    dict(hyperparameters=dict(S0=scale_S0, w0=scaled_w0, Q=params['Q']),
                         metadata=item['metadata'])
                )
            elif not quiet:
                msg = (
                    "The scaled solar hyperparameter with frequency "
                    f"w0(old)={params['w0']:.0f} is being scaled to "
                    f"w0(new)={scaled_w0:.0f}, which is not positive. "
                    f"This kernel term will be omitted."
                )
                warnings.warn(msg, GadflyUserWarning)

        # p-modes: these also depend on the granulation power where they sit
        solar_nu = solar_w0 / (2 * np.pi)  # [uHz]
        granulation_background_solar = _sho_psd(
            2 * np.pi * solar_nu[:, None],
            solar_gran_S0[None, :], solar_gran_w0[None, :], solar_gran_Q[None, :]
        ) * amp_with_wavelength

        scale_delta_nu = scale.delta_nu(M, R)
        solar_delta_nu = solar_nu - solar_nu_max
        scaled_delta_nu = solar_delta_nu * scale_delta_nu
        scaled_nu = scaled_nu_max + scaled_delta_nu
        scaled_w0 = 2 * np.pi * scaled_nu

        only_positive_omega = scaled_w0 > 0
        solar_nu = solar_nu[only_positive_omega]
        S0_fit = S0_fit[only_positive_omega]
        Q_fit = Q_fit[only_positive_omega]
        scaled_nu = scaled_nu[only_positive_omega]
        scaled_w0 = scaled_w0[only_positive_omega]

        wavelength_nm = 550.0 if mean_wavelength is None else float(mean_wavelength)

        p_mode_scale_factor = (
            scale.p_mode_intensity(
                T, scaled_nu, scaled_nu_max, scale._DNU_SUN * scale_delta_nu, wavelength_nm
            ) * scale.p_mode_amplitudes(M, T, L)
        )

        # quality factors
        scaled_Gamma = 1.02 * np.exp((T - scale._T_SUN) / 436.0)
        solar_Gamma = solar_nu / Q_fit / 2  # [uHz]
        scaled_Q = Q_fit * scaled_Gamma / solar_Gamma

        solar_psd_at_p_mode_peaks = _sho_psd(
            2 * np.pi * solar_nu, S0_fit, solar_w0[only_positive_omega], Q_fit
        )

        # Chaplin et al. (2008) Eqn 3
        A = 2 * np.sqrt(4 * np.pi * solar_nu * solar_psd_at_p_mode_peaks)
        unscaled_height = 2 * A ** 2 / (np.pi * solar_Gamma)
        scaled_height = unscaled_height * p_mode_scale_factor
        scaled_A = np.sqrt(np.pi * scaled_Gamma * scaled_height / 2)
        scaled_psd_at_p_mode_peaks = (scaled_A / 2) ** 2 / (4 * np.pi * scaled_nu)

        scaled_S0 = (
            0.5 * (np.pi / 2) ** 0.5 * scaled_psd_at_p_mode_peaks / scaled_Q ** 2
        ) * granulation_background_solar.sum(1)[only_positive_omega]

        scaled_w0 = np.ravel(np.repeat(scaled_w0[None, :], len(S0_fit), 0))
        scaled_Q = np.ravel(scaled_Q)

        for S0, w0, Q, degree in zip(scaled_S0, scaled_w0, scaled_Q, ell_labels):
            if np.all(np.array([S0, w0]) > 0):
                scaled_hyperparameters.append(
                    dict(hyperparameters=dict(S0=float(S0), w0=float(w0), Q=float(Q)),
                         metadata=dict(source='oscillation', scaled=True, degree=int(degree)))
                )

        return cls(scaled_hyperparameters, name, magnitude)


class StellarOscillatorKernel(TermConvolution):
    """
    A sum of SHO kernels approximating the stellar irradiance power spectrum,
    integrated over the exposure time (reference gadfly/core.py:336-427).
    """
    def __init__(self, hyperparameters=None, texp=None, delta=None, name=None, terms=None):
        kernel_components = []

        if hyperparameters is not None:
            self.hyperparameters = hyperparameters
            if name is None and getattr(hyperparameters, 'name', None) is not None:
                name = hyperparameters.name
            kernel_components += [SHOTerm(**p['hyperparameters']) for p in self.hyperparameters]

        if terms is not None:
            kernel_components += list(terms)

        self.name = name
        term_sum = TermSum(*kernel_components)

        if delta is None:
            if texp is None:
                default_exp = 1 * u.min
                msg = (
                    "An exposure time is required to construct the kernel. gadfly will assume "
                    f"a default exposure time of 1 min. To prevent this warning, supply "
                    f"the kernel with the `texp` keyword argument."
                )
                warnings.warn(msg, GadflyUserWarning)
                texp = default_exp
            delta = float(to_value(texp, u.inv_uHz, assume=u.s))

        super().__init__(term_sum, delta)

    def plot(self, **kwargs):
        from .psd import plot_power_spectrum
        return plot_power_spectrum(kernel=self, **kwargs)

    @classmethod
    def _from_terms(cls, terms, delta=None, name=None):
        return cls(terms=terms, delta=delta, name=name)

    def __add__(self, other):
        """Assumes ``other`` is a SHOTerm or subclass (reference gadfly/core.py:405-427)."""
        if not isinstance(other, list):
            other_names = [other.name]
            other = [other]
        else:
            other_names = [t.name for t in other]

        name = ""
        if self.name is not None:
            name += self.name
        for other_name in other_names:
            if other_name is not None:
                if len(name):
                    name += " + " + other_name
                else:
                    name += other_name

        return StellarOscillatorKernel._from_terms(
            list(self.term.terms) + other, delta=self.delta, name=name
        )


class SolarOscillatorKernel(StellarOscillatorKernel):
    """
    :class:`StellarOscillatorKernel` with the solar SOHO VIRGO/PMO6 hyperparameters
    run through :meth:`Hyperparameters.for_star` at exactly solar mass, radius,
    temperature and luminosity (reference gadfly/core.py:430-461).
    """
    def __init__(self, texp=None, delta=None, bandpass=None, name=None):
        hp = Hyperparameters.for_star(
            mass=1 * u.M_sun, radius=1 * u.R_sun,
            temperature=5777 * u.K, luminosity=1 * u.L_sun,
            bandpass=bandpass
        )
        super().__init__(hp, texp=texp, delta=delta, name=name)


class ShotNoiseKernel(SHOTerm):
    """
    A SHO term approximating shot noise: very large w0, Q = 1/2
    (reference gadfly/core.py:464-544).
    """
    w0 = 1e7   # intentionally really large [uHz]
    Q = 0.5    # value does not matter much if w0 >>> 1

    def __init__(self, *args, name=None, **kwargs):
        if name is None:
            name = "Shot noise"
        super().__init__(*args, **kwargs)
        self.name = name

    @classmethod
    def from_kepler_magnitude(cls, kepler_mag, n_cadences):
        """Shot noise for a Kepler magnitude and number of cadences; the arithmetic of
        the reference's ``from_kepler_light_curve`` (gadfly/core.py:516-520) without lightkurve."""
        norm = 2 * np.pi / n_cadences ** 0.5
        _, unscaled_S0 = cls.kepler_mag_to_noise_amplitude(kepler_mag)
        S0 = (unscaled_S0 * norm) ** 0.5
        return cls(S0=float(S0), w0=cls.w0, Q=cls.Q)

    @classmethod
    def from_kepler_light_curve(cls, light_curve):
        return cls.from_kepler_magnitude(light_curve.meta['KEPMAG'], len(light_curve.time))

    @staticmethod
    def kepler_mag_to_noise_amplitude(kepler_mag):
        """Kepler noise in 6 hour bins, Jenkins et al. (2010) (reference gadfly/core.py:522-544).
        Returns (lower, upper) in ppm^2."""
        c = 3.46 * 10 ** (0.4 * (12 - kepler_mag) + 8)
        sigma_lower = np.sqrt(
            c + 7e6 * np.max([np.ones_like(kepler_mag), kepler_mag / 14], axis=0) ** 4
        ) / c
        sigma_upper = np.sqrt(c + 7e7) / c
        return 1e6 * np.array([sigma_lower, sigma_upper])


class Filter:
    """
    Photometric bandpass transmittance (reference gadfly/core.py:547-621, a tynt
    wrapper there).  tynt's filter tables are not available offline, so a filter
    is either ``'SOHO VIRGO'`` (bolometric, flat) or built from arrays:
    ``Filter(wavelength, transmittance)``.
    """
    default_filter = 'Kepler/Kepler.K'

    def __init__(self, identifier_or_filter, transmittance=None, download=False):
        if transmittance is not None:
            self.wavelength = np.asarray(to_value(identifier_or_filter, u.um), dtype=float)
            self.transmittance = np.asarray(transmittance, dtype=float)
            return
        if identifier_or_filter is None:
            msg = (
                "An observing bandpass is required to construct the kernel. gadfly "
                f'will assume the default filter "{self.default_filter}". To prevent '
                f"this warning, supply the Hyperparameters with the `bandpass` "
                "keyword argument."
            )
            warnings.warn(msg, GadflyUserWarning)
            identifier_or_filter = self.default_filter
        if isinstance(identifier_or_filter, str) and identifier_or_filter.upper() == 'SOHO VIRGO':
            self.wavelength = np.logspace(-1.5, 1.5, 1000)  # [um]
            self.transmittance = np.ones_like(self.wavelength)
        elif hasattr(identifier_or_filter, 'wavelength') and \
                hasattr(identifier_or_filter, 'transmittance'):
            self.wavelength = np.asarray(
                to_value(identifier_or_filter.wavelength, u.um), dtype=float)
            self.transmittance = np.asarray(identifier_or_filter.transmittance, dtype=float)
        else:
            raise ValueError(
                f'The observing bandpass "{identifier_or_filter}" needs tynt\'s filter tables, '
                f"which are not bundled. Use bandpass='SOHO VIRGO', pass a "
                f"Filter(wavelength, transmittance), or give `alpha=` to Hyperparameters.for_star."
            )

    @property
    def mean_wavelength(self):
        """Transmittance-weighted mean wavelength [nm]; None for a flat bandpass."""
        if np.all(self.transmittance == 1):
            return None
        return float(np.average(self.wavelength, weights=self.transmittance)) * 1e3
