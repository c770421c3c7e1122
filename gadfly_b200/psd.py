"""
Power spectra: the kernel PSD on dense frequency grids (the hot-path part, CUDA
kernel K5) and a small observed-power-spectrum container.

Reference: gadfly/psd.py.  The hot-path call there is ``kernel.get_psd(2 pi f)``
(gadfly/psd.py:151, gadfly/tests/test_core.py:34), which here runs on the GPU through
:func:`kernel_psd` / ``Term.get_psd``.  ``PowerSpectrum`` keeps the reference's
attribute and method names (``frequency``, ``power``, ``error``, ``omega``, ``bin``,
``cutout``, ``from_light_curve``) with plain float64 arrays in uHz and ppm^2/uHz;
estimating an observed spectrum is host-side numpy (FFT, reference normalisation
gadfly/psd.py:566-587) and is only used by the sample -> PSD round-trip check.
"""
import numpy as np

from . import units as u
from .units import to_value

__all__ = ['PowerSpectrum', 'kernel_psd', 'bin_power_spectrum', 'plot_power_spectrum',
           'power_spectra', 'bin_power_spectra', 'bin_ranges']


def kernel_psd(kernels, frequency, solver=None, out=None):
    """PSD [ppm^2/uHz] of one kernel or a list of kernels at ``frequency`` [uHz].

    Returns ``[F]`` for a single kernel and ``[B, F]`` for a list; evaluated by the CUDA
    PSD kernel with the summed-Lorentzian closed form times the exposure sinc^2."""
    from .solver import KernelBatch, default_solver
    single = not isinstance(kernels, (list, tuple))
    klist = [kernels] if single else list(kernels)
    freq = np.ascontiguousarray(to_value(frequency, u.uHz), dtype=np.float64)
    omega = 2 * np.pi * freq.ravel()
    solver = solver or default_solver()
    res = solver.psd(KernelBatch(klist), omega, out=out)
    if hasattr(res, 'data_ptr'):
        return res
    return res[0].reshape(freq.shape) if single else res.reshape((len(klist),) + freq.shape)


# ---- the same estimators for B light curves at once, on the device -------------
def power_spectra(flux, d_days, include_zero_freq=False, solver=None, out=None):
    """FFT power spectra of B evenly sampled light curves ``flux[B, N]`` [ppm] (numpy array or CUDA
    tensor: a tensor never leaves the device) with the reference's normalisation
    (``PowerSpectrum._fft``, gadfly/psd.py:566-587).  -> (frequency[F] uHz, power[B, F] ppm^2/uHz, norm)."""
    from .solver import default_solver
    solver = solver or default_solver()
    if hasattr(flux, 'data_ptr'):
        B, N = (1, flux.numel()) if flux.dim() == 1 else (flux.shape[0], flux.shape[1])
    else:
        flux = np.ascontiguousarray(np.atleast_2d(flux), dtype=np.float64)
        B, N = flux.shape
    d = d_days * 86400.0 * 1e-6
    freq = np.fft.rfftfreq(N, d)
    if not include_zero_freq:
        freq = freq[1:]
    if out is None and hasattr(flux, 'data_ptr') and flux.is_cuda:
        import torch
        out = torch.empty((B, len(freq)), dtype=torch.float64, device=flux.device)
    power = solver.power_spectrum(flux, B, N, d, include_zero=include_zero_freq, out=out)
    return freq, power, d / (2 * np.pi) ** 0.5 / N


def bin_ranges(axis, bins):
    """Index ranges of ``scipy.stats.binned_statistic``'s bins on a monotone axis (the last bin is
    closed on the right): -> (edges[nb + 1], lo[nb], cnt[nb])."""
    axis = np.asarray(axis, dtype=float)
    edges = np.linspace(axis.min(), axis.max(), int(bins) + 1) if np.isscalar(bins) else np.asarray(bins, dtype=float)
    nb = len(edges) - 1
    lo = np.searchsorted(axis, edges[:-1], side='left')
    hi = np.searchsorted(axis, edges[1:], side='left')
    hi[-1] = np.searchsorted(axis, edges[-1], side='right')
    return edges, lo.astype(np.int64), (hi - lo).astype(np.int64)


def bin_power_spectra(frequency, power, bins=None, log=True, constant=1, solver=None):
    """``bin_power_spectrum`` (reference gadfly/psd.py:229-297) for B spectra ``power[B, F]`` on one
    frequency grid, on the device.  -> (bin centre frequencies[nb], stat[B, nb], err[B, nb])."""
    from .solver import default_solver
    solver = solver or default_solver()
    freq = np.asarray(frequency, dtype=float)
    axis = np.log10(freq) if log else freq
    if bins is None:
        bins = max(len(axis) // 10000, 1)
    edges, lo, cnt = bin_ranges(axis, bins)
    B = 1 if (hasattr(power, 'dim') and power.dim() == 1) or np.ndim(power) == 1 else power.shape[0]
    stat = err = None
    if hasattr(power, 'data_ptr') and power.is_cuda:
        import torch
        stat = torch.empty((B, len(lo)), dtype=torch.float64, device=power.device)
        err = torch.empty((B, len(lo)), dtype=torch.float64, device=power.device)
        axis_in = torch.as_tensor(axis, device=power.device)
    else:
        axis_in = axis
    stat, err = solver.bin_power(power, B, axis_in, lo, cnt, constant=constant, stat=stat, err=err)
    centers = 0.5 * (edges[1:] + edges[:-1])
    return (10 ** centers if log else centers), stat, err


def _spectral_binning(y, all_x, lo, hi):
    if hi > lo and all_x[hi] - all_x[lo] > 0:
        x = all_x[lo:hi + 1]
        return float(np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(x)) / (all_x[hi] - all_x[lo]))
    return float(y[0])


def _spectral_binning_err(y, all_x, lo, hi, constant=1):
    if hi > lo and all_x[hi] - all_x[lo] > 0:
        mean_x = np.nanmean(all_x[lo:hi + 1])
        gaussian_term = np.nanstd(y) / len(y) ** 0.5
        non_gaussian_term = mean_x / (all_x[hi] - all_x[lo]) / constant
        return float(gaussian_term * non_gaussian_term)
    return float(y[0])


def bin_power_spectrum(power_spectrum, bins=None, log=True, **kwargs):
    """Bin a power spectrum with (log-)spaced frequency bins: trapezoidal mean per bin and
    the reference's error estimate (reference gadfly/psd.py:186-297)."""
    freq = np.asarray(power_spectrum.frequency, dtype=float)
    power = np.asarray(power_spectrum.power, dtype=float)
    axis = np.log10(freq) if log else freq
    if bins is None:
        bins = max(len(axis) // 10000, 1)
    edges = np.linspace(axis.min(), axis.max(), bins + 1) if np.isscalar(bins) else np.asarray(bins)
    nb = len(edges) - 1
    # scipy.stats.binned_statistic convention: last bin is closed on the right
    which = np.searchsorted(edges, axis, side='right') - 1
    which[axis == edges[-1]] = nb - 1
    stat = np.full(nb, np.nan)
    err = np.full(nb, np.nan)
    order = np.argsort(which, kind='stable')
    sorted_which = which[order]
    starts = np.searchsorted(sorted_which, np.arange(nb), side='left')
    stops = np.searchsorted(sorted_which, np.arange(nb), side='right')
    for i in range(nb):
        idx = order[starts[i]:stops[i]]
        if len(idx) == 0:
            continue
        lo, hi = int(idx[0]), int(idx[-1])
        stat[i] = _spectral_binning(power[idx], axis, lo, hi)
        err[i] = _spectral_binning_err(power[idx], axis, lo, hi, **kwargs)
    centers = 0.5 * (edges[1:] + edges[:-1])
    freq_bins = 10 ** centers if log else centers
    name = (power_spectrum.name if power_spectrum.name is not None else 'Power spectrum') + ' (binned)'
    return PowerSpectrum(freq_bins, stat, err, name=name)


class PowerSpectrum:
    """An observed power spectrum (reference gadfly/psd.py:364-649)."""

    def __init__(self, frequency, power, error=None, name=None, norm=None, detrended_lc=None):
        self.frequency = np.asarray(to_value(frequency, u.uHz), dtype=float)
        self.power = np.asarray(to_value(power, u.ppm ** 2 / u.uHz), dtype=float)
        self.error = None if error is None else np.asarray(
            to_value(error, u.ppm ** 2 / u.uHz), dtype=float)
        self.name = name
        self.norm = norm
        self.detrended_lc = detrended_lc

    @property
    def omega(self):
        """Angular frequency 2 pi f with f in [uHz]."""
        return 2 * np.pi * self.frequency

    @property
    def light_curve_rms(self):
        return (self.power * self.norm * 1e6) ** 0.5

    def bin(self, bins=None, **kwargs):
        return bin_power_spectrum(self, bins, **kwargs)

    def kernel_psd(self, kernel, solver=None):
        """The kernel's PSD on this spectrum's frequency grid (GPU)."""
        return kernel_psd(kernel, self.frequency, solver=solver)

    @classmethod
    def from_light_curve(cls, light_curve, flux=None, method='fft', include_zero_freq=False,
                         name=None):
        """FFT power spectrum of an evenly sampled light curve given as ``(time [d], flux [ppm])``
        or an object with ``.time``/``.flux`` (reference gadfly/psd.py:442-587 without the
        lightkurve detrending / gap-filling front end)."""
        if method.lower() != 'fft':
            raise ValueError('only method="fft" is available (Lomb-Scargle needs astropy)')
        if flux is None:
            time, flux = light_curve.time, light_curve.flux
        else:
            time = light_curve
        time = np.asarray(time.jd if hasattr(time, 'jd') else to_value(time, u.d), dtype=float)
        flux = np.asarray(to_value(flux, u.ppm), dtype=float)
        d = float(np.median(np.diff(time)))          # [d]
        freq, power, norm = cls._fft(flux, d)
        if not include_zero_freq:
            freq, power = freq[1:], power[1:]
        return cls(freq, power, name=name, norm=norm)

    @staticmethod
    def _fft(flux_ppm, d_days):
        """(frequency [uHz], power [ppm^2/uHz], norm [1/uHz]); reference gadfly/psd.py:566-587."""
        d = d_days * 86400.0 * 1e-6                  # cadence in 1/uHz
        freq = np.fft.rfftfreq(len(flux_ppm), d)     # uHz
        fft = np.fft.rfft(flux_ppm)
        norm = d / (2 * np.pi) ** 0.5 / len(flux_ppm)
        power = np.real(fft * np.conj(fft)) * norm
        return freq, power, norm

    def cutout(self, frequency_min=None, frequency_max=None):
        fmin = 0.0 if frequency_min is None else float(to_value(frequency_min, u.uHz))
        fmax = np.inf if frequency_max is None else float(to_value(frequency_max, u.uHz))
        bounds = (self.frequency <= fmax) & (self.frequency >= fmin)
        name = (self.name if self.name is not None else 'Power spectrum') + ' (cutout)'
        err = None if self.error is None else self.error[bounds]
        return PowerSpectrum(self.frequency[bounds], self.power[bounds], err, name=name,
                             norm=self.norm)

    def plot(self, **kwargs):
        return plot_power_spectrum(obs=self, **kwargs)


def plot_power_spectrum(ax=None, kernel=None, obs=None, freq=None, n_samples=1000, **kwargs):
    """Plot kernel and/or observed power spectra (reference gadfly/psd.py:36-183).  Plotting is
    outside the hot path; only the default grids and the GPU PSD call are kept."""
    try:
        import matplotlib.pyplot as plt
    except ImportError as exc:  # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not part of gadfly_b200") from exc
    if ax is None:
        _, ax = plt.subplots(figsize=kwargs.pop('figsize', (8, 4)))
    if kernel is not None:
        if freq is None:
            freq = np.sort(np.concatenate([np.logspace(-1, 3.5, n_samples // 2),
                                           np.linspace(2000, 4500, n_samples // 2)]))
        ax.loglog(freq, kernel_psd(kernel, freq), label=kwargs.get('label_kernel', kernel.name))
    if obs is not None:
        ax.loglog(obs.frequency, obs.power, label=kwargs.get('label_obs', obs.name))
    ax.set(xlabel='Frequency [$\\mu$Hz]', ylabel='Power [ppm$^2$ / $\\mu$Hz]')
    return ax.figure, ax
