"""
Solar inputs of the hyperparameter feeder (reference gadfly/sun.py:22-62).

The numbers come from ``gadfly_b200/data/solar_fit.json``, a compact table
derived by ``tools/make_data.py`` from the reference's
``data/hyperparameters.json`` (SOHO VIRGO/PMO6 fit) and
``data/broomhall2009_table2_labeled.ecsv`` (BiSON p-mode frequencies,
Broomhall et al. 2009, Table 2).
"""
import json
import os

import numpy as np

__all__ = ['broomhall_p_mode_freqs', 'solar_fit']

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
_cache = {}


def solar_fit():
    """The SOHO VIRGO/PMO6 fit: granulation (S0, w0, Q) x 5 and per-degree p-mode (S0, Q) x 4."""
    if 'fit' not in _cache:
        with open(os.path.join(_DATA, 'solar_fit.json')) as fh:
            _cache['fit'] = json.load(fh)
    return _cache['fit']


def broomhall_p_mode_freqs():
    """(nu [uHz], degree) of the 81 BiSON p-modes (reference gadfly/sun.py:22-33)."""
    fit = solar_fit()
    return (np.asarray(fit['bison_nu_uHz'], dtype=np.float64),
            np.asarray(fit['bison_degree'], dtype=np.int64))


def _p_mode_fit_to_sho_hyperparams(p_mode_parameters):
    """Spread 4 per-degree (S0, Q) pairs over the 81 observed modes
    (reference gadfly/sun.py:36-62).  Returns ((S0s, w0s, Qs), ell_labels)."""
    p = np.asarray(p_mode_parameters, dtype=np.float64)
    S0_ell, Q_ell = p[:4], p[4:]
    freq, ell = broomhall_p_mode_freqs()
    S0s = np.zeros_like(freq)
    Qs = np.zeros_like(freq)
    for degree in range(4):
        mask = np.where(ell == degree, 1, 0)
        S0s = S0s + mask * S0_ell[degree]
        Qs = Qs + mask * Q_ell[degree]
    return np.vstack([S0s, 2 * np.pi * freq, Qs]), ell
