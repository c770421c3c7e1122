#!/usr/bin/env python
"""Device feeder (csrc/feed.cu) against the host feeder (gadfly_b200/feeder.py): which terms are
kept, (S0, w0, Q), (a, b, c, d), (a', b'), the diagonal correction, the bandpass amplitude ratio,
and the time of both.  usage: python tools/feeder_check.py [n_stars]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import feeder, scale, workloads
from gadfly_b200.solver import Solver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
solver = Solver(0)
M, R, T, L = workloads.kepler_like_stars(n, 1)
for texp in (60.0, 1800.0):
    t0 = time.perf_counter()
    hpb = feeder.for_stars(M, R, T, L)
    ref = feeder.kernel_batch_from_sho(hpb, texp * 1e-6)
    t_host = time.perf_counter() - t0
    feeder.kernel_batch_for_stars_device(solver, M[:8], R[:8], T[:8], L[:8], texp_s=texp)     # warm-up
    t0 = time.perf_counter()
    got, hp_dev = feeder.kernel_batch_for_stars_device(solver, M, R, T, L, texp_s=texp, return_hyperparameters=True)
    t_dev = time.perf_counter() - t0
    same = np.array_equal(got.j_off, ref.j_off)
    print(f"texp = {texp:g} s, {n} stars, {ref.j_off[-1]} terms: host {t_host * 1e3:.1f} ms, device {t_dev * 1e3:.1f} ms "
          f"(incl. D2H of all arrays); same terms kept: {same}")
    if not same:
        bad = np.nonzero(np.diff(got.j_off) != np.diff(ref.j_off))[0]
        print("  stars with a different term count:", bad[:10], np.diff(got.j_off)[bad[:10]], np.diff(ref.j_off)[bad[:10]])
        continue
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    print("  max rel: S0 %.2e  w0 %.2e  Q %.2e | a %.2e b %.2e c %.2e d %.2e" % (
        rel(hp_dev.S0, hpb.S0), rel(hp_dev.w0, hpb.w0), rel(hp_dev.Q, hpb.Q),
        *[rel(got.base[:, k], ref.base[:, k]) for k in range(4)]))
    # a', b', Delta-diag: the closed form cancels (cosh(c D) cos(d D) - 1 for |c + i d| D -> 0), so terms
    # far below the cadence are noise in BOTH evaluations: compared on the scale of the star's k(0)
    fin_ref = np.isfinite(ref.coef).all(1)
    fin_got = np.isfinite(got.coef).all(1)
    print("  non-finite coefficient rows: host %d, device %d, same rows: %s" % (
        (~fin_ref).sum(), (~fin_got).sum(), np.array_equal(fin_ref, fin_got)))
    star = np.repeat(np.arange(ref.B), np.diff(ref.j_off))
    a0 = np.where(fin_ref, ref.coef[:, 0], 0.0)
    k0 = np.bincount(star, weights=np.abs(a0), minlength=ref.B)
    ok = fin_ref & fin_got
    zabs = np.abs((ref.c_ + 1j * ref.d_) * np.repeat(ref.delta, np.diff(ref.j_off))) if hasattr(ref, "c_") else \
        np.hypot(ref.coef[:, 2], ref.coef[:, 3]) * np.repeat(ref.delta, np.diff(ref.j_off))
    for lo, hi in ((0, 1e-3), (1e-3, 5.0), (5.0, np.inf)):
        m = ok & (zabs >= lo) & (zabs < hi)
        if not m.any():
            continue
        da = np.abs(got.coef[m, 0] - ref.coef[m, 0])
        db = np.abs(got.coef[m, 1] - ref.coef[m, 1])
        mag = np.hypot(ref.coef[m, 0], ref.coef[m, 1])
        w2 = np.minimum(zabs[m] ** 2, 1.0)
        print("  |c + i d| Delta in [%g, %g): %d terms; a', b' relative to |a' + i b'|: %.2e, relative to k(0): %.2e, "
              "x min(|w|^2, 1) (rounding of cosh(w) - 1): %.2e" % (
                  lo, hi, m.sum(), float(np.max(np.maximum(da, db) / mag)), float(np.max(np.maximum(da, db) / k0[star[m]])),
                  float(np.max(np.maximum(da, db) / mag * w2))))
    okd = np.isfinite(ref.ddiag) & np.isfinite(got.ddiag)
    print("  ddiag on the scale of k(0): %.2e (%d stars finite on both sides, %d on the host)" % (
        float(np.max(np.abs(got.ddiag - ref.ddiag)[okd] / k0[okd])), okd.sum(), np.isfinite(ref.ddiag).sum()))


class Band:
    wavelength = np.linspace(0.4, 0.9, 200)
    transmittance = np.exp(-0.5 * ((np.linspace(0.4, 0.9, 200) - 0.65) / 0.1) ** 2)


filt = g.Filter(Band)
t0 = time.perf_counter(); a_host = scale.amplitude_with_wavelength_many(filt, T); t_host = time.perf_counter() - t0
wl, tr = scale.bandpass_grid(filt)
solver.bandpass_amplitude(T[:8], wl, tr)
t0 = time.perf_counter(); a_dev = solver.bandpass_amplitude(T, wl, tr); t_dev = time.perf_counter() - t0
print(f"bandpass amplitude ratio, {n} temperatures x {len(wl)} wavelengths: host {t_host * 1e3:.0f} ms, device {t_dev * 1e3:.1f} ms, "
      f"max rel {np.max(np.abs(a_dev / a_host - 1)):.2e}")
