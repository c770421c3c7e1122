import sys; sys.path.insert(0, '.')
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import batch, solver as S
import oracle
k_sun = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
hp = g.Hyperparameters.for_star(0.9, 10.0, 4919.0, 52.3, bandpass='SOHO VIRGO', quiet=True)
k_giant = g.StellarOscillatorKernel(hp, texp=1 * g.units.min)
sol = S.Solver(0)
rng = np.random.default_rng(5)
for name, k in (("sun", k_sun), ("giant", k_giant)):
    for kind in ("uniform", "jitter", "ragged", "gaps"):
        for N in (60, 257, 1000):
            if kind == "uniform": t = np.arange(N) * 9e-5
            elif kind == "jitter": t = np.arange(N) * 9e-5 + rng.uniform(0, 1e-12, N)
            elif kind == "ragged":
                t = np.sort(rng.uniform(0, N * 9e-5, N)); t = np.cumsum(np.maximum(np.diff(t, prepend=0.0), 6.1e-5))
            else:
                t = np.arange(N) * 9e-5; t[N // 2:] += 0.5; t[3 * N // 4:] += 50.0
            y = rng.standard_normal(N) * 50
            dg = np.full(N, 1.5)
            out = {}
            for flags in (0, S.FLAG_REFERENCE_ORDER):
                ll, logdet, quad, status = batch.log_likelihood([k], t, y, dg, solver=sol, return_parts=True, flags=flags)
                out[flags] = (logdet[0], quad[0])
            o_logdet, o_quad, _ = oracle.stream(0, k.scan_coefficients(), t, y, diag=dg)
            print(f"{name:6s} {kind:8s} N={N:5d} fast: dlogdet={abs(out[0][0]/o_logdet-1):.1e} dquad={abs(out[0][1]/o_quad-1):.1e} | ref: {abs(out[2][0]/o_logdet-1):.1e} {abs(out[2][1]/o_quad-1):.1e}")
