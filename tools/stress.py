#!/usr/bin/env python
"""Randomised stress of the scan kernels against the CPU oracle: random widths (J = 2 .. 172),
lengths, cadence patterns (uniform, jittered, gaps, cadence changes), batch sizes and modes.
Run it under `timeout`: a hang is a finding.
usage: python tools/stress.py [seconds] [seed] [long_n] [max_batches]   (max_batches > 0: stop after that many
batches -- a deterministic case list for tests/test_gpu_baseline.py)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import batch, solver as S
from gadfly_b200.solver import Geometry, KernelBatch, default_solver
import oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
long_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0       # > 0: few long sequences of up to this length
max_batches = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(seed)
solver = default_solver()
sun = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
sun_terms = list(sun.term.terms)


def random_kernel():
    kind = rng.integers(0, 3)
    if kind == 0:       # a slice of the solar kernel
        n = int(rng.integers(1, 87))
        start = int(rng.integers(0, 87 - n))
        return g.StellarOscillatorKernel(terms=sun_terms[start:start + n], delta=sun.delta)
    n = int(rng.integers(1, 20 if kind == 1 else 87))
    terms = [g.SHOTerm(S0=float(10 ** rng.uniform(0, 3)), w0=float(10 ** rng.uniform(0.3, 3.8)),
                       Q=float(10 ** rng.uniform(-0.25, 2.5))) for _ in range(n)]
    if rng.random() < 0.3:
        return g.TermSum(*terms) if len(terms) > 1 else terms[0]
    return g.StellarOscillatorKernel(terms=terms, delta=6e-5)


def random_times(n):
    kind = rng.integers(0, 4)
    dt = np.full(n, 6e-5 * float(rng.choice([1, 1.44, 30])))
    if kind == 1:
        dt *= 1 + 1e-3 * rng.standard_normal(n)
    elif kind == 2 and n > 4:
        dt[rng.integers(1, n, size=max(1, n // 50))] += 10 ** rng.uniform(-3, 1.5)
    elif kind == 3 and n > 20:
        dt[n // 2:] *= 7.0
    return float(rng.choice([0.0, 3.0, 2.1e5])) + np.cumsum(np.abs(dt) + 6.05e-5 * (kind == 1))


t_end = time.time() + budget
cases = worst = dumped = api = batches = 0
worst_api = 0.0
while time.time() < t_end and not (max_batches and batches >= max_batches):
    batches += 1
    B = int(rng.choice([1, 2, 5, 40, 160]))
    nmax = int(rng.choice([20, 80, 400, 3000]))
    if long_n:
        nmax, B = long_n, int(rng.integers(1, 3))
    if nmax > 400:
        B = min(B, 5)                        # the oracle needs ~25 us per step at J = 172
    nk = int(rng.integers(1, 4))
    kernels = [random_kernel() for _ in range(nk)]
    ks = [kernels[int(rng.integers(0, nk))] for _ in range(B)]
    lengths = [int(x) for x in rng.integers(1, nmax, B)]
    ts = [random_times(n) for n in lengths]
    scans = [k.scan_coefficients() for k in ks]
    k0 = [np.sum(s[0]) + np.sum(s[2]) + s[6] for s in scans]
    diags = [np.full(n, 1e-4 * abs(a) * 10 ** rng.uniform(0, 2)) for n, a in zip(lengths, k0)]
    nrm = [rng.standard_normal(n) for n in lengths]
    t, dg, nn = map(np.concatenate, (ts, diags, nrm))
    flags = int(rng.choice([0, 0, 0, S.FLAG_WIDE_KERNEL, S.FLAG_REFERENCE_ORDER]))
    rows, status = batch.sample(ks, t, dg, lengths=lengths, normals=nn, solver=solver, subtract_mean=False,
                                flags=flags)
    refs = [oracle.stream(1, scans[b], ts[b], nrm[b], diag=diags[b]) for b in range(B)]
    ys = []
    for b in range(B):
        x_ref, _, st = refs[b]
        assert (status[b] == 0) == (st == 0), ("status", b, status[b], st, flags)
        if st == 0:
            err = np.max(np.abs(rows[b] - x_ref)) / max(np.max(np.abs(x_ref)), 1e-300)
            worst = max(worst, err)
            if err > 5e-9 and ts[b][0] < 10 and dumped < 5:
                os.makedirs("gpurun_out", exist_ok=True)
                np.savez(f"gpurun_out/stress_case_{dumped}.npz", scan=np.array(list(scans[b][:6]), dtype=object),
                         ddiag=scans[b][6], t=ts[b], diag=diags[b], nrm=nrm[b], x_gpu=rows[b], flags=flags)
                dumped += 1
            if err > 1e-9:
                # how well conditioned was it?  smallest pivot over k(0) from the oracle's own factor
                gp = oracle.OracleGP(scans[b], ts[b], diag=diags[b])
                dmin = float(np.min(gp.d)) / abs(k0[b]) if hasattr(gp, "d") else float("nan")
                print(f"  sample dev {err:.1e}  J={ks[b].J} N={lengths[b]} flags={flags} min d/k0={dmin:.1e} "
                      f"t0={ts[b][0]:.3g} dt0={ts[b][1] - ts[b][0] if lengths[b] > 1 else 0:.3g}", flush=True)
            assert err <= 1e-5, ("sample", b, err, ks[b].J, lengths[b], flags)
        ys.append(np.where(np.isfinite(x_ref), x_ref, 0.0) if st == 0 else nrm[b])
    ll, logdet, quad, status = batch.log_likelihood(ks, t, np.concatenate(ys), dg, lengths=lengths, solver=solver,
                                                    return_parts=True, flags=flags)
    for b in range(B):
        o_ld, o_q, st = oracle.stream(0, scans[b], ts[b], ys[b], diag=diags[b])
        assert (status[b] == 0) == (st == 0), ("status ll", b, status[b], st, flags)
        if st == 0:
            err = max(abs(logdet[b] - o_ld) / max(abs(o_ld), 1e-300), abs(quad[b] - o_q) / max(abs(o_q), 1e-300))
            worst = max(worst, err)
            if err > 1e-9:
                print(f"  loglike dev {err:.1e}  J={ks[b].J} N={lengths[b]} flags={flags}", flush=True)
            assert err <= 1e-5, ("loglike", b, err, ks[b].J, lengths[b], flags)
    cases += B
    # ---- every few batches: the stored-factor API (factor + sweeps) and the fused Philox draws ----
    if cases % 7 == 0:
        k = ks[0]
        n = max(lengths[0], 2)
        tt = random_times(n)
        dgn = np.full(n, 1e-3 * abs(k0[0]))
        ogp = oracle.OracleGP(scans[0], tt, diag=dgn)
        if np.all(ogp.d > 0):
            gp = g.GaussianProcess(k, solver=solver)
            gp.compute(tt, diag=dgn, quiet=True)
            yv = rng.standard_normal(n) * np.sqrt(abs(k0[0]))
            checks = {"log_likelihood": (gp.log_likelihood(yv), ogp.log_likelihood(yv)),
                      "dot_tril": (gp.dot_tril(yv), ogp.dot_tril(yv)),
                      "apply_inverse": (gp.apply_inverse(yv), ogp.apply_inverse(yv))}
            for name, (a, b_) in checks.items():
                a, b_ = np.atleast_1d(a), np.atleast_1d(b_)
                err = np.max(np.abs(a - b_)) / max(np.max(np.abs(b_)), 1e-300)
                worst_api = max(worst_api, err)
                lim = 1e-6 if name == "apply_inverse" else 1e-8       # K^-1 y carries cond(K) eps
                assert err <= lim, (name, err, k.J, n)
            from gadfly_b200 import philox
            xs, st = batch.sample([k], tt, dgn, seed=cases, seq0=3, solver=solver, subtract_mean=False)
            x_ref = oracle.stream(1, scans[0], tt, philox.normals(cases, 3, n), diag=dgn)[0]
            err = np.max(np.abs(xs[0] - x_ref)) / max(np.max(np.abs(x_ref)), 1e-300)
            worst_api = max(worst_api, err)
            assert err <= 1e-8, ("philox sample", err, k.J, n)
            # kernel PSD on a random grid (rtol 1e-12 against the oracle)
            om = 2 * np.pi * np.sort(10 ** rng.uniform(-1, 4, 257))
            got = solver.psd(KernelBatch([k]), om)[0]
            ref = oracle.psd(k.base_coefficients(), om, k.exposure)
            err = np.max(np.abs(got / ref - 1))
            assert err <= 1e-12, ("psd", err, k.J)
            api += 1
print(f"api checks: {api}, worst {worst_api:.2e}")
print(f"stress ok: {cases} sequences in {batches} batches, worst relative deviation {worst:.2e}")
