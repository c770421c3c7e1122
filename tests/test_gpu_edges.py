"""GPU parity on the edges of the scan kernel's fast paths: very short sequences (row-ring
hand-overs), cadence changes / jitter / gaps (producer tables, renormalisation), absolute time
stamps (large phases), overdamped and white-noise-like extra terms (reference
gadfly/core.py:405-427, 464-544)."""
import numpy as np
import pytest

import gadfly_b200 as g
from gadfly_b200 import batch, solver as S
from gadfly_b200.solver import Geometry, KernelBatch
import oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _maxrel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b))


def _check(solver, kernel, t, diag=None, seed=0, rtol=RTOL):
    rng = np.random.default_rng(seed)
    N = len(t)
    scan = kernel.scan_coefficients()
    nrm = rng.standard_normal(N)
    dg = None if diag is None else np.broadcast_to(np.asarray(diag, dtype=float), (N,)).copy()
    x_ref, ld_ref, st_ref = oracle.stream(1, scan, t, nrm, diag=dg)
    assert st_ref == 0
    x, status = batch.sample([kernel], t, dg, normals=nrm, solver=solver, subtract_mean=False)
    assert status[0] == 0
    assert _maxrel(x[0], x_ref) <= rtol
    ll, logdet, quad, status = batch.log_likelihood([kernel], t, x_ref, dg, solver=solver, return_parts=True)
    o_ld, o_q, _ = oracle.stream(0, scan, t, x_ref, diag=dg)
    assert status[0] == 0
    assert logdet[0] == pytest.approx(o_ld, rel=rtol) and quad[0] == pytest.approx(o_q, rel=rtol)


@pytest.mark.parametrize("N", [1, 2, 3, 7, 8, 9, 15, 16, 17, 18, 23, 24, 25, 33])
def test_short_sequences(solver, solar_kernel, N):
    _check(solver, solar_kernel, np.arange(N) * 6e-5, seed=N)


def test_cadence_change_jitter_and_gaps(solver, solar_kernel):
    rng = np.random.default_rng(3)
    dt = np.concatenate([
        np.full(200, 6e-5),                         # 1 min
        np.full(150, 1.8e-3),                       # 30 min: new tables
        np.full(100, 6e-5) * (1 + 1e-13 * rng.standard_normal(100)),    # jitter inside the first-order window
        np.full(100, 6e-5) * (1 + 1e-3 * rng.standard_normal(100)),     # jitter outside: exact rows
        [0.5], np.full(60, 6e-5), [40.0], np.full(60, 6e-5),            # gaps of 5.8 and 463 days
    ])
    t = np.cumsum(dt)
    _check(solver, solar_kernel, t, diag=25.0)


def test_absolute_time_stamps(solver, solar_kernel):
    """BJD-like time stamps (2.1e5 1/uHz): phases d t ~ 5e9 rad, outside the table fast path and
    the Cody-Waite range -- the rows must still be the correctly rounded cos / sin of d * t."""
    t = 2.1e5 + np.arange(400) * 6e-5
    _check(solver, solar_kernel, t, diag=25.0)


def test_extra_terms_overdamped_and_shot_noise(solver, solar_kernel):
    """kernel + term (reference gadfly/core.py:405-427): an overdamped SHO (two real terms) and a
    critically damped one (Q = 0.5: the f = sqrt(eps) branch the ShotNoiseKernel takes,
    gadfly/core.py:471-474)."""
    over = g.SHOTerm(S0=2000.0, w0=30.0, Q=0.3)
    shot = g.SHOTerm(S0=0.05, w0=2.0e4, Q=0.5)
    k = solar_kernel + over
    assert k.J == solar_kernel.J + 2
    _check(solver, k, np.arange(500) * 6e-5)
    k2 = solar_kernel + shot
    assert k2.J == solar_kernel.J + 2
    _check(solver, k2, np.arange(500) * 8.64e-5)
    gp = g.GaussianProcess(k2, t=np.arange(300) * 8.64e-5, solver=solver)
    assert np.isfinite(gp.log_likelihood(np.zeros(300)))
    # The ShotNoiseKernel's own w0 = 1e7 under a 1-min exposure is outside FP64 for the exposure
    # transform itself (c * delta = 600: a' ~ e^600 cancels against the diagonal correction): the
    # oracle and the GPU path must agree that the first pivot is not positive.
    k3 = solar_kernel + g.SHOTerm(S0=1e-3, w0=1e7, Q=0.5)
    t = np.arange(64) * 8.64e-5
    assert oracle.stream(0, k3.scan_coefficients(), t, np.zeros(64))[2] == 1
    _, _, _, status = batch.log_likelihood([k3], t, np.zeros(64), solver=solver, return_parts=True)
    assert status[0] == 1


def test_shared_light_curve_many_hyperparameter_sets(solver, solar_kernel):
    """BASELINE configs[3] in small: one light curve, many hyper-parameter sets, y and diag passed
    once (GF_FLAG_SHARED_Y) -- same numbers as replicating them per set."""
    from gadfly_b200 import solver as S
    from gadfly_b200.solver import Geometry, KernelBatch
    N, B = 900, 7
    rng = np.random.default_rng(21)
    t = np.arange(N) * 6e-5
    y = rng.standard_normal(N) * 250
    dg = np.full(N, 15.0 ** 2)
    kernels = [g.StellarOscillatorKernel(
        terms=[g.SHOTerm(S0=p.S0 * f, w0=p.w0 * (2 - f), Q=p.Q) for p in solar_kernel.term.terms],
        delta=solar_kernel.delta) for f in np.linspace(0.9, 1.1, B)]
    kb = KernelBatch(kernels)
    geom = Geometry.shared_t(B, N)
    ld0, q0, s0 = solver.loglike(kb, geom, t, np.tile(y, B), np.tile(dg, B))
    ld1, q1, s1 = solver.loglike(kb, geom, t, y, dg, flags=S.FLAG_SHARED_Y)
    ld2, q2, s2 = solver.loglike(kb, geom, t, y, dg, flags=S.FLAG_SHARED_Y | S.FLAG_REFERENCE_ORDER)
    assert s0.tolist() == [0] * B and s1.tolist() == [0] * B and s2.tolist() == [0] * B
    np.testing.assert_array_equal(ld1, ld0)
    np.testing.assert_array_equal(q1, q0)
    np.testing.assert_allclose(ld2, ld0, rtol=RTOL)
    np.testing.assert_allclose(q2, q0, rtol=RTOL)
    assert np.ptp(-0.5 * (q0 + ld0)) > 1.0          # the sets really differ


def test_full_length_light_curve(solver, solar_kernel):
    """BASELINE's full sequence length: 2^20 points of 1-min cadence (728 days), fused Philox
    sample and fused log-likelihood against the CPU oracle on the same inputs -- 1M steps of
    lazy-decay frames, producer tables and renormalisations without drift."""
    from gadfly_b200 import philox
    N = 1 << 20
    t = np.arange(N) * 6e-5
    scan = solar_kernel.scan_coefficients()
    x, status = batch.sample([solar_kernel], t, seed=5, solver=solver, subtract_mean=False)
    nrm = philox.normals(5, 0, N)
    x_ref, ld_ref, st_ref = oracle.stream(1, scan, t, nrm, fast=True)
    assert status[0] == 0 and st_ref == 0
    assert _maxrel(x[0], x_ref) <= RTOL
    ll, logdet, quad, status = batch.log_likelihood([solar_kernel], t, x_ref, solver=solver, return_parts=True)
    o_ld, o_q, _ = oracle.stream(0, scan, t, x_ref, fast=True)
    assert status[0] == 0
    assert logdet[0] == pytest.approx(o_ld, rel=RTOL) and quad[0] == pytest.approx(o_q, rel=RTOL)
    assert quad[0] == pytest.approx(np.sum(nrm * nrm), rel=1e-8)       # |L^-1 L n|^2 = |n|^2


def _narrow_kernels():
    rng = np.random.default_rng(17)
    def sho(n):
        return [g.SHOTerm(S0=float(10 ** rng.uniform(0, 3)), w0=float(10 ** rng.uniform(0.5, 3.5)),
                          Q=float(10 ** rng.uniform(-0.2, 2))) for _ in range(n)]
    return {2: g.SHOTerm(S0=3.0, w0=40.0, Q=2.5),
            10: None,                                               # granulation-only, set by the test
            16: g.StellarOscillatorKernel(terms=sho(8), delta=6e-5),
            32: g.StellarOscillatorKernel(terms=sho(16), delta=6e-5)}


@pytest.mark.parametrize("J", [2, 10, 16, 32])
def test_narrow_kernels_one_warp_per_sequence(solver, solar_kernel, J):
    """Batches whose kernels all have J <= 32 take the one-warp-per-sequence kernel
    (csrc/scan_small.cu): against the oracle, and against the wide kernel on the same inputs."""
    from gadfly_b200 import solver as S, philox
    from gadfly_b200.solver import Geometry, KernelBatch
    k = _narrow_kernels()[J]
    if k is None:
        k = g.StellarOscillatorKernel(terms=list(solar_kernel.term.terms[:5]), delta=solar_kernel.delta)
    assert k.J == J
    B = 37                                      # more sequences than warps of a CTA, ragged lengths
    rng = np.random.default_rng(J)
    lengths = [int(x) for x in rng.integers(1, 200, B)]
    lengths[0], lengths[1], lengths[2] = 1, 32, 33
    ts = [np.cumsum(rng.uniform(6.1e-5, 3e-4, n)) for n in lengths]
    scan = k.scan_coefficients()
    k0 = np.sum(scan[0]) + np.sum(scan[2]) + scan[6]
    diags = [np.full(n, 1e-4 * k0) for n in lengths]
    nrm = [rng.standard_normal(n) for n in lengths]
    t, dg, nn = map(np.concatenate, (ts, diags, nrm))
    rows, status = batch.sample([k] * B, t, dg, lengths=lengths, normals=nn, solver=solver, subtract_mean=False)
    assert status.tolist() == [0] * B
    xs = []
    for b in range(B):
        x_ref = oracle.stream(1, scan, ts[b], nrm[b], diag=diags[b])[0]
        assert _maxrel(rows[b], x_ref) <= RTOL, b
        xs.append(x_ref)
    y = np.concatenate(xs)
    ll, logdet, quad, status = batch.log_likelihood([k] * B, t, y, dg, lengths=lengths, solver=solver,
                                                    return_parts=True)
    ll_w, logdet_w, quad_w, status_w = batch.log_likelihood([k] * B, t, y, dg, lengths=lengths, solver=solver,
                                                            return_parts=True, flags=S.FLAG_WIDE_KERNEL)
    assert status.tolist() == [0] * B and status_w.tolist() == [0] * B
    for b in range(B):
        o_ld, o_q, _ = oracle.stream(0, scan, ts[b], xs[b], diag=diags[b])
        assert logdet[b] == pytest.approx(o_ld, rel=RTOL) and quad[b] == pytest.approx(o_q, rel=RTOL), b
    np.testing.assert_allclose(logdet, logdet_w, rtol=RTOL)
    np.testing.assert_allclose(quad, quad_w, rtol=RTOL)
    # fused Philox draws on the narrow path
    N = 100
    tt = np.arange(N) * 6e-5
    dfl = np.full(N, 1e-4 * k0)       # white-noise floor: keeps the smooth single-term cases well conditioned
    x, st = batch.sample([k] * 3, tt, np.tile(dfl, (3, 1)), seed=9, seq0=4, solver=solver, subtract_mean=False)
    for b in range(3):
        assert _maxrel(x[b], oracle.stream(1, scan, tt, philox.normals(9, 4 + b, N), diag=dfl)[0]) <= RTOL
    # stored factor (GaussianProcess.compute) and the sweeps on it, narrow path against the wide one
    from gadfly_b200.solver import KernelBatch as KB
    geom1 = Geometry.shared_t(1, N)
    d_n, W_n, w_off, ld_n, st_n = solver.factor(KB([k]), geom1, tt, dfl)
    d_w, W_w, _, ld_w, st_w = solver.factor(KB([k]), geom1, tt, dfl, flags=S.FLAG_WIDE_KERNEL)
    assert st_n[0] == 0 and st_w[0] == 0
    np.testing.assert_allclose(d_n, d_w, rtol=RTOL)
    np.testing.assert_allclose(W_n, W_w, rtol=1e-7, atol=1e-9 * np.max(np.abs(W_w)))
    gp = g.GaussianProcess(k, t=tt, diag=dfl, solver=solver)
    o_ld, o_q, _ = oracle.stream(0, scan, tt, x[1], diag=dfl)
    assert gp.log_likelihood(x[1]) == pytest.approx(oracle.log_likelihood_from_stream(o_ld, o_q, N), rel=RTOL)
    # shared light curve, and a sequence that is not positive definite
    kb = KernelBatch([k] * 4)
    geom = Geometry.shared_t(4, N)
    ld1, q1, s1 = solver.loglike(kb, geom, tt, x[0], dfl, flags=S.FLAG_SHARED_Y)
    o_ld, o_q, _ = oracle.stream(0, scan, tt, x[0], diag=dfl)
    assert s1.tolist() == [0] * 4 and ld1[3] == pytest.approx(o_ld, rel=RTOL) and q1[3] == pytest.approx(o_q, rel=RTOL)
    tbad = np.array([0.0, 0.0, 1.0, 2.0])
    _, _, _, st = batch.log_likelihood([g.SHOTerm(S0=1.0, w0=2.0, Q=3.0)] * 2, np.concatenate([tbad, tbad]),
                                       np.ones(8), np.concatenate([np.zeros(4), np.ones(4)]), lengths=[4, 4],
                                       solver=solver, return_parts=True)
    assert st.tolist() == [2, 0]


def test_rounded_phase_regressions(solver, solar_kernel):
    """Two cases the randomised stress tool found (profiles/r1_v6_stress.txt): the rows must be
    cos / sin of the ROUNDED phase fl(d t) -- an FMA contraction of d * t once made them drift."""
    # p-mode-only kernel, time stamps accumulated by cumsum (ulp-level irregular): fast producer path
    pm = g.StellarOscillatorKernel(terms=list(solar_kernel.term.terms[40:84]), delta=solar_kernel.delta)
    N = 2202
    t = np.cumsum(np.full(N, 8.64e-5))
    scan = pm.scan_coefficients()
    k0 = np.sum(scan[2]) + scan[6]
    _check(solver, pm, t, diag=5e-3 * k0, seed=3, rtol=1e-11)
    # JD-like absolute time stamps (phases of 1e9 rad) on narrow random kernels and on the solar kernel
    rng = np.random.default_rng(23)
    for nterm in (1, 3, 8):
        terms = [g.SHOTerm(S0=float(10 ** rng.uniform(0, 3)), w0=float(10 ** rng.uniform(2.5, 3.8)),
                           Q=float(10 ** rng.uniform(0, 2.5))) for _ in range(nterm)]
        k = g.StellarOscillatorKernel(terms=terms, delta=6e-5)
        sc = k.scan_coefficients()
        _check(solver, k, 2.1e5 + np.cumsum(np.full(300, 6e-5)), diag=1e-3 * (np.sum(sc[2]) + sc[6]), seed=nterm)
    _check(solver, solar_kernel, 2.1e5 + np.cumsum(np.full(1500, 6e-5)), diag=25.0, seed=9)


def test_kernels_wider_than_the_register_resident_scans(solver, solar_kernel):
    """``kernel + more terms`` (reference gadfly/core.py:405-427) beyond J = 176: the solar kernel plus
    10 / 60 extra SHO terms (J = 192 / 292) runs on the wide kernel (state in L2-resident scratch) --
    fused log-likelihood, fused sample and the stored-factor API (factor + sweeps with six terms per
    lane), against the oracle at the north star's 1e-9; a batch mixing narrow and wide kernels; and
    beyond GF_MAX_J_WIDE the library still refuses."""
    rng = np.random.default_rng(8)
    N = 300
    t = np.cumsum(6e-5 * (1 + 0.2 * rng.random(N)))
    for extra in (10, 60):
        terms = list(solar_kernel.term.terms) + [
            g.SHOTerm(S0=float(10 ** rng.uniform(-1, 1)), w0=float(10 ** rng.uniform(1, 4)), Q=float(10 ** rng.uniform(0, 2.5)))
            for _ in range(extra)]
        k = g.StellarOscillatorKernel(terms=terms, delta=solar_kernel.delta)
        assert k.J == 172 + 2 * extra
        scan = k.scan_coefficients()
        nrm = rng.standard_normal(N)
        x_ref, o_ld, st = oracle.stream(1, scan, t, nrm, diag=np.full(N, 9.0))
        assert st == 0
        kb = KernelBatch([k, solar_kernel])
        geom = Geometry.shared_t(2, N)
        diag = np.full(2 * N, 9.0)
        x, logdet, status = solver.sample(kb, geom, t, diag, normals=np.concatenate([nrm, nrm]))
        assert status.tolist() == [0, 0]
        assert np.max(np.abs(x[:N] - x_ref)) <= 1e-9 * np.max(np.abs(x_ref))
        assert logdet[0] == pytest.approx(o_ld, rel=1e-11)
        # the narrow kernel in the same batch
        x_sun = oracle.stream(1, solar_kernel.scan_coefficients(), t, nrm, diag=np.full(N, 9.0))[0]
        assert np.max(np.abs(x[N:] - x_sun)) <= 1e-9 * np.max(np.abs(x_sun))
        logdet2, quad, status = solver.loglike(kb, geom, t, np.concatenate([x_ref, x_sun]), diag)
        o_ld2, o_q, _ = oracle.stream(0, scan, t, x_ref, diag=np.full(N, 9.0))
        assert quad[0] == pytest.approx(o_q, rel=1e-9) and logdet2[0] == pytest.approx(o_ld2, rel=1e-11)
        # stored factor: compute + log_likelihood + dot_tril + apply_inverse
        gp = g.GaussianProcess(k, t=t, diag=np.full(N, 9.0), solver=solver)
        ogp = oracle.OracleGP(scan, t, diag=np.full(N, 9.0))
        assert gp.log_likelihood(x_ref) == pytest.approx(ogp.log_likelihood(x_ref), rel=1e-9)
        assert np.max(np.abs(gp.dot_tril(nrm) - ogp.dot_tril(nrm))) <= 1e-9 * np.max(np.abs(x_ref))
        ai, oai = gp.apply_inverse(x_ref), ogp.apply_inverse(x_ref)
        assert np.max(np.abs(ai - oai)) <= 1e-7 * np.max(np.abs(oai))
    too_many = list(solar_kernel.term.terms) * 3
    with pytest.raises(ValueError):
        KernelBatch([g.StellarOscillatorKernel(terms=too_many, delta=solar_kernel.delta)])


def test_blocked_scan_kernel_matches_oracle(solver, solar_kernel):
    """GF_FLAG_BLOCKED (csrc/scan_blk.cu): the 4-step blocked recurrence -- rank-4 updates, four
    matrix-vector products against the same state, one batch of dot products and a scalar 4 x 4
    recursion per hand-over -- is the arithmetic of the step-by-step recurrence up to summation order.
    Experimental (bulk-synchronous, slower than the default kernel): checked here
    against the oracle on a cadence with gaps (frame changes inside ring halves, blocks cut short),
    lengths that are not multiples of 4, a mixed-width batch, and a not-positive-definite case."""
    from gadfly_b200 import philox
    rng = np.random.default_rng(21)
    narrow = g.StellarOscillatorKernel(terms=list(solar_kernel.term.terms)[:30], delta=solar_kernel.delta)
    for N, kernels in ((1, [solar_kernel]), (6, [solar_kernel, narrow]), (1003, [solar_kernel, narrow, solar_kernel])):
        t = np.cumsum(rng.choice([6e-5, 6e-5, 6e-5, 1.2e-4, 3e-3, 0.4], N))
        B = len(kernels)
        kb = KernelBatch(kernels)
        geom = Geometry.shared_t(B, N)
        diag = np.full(B * N, 20.0 ** 2)
        y = rng.standard_normal(B * N) * 300.0
        ld, q, st = solver.loglike(kb, geom, t, y, diag=diag, flags=S.FLAG_BLOCKED)
        x, _, sx = solver.sample(kb, geom, t, diag=diag, seed=4, flags=S.FLAG_BLOCKED)
        for b, kern in enumerate(kernels):
            scan = kern.scan_coefficients()
            ref = oracle.stream(0, scan, t, y[b * N:(b + 1) * N], diag=diag[:N])
            assert st[b] == 0 and ref[2] == 0
            assert ld[b] == pytest.approx(ref[0], rel=1e-9)
            assert q[b] == pytest.approx(ref[1], rel=1e-9)
            xr = oracle.stream(1, scan, t, philox.normals(4, b, N), diag=diag[:N])[0]
            assert np.max(np.abs(x[b * N:(b + 1) * N] - xr)) <= 1e-9 * np.max(np.abs(xr))
    # first non-positive pivot: same row as the default kernel and the oracle report
    N = 700
    t = np.arange(N) * 6e-5
    bad = np.full(N, 25.0); bad[333:] = -4.0e5
    kb = KernelBatch([solar_kernel])
    y = rng.standard_normal(N) * 300.0
    _, _, st_blk = solver.loglike(kb, Geometry.shared_t(1, N), t, y, diag=bad, flags=S.FLAG_BLOCKED)
    _, _, st_def = solver.loglike(kb, Geometry.shared_t(1, N), t, y, diag=bad)
    assert st_blk[0] == st_def[0] == oracle.stream(0, solar_kernel.scan_coefficients(), t, y, diag=bad)[2] > 0
