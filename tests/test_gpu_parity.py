"""GPU parity tests: every kernel of the hot path, called through the C ABI
(``libgadfly_b200.so`` via ctypes), against the CPU oracle on the same inputs and against the
committed golden vectors.  Tolerances are the north star's: log-likelihood and samples rtol 1e-9
(identical normal draws), PSD rtol 1e-12."""
import numpy as np
import pytest

import gadfly_b200 as g
from gadfly_b200 import batch, philox, solver as S
from gadfly_b200.solver import Geometry, KernelBatch
import oracle
from oracle import terms_oracle as T
from conftest import golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9
CASES = ["sun_n384", "sun_ragged", "giant_n512", "gran_only"]


def _kernel_from_golden(gd):
    terms = [g.SHOTerm(S0=r[0], w0=r[1], Q=r[2]) for r in gd["sho"]]
    delta = float(gd["delta"])
    if delta > 0:
        return g.StellarOscillatorKernel(terms=terms, delta=delta)
    return g.TermSum(*terms)


def _maxrel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b))


@pytest.mark.parametrize("flags", [0, S.FLAG_REFERENCE_ORDER], ids=["fast", "reforder"])
@pytest.mark.parametrize("name", CASES)
def test_loglike_and_sample_vs_golden_and_oracle(solver, name, flags):
    gd = golden(f"dense_{name}.npz")
    k = _kernel_from_golden(gd)
    kb = KernelBatch([k])
    N = len(gd["t"])
    geom = Geometry.shared_t(1, N)
    logdet, quad, status = solver.loglike(kb, geom, gd["t"], gd["y"], gd["diag"], flags=flags)
    assert status[0] == 0
    ll = -0.5 * (quad[0] + logdet[0] + N * np.log(2 * np.pi))
    # dense Cholesky golden
    assert ll == pytest.approx(float(gd["loglike"]), rel=RTOL)
    assert logdet[0] == pytest.approx(float(gd["logdet"]), rel=RTOL)
    # oracle on the same inputs
    o_logdet, o_quad, o_status = oracle.stream(0, k.scan_coefficients(), gd["t"], gd["y"], diag=gd["diag"])
    assert logdet[0] == pytest.approx(o_logdet, rel=1e-11)
    assert quad[0] == pytest.approx(o_quad, rel=RTOL)
    x, ld2, status = solver.sample(kb, geom, gd["t"], gd["diag"], normals=gd["normals"], flags=flags)
    assert status[0] == 0 and ld2[0] == pytest.approx(o_logdet, rel=1e-11)
    o_x = oracle.stream(1, k.scan_coefficients(), gd["t"], gd["normals"], diag=gd["diag"])[0]
    assert _maxrel(x, o_x) <= RTOL
    assert _maxrel(x, gd["dot_tril"]) <= 5e-9      # dense Cholesky itself is ~cond*eps accurate


@pytest.mark.parametrize("flags", [0, S.FLAG_REFERENCE_ORDER], ids=["fast", "reforder"])
def test_batched_mixed_widths_shared_and_ragged(solver, solar_kernel, giant_kernel, flags):
    """Ragged batch: different kernels (J = 172, 124, 10, 2), different lengths incl. N = 1
    and an empty sequence, own time stamps per sequence."""
    rng = np.random.default_rng(5)
    gran = g.StellarOscillatorKernel(terms=list(solar_kernel.term.terms[:5]), delta=6e-5)
    one = g.SHOTerm(S0=3.0, w0=40.0, Q=2.5)
    kernels = [solar_kernel, giant_kernel, gran, one, solar_kernel, giant_kernel]
    lengths = [300, 257, 1000, 64, 1, 0]
    ts = [np.sort(rng.uniform(0, n * 9e-5, n)) for n in lengths]
    ts = [np.cumsum(np.maximum(np.diff(t, prepend=0.0), 6.1e-5)) for t in ts]
    ys = [rng.standard_normal(n) * 50 for n in lengths]
    # white-noise floor of 1e-4 k(0): with a negligible floor the giant's covariance is so
    # ill-conditioned (d_n / a_n ~ 1e-7) that one-ulp differences in sin/cos already move log det
    # by 1e-9 -- for the oracle as much as for the GPU path (see test_ill_conditioned_...)
    k0s = [np.sum(k.scan_coefficients()[0]) + np.sum(k.scan_coefficients()[2]) + k.scan_coefficients()[6]
           for k in kernels]
    diags = [np.full(n, 1e-4 * k0 * (1 + b)) for b, (n, k0) in enumerate(zip(lengths, k0s))]
    t, y, dg = map(np.concatenate, (ts, ys, diags))
    ll, logdet, quad, status = batch.log_likelihood(kernels, t, y, dg, lengths=lengths, solver=solver,
                                                    return_parts=True, flags=flags)
    assert status.tolist() == [0] * 6
    for b, k in enumerate(kernels):
        if lengths[b] == 0:
            assert logdet[b] == 0 and quad[b] == 0
            continue
        o_logdet, o_quad, _ = oracle.stream(0, k.scan_coefficients(), ts[b], ys[b], diag=diags[b])
        assert logdet[b] == pytest.approx(o_logdet, rel=RTOL), b
        assert quad[b] == pytest.approx(o_quad, rel=RTOL), b
    nrm = rng.standard_normal(len(t))
    rows, status = batch.sample(kernels, t, dg, lengths=lengths, normals=nrm, solver=solver,
                                subtract_mean=False, flags=flags)
    off = np.concatenate([[0], np.cumsum(lengths)])
    for b, k in enumerate(kernels):
        if lengths[b] == 0:
            continue
        o_x = oracle.stream(1, k.scan_coefficients(), ts[b], nrm[off[b]:off[b + 1]], diag=diags[b])[0]
        assert _maxrel(rows[b], o_x) <= RTOL, b


def test_many_sequences_more_than_sms(solver, giant_kernel):
    """More sequences than SMs, shared time grid: the work queue path."""
    B, N = 333, 96
    rng = np.random.default_rng(8)
    t = np.arange(N) * 1.2e-4
    scan = giant_kernel.scan_coefficients()
    k0 = np.sum(scan[2]) + scan[6]
    y = rng.standard_normal((B, N)) * np.sqrt(k0)
    dg = np.full((B, N), 1e-4 * k0)     # white-noise floor: keeps the problem well conditioned
    ll = batch.log_likelihood([giant_kernel] * B, t, y, dg, solver=solver)
    for b in [0, 1, 147, 148, 200, 332]:
        ld, q, st = oracle.stream(0, scan, t, y[b], diag=dg[b])
        assert ll[b] == pytest.approx(oracle.log_likelihood_from_stream(ld, q, N), rel=RTOL)


def test_fused_philox_sampling_matches_host_stream(solver, solar_kernel):
    B, N = 3, 700
    t = np.arange(N) * 6e-5
    x, status = batch.sample([solar_kernel] * B, t, seed=1234, seq0=10, solver=solver, subtract_mean=False)
    scan = solar_kernel.scan_coefficients()
    for b in range(B):
        n = philox.normals(1234, 10 + b, N)
        ref = oracle.stream(1, scan, t, n)[0]
        assert _maxrel(x[b], ref) <= RTOL
    # different sequences / seeds give different draws
    assert _maxrel(x[0], x[1]) > 0.1
    x2, _ = batch.sample([solar_kernel], t, seed=1235, seq0=10, solver=solver, subtract_mean=False)
    assert _maxrel(x2[0], x[0]) > 0.1


def test_non_positive_definite_status(solver):
    k = g.SHOTerm(S0=1.0, w0=2.0, Q=3.0)
    t = np.array([0.0, 0.0, 1.0, 2.0])
    ll, logdet, quad, status = batch.log_likelihood([k, k], np.concatenate([t, t + 0.0]),
                                                    np.ones(8), np.concatenate([np.zeros(4), np.ones(4)]),
                                                    lengths=[4, 4], solver=solver, return_parts=True)
    assert status[0] == 2 and status[1] == 0
    assert ll[0] == -np.inf and np.isfinite(ll[1])
    with pytest.raises(g.LinAlgError):
        batch.log_likelihood([k], t, np.ones(4), solver=solver, quiet=False)


@pytest.mark.parametrize("name", CASES)
def test_factor_and_sweeps_vs_oracle(solver, name):
    gd = golden(f"dense_{name}.npz")
    k = _kernel_from_golden(gd)
    kb = KernelBatch([k])
    N = len(gd["t"])
    geom = Geometry.shared_t(1, N)
    d, W, w_off, logdet, status = solver.factor(kb, geom, gd["t"], gd["diag"])
    ogp = oracle.OracleGP(k.scan_coefficients(), gd["t"], diag=gd["diag"])
    assert status[0] == 0
    np.testing.assert_allclose(d, ogp.d, rtol=RTOL)
    assert _maxrel(W.reshape(N, -1), ogp.W) <= RTOL
    rng = np.random.default_rng(2)
    Y = rng.standard_normal(N)
    for op, fn in enumerate([oracle.solve_lower, oracle.matmul_lower, oracle.solve_upper, oracle.matmul_upper]):
        Z = solver.sweep(op, kb, geom, w_off, gd["t"], W, Y)
        ref = fn(ogp.t, ogp.c, ogp.U, ogp.W, Y)
        assert _maxrel(Z, ref) <= RTOL, op


def test_gaussian_process_api_vs_golden(solver, solar_kernel):
    """The celerite2-style calls the tutorials make (reference docs/gadfly/start.rst:30-70)."""
    gd = golden("dense_sun_n384.npz")
    gp = g.GaussianProcess(solar_kernel, t=gd["t"], diag=gd["diag"], solver=solver)
    assert gp.log_likelihood(gd["y"]) == pytest.approx(float(gd["loglike"]), rel=RTOL)
    assert _maxrel(gp.dot_tril(gd["normals"]), gd["dot_tril"]) <= 5e-9
    assert _maxrel(gp.apply_inverse(gd["y"]), gd["apply_inverse"]) <= 1e-6
    Y2 = np.stack([gd["normals"], gd["y"]], axis=1)
    out = gp.dot_tril(Y2)
    assert out.shape == Y2.shape and _maxrel(out[:, 0], gd["dot_tril"]) <= 5e-9
    # sample(): numpy's global generator, as celerite2 draws it; mean subtracted (gadfly/gp.py:392)
    np.random.seed(42)
    x = gp.sample()
    np.random.seed(42)
    n = np.random.randn(len(gd["t"]))
    ref = oracle.OracleGP(solar_kernel.scan_coefficients(), gd["t"], diag=gd["diag"]).sample_from_normals(n)
    assert _maxrel(x, ref) <= RTOL and abs(x.mean()) < 1e-9 * np.abs(x).max()
    xs = gp.sample(size=3)
    assert xs.shape == (3, len(gd["t"]))
    q = gp.sample(return_quantity=True)
    assert q.unit == g.units.ppm
    # units at the boundary (reference gadfly/gp.py:61-126)
    t_days = gd["t"] / 0.0864 * g.units.d
    gp2 = g.GaussianProcess(solar_kernel, t=t_days, diag=gd["diag"], solver=solver)
    assert gp2.log_likelihood(gd["y"] * g.units.ppm) == pytest.approx(float(gd["loglike"]), rel=1e-7)
    with pytest.raises(ValueError):
        gp.log_likelihood(gd["y"][:-1])
    with pytest.raises(ValueError):
        gp.compute(gd["t"], yerr=np.ones(len(gd["t"])), diag=gd["diag"])
    pred = gp.predict(gd["y"])
    assert pred.shape == gd["y"].shape


def test_predict_at_new_times_vs_dense(solver, solar_kernel):
    """gp.predict(y, t=new times, return_var / return_cov) -- the gap-filling call of the
    reference's tutorial (docs/gadfly/synth.rst:193-201, gadfly/gp.py:243-306) -- against dense
    linear algebra on the kernel definition.  The mean uses the semiseparable form of the kernel
    for every lag (celerite2's general_matmul), the variance the exact k_delta(tau)."""
    from oracle import dense
    N = 300
    t = np.arange(N) * 1.8e-4                       # 3-min cadence, exposure 1 min
    diag = np.full(N, 20.0 ** 2)
    base = solar_kernel.base_coefficients()
    conv = tuple(solar_kernel.get_coefficients())
    delta = solar_kernel.delta
    K = dense.covariance_semiseparable(solar_kernel.scan_coefficients(), t, diag=diag)
    rng = np.random.default_rng(8)
    y = np.linalg.cholesky(K) @ rng.standard_normal(N)
    alpha = np.linalg.solve(K, y)
    ts = np.sort(np.concatenate([t[:-1] + 0.9e-4, [-5e-4, -1e-4, t[-1] + 2e-4, t[-1] + 0.5],
                                 t[[0, 17, N - 1]], t[[40, 41]] + 2e-5]))
    gp = g.GaussianProcess(solar_kernel, t=t, diag=diag, solver=solver)
    mu, var = gp.predict(y, t=ts, return_var=True)
    lag = ts[:, None] - t[None, :]
    mu_ref = T.get_value(conv, lag) @ alpha
    assert _maxrel(mu, mu_ref) <= RTOL
    Ks = T.get_value_convolved(base, delta, lag.ravel()).reshape(lag.shape)
    k0 = float(T.get_value_convolved(base, delta, np.zeros(1))[0])
    var_ref = k0 - np.einsum("ij,ji->i", Ks, np.linalg.solve(K, Ks.T))
    assert np.max(np.abs(var - var_ref)) <= 1e-8 * k0
    far = var[np.searchsorted(ts, t[-1] + 0.5)]          # 5.8 days after the last observation
    assert np.all(var > -1e-8 * k0) and 0.9 * k0 < far <= k0
    mu2, cov = gp.predict(y, t=ts[:60], return_cov=True)
    lag60 = ts[:60, None] - ts[None, :60]
    cov_ref = T.get_value_convolved(base, delta, lag60.ravel()).reshape(lag60.shape) \
        - Ks[:60] @ np.linalg.solve(K, Ks[:60].T)
    assert np.max(np.abs(cov - cov_ref)) <= 1e-8 * k0 and _maxrel(mu2, mu_ref[:60]) <= RTOL
    np.testing.assert_allclose(np.diag(cov), var[:60], atol=1e-8 * k0)
    # a different kernel for the prediction: the granulation component only
    gran = g.StellarOscillatorKernel(terms=list(solar_kernel.term.terms[:5]), delta=delta)
    mu_g = gp.predict(y, t=ts, kernel=gran)
    assert _maxrel(mu_g, T.get_value(tuple(gran.get_coefficients()), lag) @ alpha) <= RTOL
    # quantities in, quantities out
    q, qv = gp.predict(y * g.units.ppm, t=ts / 0.0864 * g.units.d, return_var=True, return_quantity=True)
    assert q.unit == g.units.ppm and _maxrel(q.value, mu_ref) <= 1e-7
    with pytest.raises(ValueError):
        gp.predict(y, t=ts[::-1])
    # observed times: unchanged celerite2 shortcut, and its variance
    mu0, var0 = gp.predict(y, return_var=True)
    assert _maxrel(mu0, y - diag * alpha) <= 1e-8
    Kc = dense.covariance(base, t, delta=delta)
    var0_ref = k0 - np.einsum("ij,ji->i", Kc, np.linalg.solve(K, Kc))
    assert np.max(np.abs(var0 - var0_ref)) <= 1e-7 * k0


def test_gp_not_positive_definite_raises_or_quiet(solver):
    k = g.SHOTerm(S0=1.0, w0=2.0, Q=3.0)
    t = np.array([0.0, 0.0, 1.0])
    with pytest.raises(g.LinAlgError):
        g.GaussianProcess(k, t=t, solver=solver)
    gp = g.GaussianProcess(k, solver=solver)
    gp.compute(t, quiet=True)
    assert gp.log_likelihood(np.ones(3)) == -np.inf


def test_psd_vs_oracle_and_reference_closed_form(solver, solar_kernel, giant_kernel):
    gd = golden("ref_sho_psd.npz")
    ref_kernel = g.TermSum(*[g.SHOTerm(S0=r[0], w0=r[1], Q=r[2]) for r in gd["params"]])
    got = ref_kernel.get_psd(gd["omega"])
    np.testing.assert_allclose(got, oracle.psd(ref_kernel.base_coefficients(), gd["omega"]), rtol=1e-12)
    far = gd["omega"] < 2 * np.pi * 800.0
    np.testing.assert_allclose(got[far], gd["psd_sum"][far], rtol=1e-12)   # reference _sho_psd
    np.testing.assert_allclose(got, gd["psd_sum"], rtol=1e-8)
    # exposure-integrated kernels, batched, dense grid incl. omega = 0
    omega = 2 * np.pi * np.concatenate([[0.0], np.linspace(0.01, 8333.0, 20001)])
    kb = KernelBatch([solar_kernel, giant_kernel])
    out = solver.psd(kb, omega)
    for b, k in enumerate([solar_kernel, giant_kernel]):
        np.testing.assert_allclose(out[b], oracle.psd(k.base_coefficients(), omega, k.delta), rtol=1e-12)
    assert out.shape == (2, len(omega))
    np.testing.assert_allclose(g.kernel_psd(solar_kernel, omega / (2 * np.pi), solver=solver), out[0], rtol=0)


def test_device_pointers_and_async(solver, solar_kernel):
    """Inputs and outputs already resident in HBM (torch tensors) give the same numbers."""
    import torch
    B, N = 4, 256
    rng = np.random.default_rng(3)
    t = np.arange(N) * 6e-5
    y = rng.standard_normal((B, N)) * 300
    kb = KernelBatch([solar_kernel] * B)
    geom = Geometry.shared_t(B, N)
    ref = solver.loglike(kb, geom, t, y)
    dev = torch.device('cuda', solver.device)
    td, yd = torch.from_numpy(t).to(dev), torch.from_numpy(y).to(dev)
    ld = torch.empty(B, dtype=torch.float64, device=dev)
    qd = torch.empty(B, dtype=torch.float64, device=dev)
    sd = torch.empty(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    solver.loglike(kb, geom, td, yd, logdet=ld, quad=qd, status=sd, flags=S.FLAG_ASYNC)
    solver.synchronize()
    np.testing.assert_array_equal(ld.cpu().numpy(), ref[0])
    np.testing.assert_array_equal(qd.cpu().numpy(), ref[1])
    assert solver.last_kernel_ms > 0 and solver.launch_count > 0


def test_psd_device_tensors(solver, solar_kernel, giant_kernel):
    """Frequency grid and output resident in HBM: same numbers as the host-array call."""
    import torch
    dev = torch.device('cuda', solver.device)
    kb = KernelBatch([solar_kernel, giant_kernel])
    omega = 2 * np.pi * np.linspace(0.5, 6000.0, 5000)
    ref = solver.psd(kb, omega)
    od = torch.from_numpy(omega).to(dev)
    out = torch.empty(2 * len(omega), dtype=torch.float64, device=dev)
    solver.psd(kb, od, out=out)
    np.testing.assert_array_equal(out.cpu().numpy().reshape(2, -1), ref)


def test_batched_feeder_end_to_end(solver):
    """Stars -> KernelBatch.for_stars (batched feeder) -> fused log-likelihood, against the
    per-star kernel objects through the same kernel and against the CPU oracle."""
    import warnings
    stars = [(1.0, 1.0, 5777.0, 1.0), (1.1, 1.6, 6100.0, 3.2), (0.9, 1.4, 5500.0, 2.2), (1.3, 2.5, 5900.0, 6.0)]
    M, R, Tt, L = (np.array(x) for x in zip(*stars))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kernels = [g.StellarOscillatorKernel(
            g.Hyperparameters.for_star(m, r, t, l, bandpass='SOHO VIRGO', quiet=True), texp=1 * g.units.min)
            for m, r, t, l in stars]
    kb_ref, kb = KernelBatch(kernels), KernelBatch.for_stars(M, R, Tt, L)
    for b in (kb_ref, kb):
        b.ddiag = b.ddiag + 30.0 ** 2          # yerr = 30 ppm
    N = 1500
    t = np.arange(N) * 6e-5
    y = np.random.default_rng(11).standard_normal((len(stars), N)) * 200
    geom = Geometry.shared_t(len(stars), N)
    ld0, q0, s0 = solver.loglike(kb_ref, geom, t, y)
    ld1, q1, s1 = solver.loglike(kb, geom, t, y)
    assert s0.tolist() == [0] * 4 and s1.tolist() == [0] * 4
    np.testing.assert_allclose(ld1, ld0, rtol=1e-9)
    np.testing.assert_allclose(q1, q0, rtol=1e-8)
    for b, k in enumerate(kernels):
        sc = list(k.scan_coefficients())
        sc[-1] += 30.0 ** 2
        logdet, quad, status = oracle.stream(0, tuple(sc), t, y[b])
        assert status == 0
        assert ld0[b] == pytest.approx(logdet, rel=RTOL) and q0[b] == pytest.approx(quad, rel=RTOL)


def test_round_trip_sample_psd_statistics(solver, solar_kernel):
    """The reference's only hot-path test (gadfly/tests/test_core.py:17-49): samples drawn from
    the kernel have the kernel's power spectrum (binned FFT PSD within 5 sigma, 3-1000 uHz)."""
    np.random.seed(42)
    N = 100_000
    t_days = np.linspace(0, 100, N)
    gp = g.GaussianProcess(solar_kernel, t=t_days * g.units.d, solver=solver)
    for _ in range(3):
        flux = gp.sample()
        ps = g.PowerSpectrum.from_light_curve(t_days, flux).bin(15)
        model = solar_kernel.get_psd(2 * np.pi * ps.frequency)
        ok = (ps.frequency < 1e3) & (ps.frequency > 3)
        dev = np.abs((model[ok] - ps.power[ok]) / np.nanmax(ps.error))
        assert np.nanmax(dev) < 5


def test_large_property_checks(solver, solar_kernel):
    """Size-independent properties at a size the oracle does not finish quickly:
    L^-1 (L n) = n through the fused sample + loglike kernels, and log-det additivity of a
    shared factor across replicas."""
    N, B = 20000, 4
    t = np.arange(N) * 6e-5
    x, status = batch.sample([solar_kernel] * B, t, seed=99, solver=solver, subtract_mean=False)
    assert status.tolist() == [0] * B
    ll, logdet, quad, status = batch.log_likelihood([solar_kernel] * B, t, x, solver=solver, return_parts=True)
    # quad = |L^-1 x|^2_D^-1 = sum n^2  for x = L sqrt(D) n
    for b in range(B):
        n = philox.normals(99, b, N)
        assert quad[b] == pytest.approx(np.sum(n * n), rel=1e-8)
    assert np.ptp(logdet) <= 1e-12 * abs(logdet[0])


# ---- ground truth from the kernel DEFINITION in extended precision (tools/make_golden_definition.py) ----
# (log-det, log-likelihood, samples / K^-1 y): the north star's 1e-9 on the three well-conditioned
# cases -- against the DEFINITION, not against the oracle; at the reference's default 1-min exposure
# celerite2's own FP64 closed form for the exposure-integrated coefficients cancels (SURVEY.md 0.6:
# 6e-9 on the slowest granulation term), which bounds ANY implementation that follows it: 2.8e-9 on
# the samples for the CPU oracle too (tests/test_oracle.py), logL stays at 5e-12
DEF_CASES = [("def_solar_200s", RTOL, RTOL, RTOL), ("def_subgiant", RTOL, RTOL, RTOL),
             ("def_giant", RTOL, RTOL, RTOL), ("def_solar_sc", RTOL, RTOL, 2e-8)]


@pytest.mark.parametrize("name,tol_ld,tol_ll,tol_x", DEF_CASES)
def test_gp_facade_matches_definition_golden(solver, name, tol_ld, tol_ll, tol_x):
    """(S0, w0, Q), t, y in -> SHOTerm / TermConvolution coefficients (host), rows + factor + sweeps
    (CUDA) -> logL, samples, K^-1 y, against dense longdouble linear algebra on the kernel definition."""
    gd = golden(name + ".npz")
    terms = [g.SHOTerm(S0=a, w0=b, Q=c) for a, b, c in zip(gd["S0"], gd["w0"], gd["Q"])]
    k = g.StellarOscillatorKernel(terms=terms, delta=float(gd["delta"]))
    kb = KernelBatch([k])
    N = len(gd["t"])
    geom = Geometry.shared_t(1, N)
    diag = gd["diag"] if np.any(gd["diag"]) else None
    logdet, quad, status = solver.loglike(kb, geom, gd["t"], gd["y"], diag)
    assert status[0] == 0
    assert logdet[0] == pytest.approx(float(gd["logdet"]), rel=tol_ld)
    ll = -0.5 * (quad[0] + logdet[0] + N * np.log(2 * np.pi))
    assert ll == pytest.approx(float(gd["logl"]), rel=tol_ll)
    x, _, status = solver.sample(kb, geom, gd["t"], diag, normals=gd["normals"])
    assert status[0] == 0 and _maxrel(x, gd["x"]) <= tol_x
    # the user-facing class: compute + log_likelihood + apply_inverse
    gp = g.GaussianProcess(k, t=gd["t"], diag=diag)
    assert gp.log_likelihood(gd["y"]) == pytest.approx(float(gd["logl"]), rel=tol_ll)
    assert _maxrel(gp.apply_inverse(gd["y"]), gd["alpha"]) <= 10 * tol_x


# ---- k right-hand sides per sequence on ONE factor (gf_sample_multi / gf_loglike_multi) -------------
def test_multi_rhs_shares_one_factor(solver, solar_kernel, giant_kernel):
    """SURVEY.md 8b's k: realisations / data vectors sharing a factor (celerite2 sample(size=k),
    reference gadfly/gp.py:372-395).  Ragged batch of two kernels, k = 5: every realisation equals the
    fused single-realisation kernel's and the oracle's for the same normal draws; Philox realisation
    index seq0 + b k + r; log-likelihood pieces of k data vectors."""
    rng = np.random.default_rng(17)
    kernels = [solar_kernel, giant_kernel]
    lengths = [700, 333]
    k = 5
    ts = [np.cumsum(6e-5 * (1 + 0.3 * rng.random(n))) for n in lengths]
    t = np.concatenate(ts)
    kb = KernelBatch(kernels)
    geom = Geometry.ragged(lengths)
    diag = np.concatenate([np.full(n, 25.0) for n in lengths])
    nrm = [rng.standard_normal((k, n)) for n in lengths]
    x, logdet, status = solver.sample_multi(kb, geom, t, k, diag=diag, normals=np.concatenate([a.ravel() for a in nrm]))
    assert np.all(status == 0)
    off = 0
    ys = []
    for b, (kern, n) in enumerate(zip(kernels, lengths)):
        scan = kern.scan_coefficients()
        xb = x[off:off + k * n].reshape(k, n)
        off += k * n
        for r in range(k):
            ref, o_ld, st = oracle.stream(1, scan, ts[b], nrm[b][r], diag=np.full(n, 25.0))
            assert st == 0 and _maxrel(xb[r], ref) <= RTOL
            assert logdet[b] == pytest.approx(o_ld, rel=1e-11)
        ys.append(xb)
    # Philox draws: realisation (b, r) is sequence seq0 + b k + r of the host generator
    xp, _, _ = solver.sample_multi(kb, geom, t, k, diag=diag, seed=5, seq0=100)
    ref = oracle.stream(1, kernels[1].scan_coefficients(), ts[1], philox.normals(5, 100 + 1 * k + 3, lengths[1]),
                        diag=np.full(lengths[1], 25.0))[0]
    got = xp[k * lengths[0] + 3 * lengths[1]:k * lengths[0] + 4 * lengths[1]]
    assert _maxrel(got, ref) <= RTOL
    # k data vectors per sequence
    y = np.concatenate([a.ravel() for a in ys])
    logdet2, quad, status = solver.loglike_multi(kb, geom, t, y, k, diag=diag)
    assert np.all(status == 0)
    for b, (kern, n) in enumerate(zip(kernels, lengths)):
        for r in range(k):
            o_ld, o_q, st = oracle.stream(0, kern.scan_coefficients(), ts[b], ys[b][r], diag=np.full(n, 25.0))
            assert quad[b * k + r] == pytest.approx(o_q, rel=RTOL)
            assert logdet2[b] == pytest.approx(o_ld, rel=1e-11)
    # the batch facade: [B, size, N]
    xs, st = batch.sample([solar_kernel] * 3, np.arange(256) * 6e-5, size=4, seed=9, solver=solver, subtract_mean=False)
    assert xs.shape == (3, 4, 256) and np.all(st == 0)
    one, _ = batch.sample([solar_kernel], np.arange(256) * 6e-5, seed=9, seq0=2 * 4 + 1, solver=solver, subtract_mean=False)
    assert _maxrel(xs[2, 1], one[0]) <= RTOL


# ---- the observed power spectrum and its binning on the device ----------------------
def test_observed_power_spectrum_and_binning_on_device(solver):
    """gf_power_spectrum_batched / gf_bin_power_batched against the host restatement of the reference's
    ``PowerSpectrum._fft`` (gadfly/psd.py:566-587: numpy rfft, norm d / sqrt(2 pi) / N) and
    ``bin_power_spectrum`` (gadfly/psd.py:229-297), odd and even lengths, linear and log bins."""
    from gadfly_b200 import psd as P
    rng = np.random.default_rng(4)
    for N in (4096, 10007):
        flux = rng.standard_normal((3, N)) * 100.0 + 5 * np.sin(np.arange(N) * 0.01)[None, :]
        d_days = 1.0 / 1440.0
        freq, power, norm = P.power_spectra(flux, d_days, solver=solver)
        for b in range(3):
            ref = g.PowerSpectrum.from_light_curve(np.arange(N) * d_days, flux[b])
            np.testing.assert_allclose(freq, ref.frequency, rtol=1e-14)
            assert np.max(np.abs(power[b] - ref.power)) <= 1e-12 * np.max(ref.power)
            assert norm == pytest.approx(ref.norm, rel=1e-15)
        for log, bins in ((True, 15), (False, 40)):
            fb, stat, err = P.bin_power_spectra(freq, power, bins=bins, log=log, solver=solver)
            for b in range(3):
                ref = P.bin_power_spectrum(g.PowerSpectrum(freq, power[b]), bins=bins, log=log)
                np.testing.assert_allclose(fb, ref.frequency, rtol=1e-13)
                np.testing.assert_allclose(stat[b], ref.power, rtol=1e-11, equal_nan=True)
                np.testing.assert_allclose(err[b], ref.error, rtol=1e-9, equal_nan=True)


def test_round_trip_stays_on_the_device(solver, solar_kernel):
    """The reference's round trip (gadfly/tests/test_core.py:17-49) without leaving HBM: B Philox draws
    from the fused sample kernel (CUDA tensor out) -> FFT power spectra -> 15 log bins -> the kernel PSD
    at the bin centres, all through device pointers; binned spectra within 5 sigma of the model for
    3 < f < 1000 uHz, as the reference asserts."""
    import torch
    from gadfly_b200 import psd as P
    dev = torch.device("cuda", solver.device)
    B, N = 6, 100_000
    t_days = np.linspace(0, 100, N)
    t = torch.as_tensor(t_days * 0.0864, device=dev)             # days -> 1/uHz (reference gadfly/gp.py:79-86)
    x = torch.empty(B * N, dtype=torch.float64, device=dev)
    kb = KernelBatch([solar_kernel] * B)
    _, _, status = solver.sample(kb, Geometry.shared_t(B, N), t, seed=42, out=x)
    assert np.all(status == 0)
    flux = x.view(B, N)
    flux = flux - flux.mean(dim=1, keepdim=True)                  # the reference's mean subtraction (gp.py:392)
    freq, power, _ = P.power_spectra(flux.contiguous(), float(t_days[1] - t_days[0]), solver=solver)
    assert power.is_cuda
    fb, stat, err = P.bin_power_spectra(freq, power, bins=15, solver=solver)
    assert stat.is_cuda
    model = solver.psd(KernelBatch([solar_kernel]), torch.as_tensor(2 * np.pi * fb, device=dev),
                       out=torch.empty((1, len(fb)), dtype=torch.float64, device=dev))
    ok = torch.as_tensor((fb < 1e3) & (fb > 3), device=dev)
    dev_sigma = ((model[0][None, :] - stat).abs() / torch.nan_to_num(err, nan=0.0).max(dim=1, keepdim=True).values)[:, ok]
    assert float(dev_sigma.max()) < 5


# ---- gradients of log L with respect to (S0, w0, Q) ---------------------------------
def test_log_likelihood_gradient_vs_definition(solver):
    """batch.log_likelihood_gradient (4 P + 1 perturbed kernels scanned in one batched launch against the
    one light curve) against tests/golden/def_grad.npz: d log L / d ln p by central differences of the
    longdouble kernel DEFINITION (tools/make_golden_definition.py), 6 terms x (S0, w0, Q)."""
    gd = golden("def_grad.npz")
    grad, ll = batch.log_likelihood_gradient(gd["S0"], gd["w0"], gd["Q"], float(gd["delta"]), gd["t"], gd["y"],
                                             diag=gd["diag"], solver=solver, return_value=True)
    assert ll == pytest.approx(float(gd["logl"]), rel=RTOL)
    ref = gd["grad"]
    assert grad.shape == ref.shape == (3, 6)
    assert np.max(np.abs(grad - ref)) <= 1e-6 * np.linalg.norm(ref)
    # every component that matters individually, too
    big = np.abs(ref) > 1e-3 * np.max(np.abs(ref))
    np.testing.assert_allclose(grad[big], ref[big], rtol=1e-5)
    # a subset of the parameters
    g_w0 = batch.log_likelihood_gradient(gd["S0"], gd["w0"], gd["Q"], float(gd["delta"]), gd["t"], gd["y"],
                                         diag=gd["diag"], wrt=("w0",), solver=solver)
    np.testing.assert_allclose(g_w0[0], grad[1], rtol=1e-12)


def test_async_calls_with_tickets(solver, solar_kernel):
    """``log_likelihood(wait=False)``: several calls queued back to back on one handle (host buffers,
    ping-pong staging), each result read through its own completion ticket, equal to the blocking call."""
    rng = np.random.default_rng(3)
    N, B = 3000, 5
    t = np.arange(N) * 6e-5
    ys = [rng.standard_normal((B, N)) * 280.0 for _ in range(4)]
    ref = [batch.log_likelihood([solar_kernel] * B, t, y, solver=solver) for y in ys]
    pend = [batch.log_likelihood([solar_kernel] * B, t, y, solver=solver, wait=False) for y in ys]
    for p, r in zip(reversed(pend), reversed(ref)):
        np.testing.assert_array_equal(p.result(), r)


def test_async_call_does_not_wait_for_pageable_outputs(solver, solar_kernel):
    """GF_FLAG_ASYNC with ordinary (pageable) numpy outputs: the call returns while its kernel is
    still running -- a direct device-to-host copy into pageable memory would block until the kernel
    has finished -- and the outputs are delivered by ``wait`` / ``synchronize`` through the handle's
    pinned ring, equal to the blocking call's, also when the caller drops the returned arrays."""
    import time
    N, B = 60000, 148
    t = np.arange(N) * 6e-5
    kb = KernelBatch([solar_kernel] * B)
    geom = Geometry.shared_t(B, N)
    import torch
    x_dev = torch.empty(B * N, dtype=torch.float64, device="cuda")
    _, ld_ref, st_ref = solver.sample(kb, geom, t, seed=9, out=x_dev)
    t0 = time.perf_counter()
    solver.sample(kb, geom, t, seed=9, out=x_dev)
    blocking = time.perf_counter() - t0
    assert blocking > 0.02                                   # ~60 ms of kernel
    t0 = time.perf_counter()
    _, ld, st = solver.sample(kb, geom, t, seed=9, out=x_dev, flags=S.FLAG_ASYNC)
    issued = time.perf_counter() - t0
    solver.sample(kb, geom, t, seed=10, out=x_dev, flags=S.FLAG_ASYNC)       # returned arrays dropped
    tk = solver.ticket()
    _, ld3, st3 = solver.sample(kb, geom, t, seed=11, out=x_dev, flags=S.FLAG_ASYNC)
    assert issued < 0.5 * blocking, (issued, blocking)
    solver.wait(tk)
    np.testing.assert_array_equal(ld, ld_ref)
    np.testing.assert_array_equal(st, st_ref)
    solver.synchronize()
    np.testing.assert_array_equal(ld3, ld_ref)               # same kernel, other draws: same log det


def test_fit_with_device_gradients(solver):
    """What the gradients are for (the reference fits its solar kernel with celerite2.jax + BFGS,
    notebooks/virgo_lc.ipynb:48): maximise log L over ln(S0, w0, Q) of a 4-term kernel with scipy's
    L-BFGS-B, value and gradient from ``batch.log_likelihood_gradient`` (one batched launch per
    evaluation).  Starting 20-30 % off, the fit climbs to within a few units of the log-likelihood at
    the parameters the data were drawn from (or above it)."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(5)
    S0 = np.array([600.0, 8.0, 2.0e-3, 3.0e-3])
    w0 = np.array([9.0, 150.0, 17500.0, 19800.0])
    Q = np.array([0.6, 0.7, 300.0, 450.0])
    delta, N = 2e-4, 3000
    t = np.arange(N) * 2.5e-4
    diag = np.full(N, 9.0)
    truth = g.StellarOscillatorKernel(terms=[g.SHOTerm(S0=a, w0=b, Q=c) for a, b, c in zip(S0, w0, Q)], delta=delta)
    y, st = batch.sample([truth], t, diag, seed=77, solver=solver, subtract_mean=False)
    y = y[0] + 3.0 * rng.standard_normal(N)
    assert st[0] == 0
    ll_true = batch.log_likelihood([truth], t, y[None, :], diag, solver=solver)[0]
    x_true = np.log(np.concatenate([S0, w0, Q]))
    x0 = x_true + np.concatenate([0.3 * rng.uniform(-1, 1, 4), [0.2, -0.2, 2e-3, -2e-3], 0.3 * rng.uniform(-1, 1, 4)])
    x0[8] = max(x0[8], np.log(0.55))      # granulation terms stay underdamped (Q > 0.5)
    x0[9] = max(x0[9], np.log(0.55))
    calls = []

    def objective(x):
        p = np.exp(x)
        grad, ll = batch.log_likelihood_gradient(p[:4], p[4:8], p[8:], delta, t, y, diag=diag, solver=solver,
                                                 return_value=True)
        calls.append(ll)
        return -ll, -grad.ravel()

    bounds = [(None, None)] * 8 + [(np.log(0.51), None)] * 4
    res = minimize(objective, x0, jac=True, method="L-BFGS-B", bounds=bounds, options=dict(maxiter=60))
    assert calls[0] < ll_true - 20.0            # the start is clearly worse than the truth
    assert -res.fun > ll_true - 5.0             # the fit is as good as the truth (12 parameters)
    assert np.all(np.abs(res.x[4:8] - x_true[4:8]) < np.array([0.5, 0.5, 2e-3, 2e-3]))   # frequencies recovered


def test_device_feeder_matches_host_feeder(solver):
    """(f2) csrc/feed.cu through gf_feed_stars against gadfly_b200/feeder.py (itself pinned to the
    per-star ``Hyperparameters.for_star`` path and the mpmath fixture on the CPU side): 512
    Kepler-like stars + the documentation's stars.  Same terms kept; (S0 w0 Q), a, b, Q to 1e-12
    (w0, c, d on the scale of the star's highest frequency: the scaled mode frequencies are a
    difference that can land near zero).  The exposure transform is celerite2's closed form, whose
    cosh(w) cos(.) - 1 loses 1/|w|^2 (w = (c + i d) Delta) in ANY FP64 evaluation, so a', b' are
    compared with that factor and on the scale of k(0); Delta-diag and log-likelihoods at a 10-min
    exposure on stars where every term is resolved (|w| >= 1e-3)."""
    from gadfly_b200 import feeder, workloads
    M, R, T, L = workloads.kepler_like_stars(512, 4)
    doc = np.array([(1.0, 1.0, 5777.0, 1.0), (0.9, 10.0, 4919.0, 52.3), (1.32, 11.26, 4923.0, 66.8),
                    (1.1, 1.6, 6100.0, 3.2), (2.0, 20.7, 4364.0, 146.0)])
    M, R, T, L = (np.concatenate([doc[:, k], x]) for k, x in enumerate((M, R, T, L)))
    hpb = feeder.for_stars(M, R, T, L)
    ref = feeder.kernel_batch_from_sho(hpb, 6e-5)
    got, hp = feeder.kernel_batch_for_stars_device(solver, M, R, T, L, texp_s=60.0, return_hyperparameters=True)
    assert np.array_equal(got.j_off, ref.j_off)
    assert got.j_off[1] == 86 and got.j_off[2] - got.j_off[1] == 62           # Sun, KIC 9333184
    widths = np.diff(ref.j_off)
    star = np.repeat(np.arange(ref.B), widths)
    w0max = np.maximum.reduceat(hpb.w0, ref.j_off[:-1])[star]
    assert np.max(np.abs(hp.w0 - hpb.w0) / w0max) < 1e-13
    np.testing.assert_allclose(hp.Q, hpb.Q, rtol=1e-13)
    np.testing.assert_allclose(hp.S0 * hp.w0, hpb.S0 * hpb.w0, rtol=1e-12)
    np.testing.assert_allclose(got.base[:, :2], ref.base[:, :2], rtol=1e-12)
    assert np.max(np.abs(got.base[:, 2:] - ref.base[:, 2:]) / w0max[:, None]) < 1e-13
    k0 = np.add.reduceat(np.abs(ref.coef[:, 0]), ref.j_off[:-1])
    mag = np.hypot(ref.coef[:, 0], ref.coef[:, 1])
    w = np.hypot(ref.coef[:, 2], ref.coef[:, 3]) * 6e-5
    err = np.max(np.abs(got.coef[:, :2] - ref.coef[:, :2]), axis=1)
    assert np.all(err <= 1e-13 * mag / np.minimum(w * w, 1.0) + 1e-11 * k0[star])
    # Delta-diag and the log-likelihoods downstream, at a 10-min exposure, on the stars whose slowest
    # term is still resolved (|w| >= 1e-3: the closed form keeps >= 10 digits)
    ref10 = feeder.kernel_batch_from_sho(hpb, 6e-4)
    got10 = feeder.kernel_batch_for_stars_device(solver, M, R, T, L, texp_s=600.0)
    w10 = np.hypot(ref10.coef[:, 2], ref10.coef[:, 3]) * 6e-4
    resolved = np.minimum.reduceat(w10, ref10.j_off[:-1]) >= 1e-3
    assert resolved.sum() > 50
    k10 = np.add.reduceat(np.abs(ref10.coef[:, 0]), ref10.j_off[:-1])
    assert np.max(np.abs(got10.ddiag - ref10.ddiag)[resolved] / k10[resolved]) < 1e-9
    idx = np.nonzero(resolved)[0][:6]
    N = 4000
    t = np.arange(N) * 6e-4
    y = np.random.default_rng(8).standard_normal((len(idx), N)) * np.sqrt(k10[idx])[:, None]
    ll_ref = batch.log_likelihood(ref10.take(idx), t, y, solver=solver)
    ll_got = batch.log_likelihood(got10.take(idx), t, y, solver=solver)
    np.testing.assert_allclose(ll_got, ll_ref, rtol=1e-9)
    # KernelBatch.for_stars(solver=...) is the same call
    kb = KernelBatch.for_stars(M[:5], R[:5], T[:5], L[:5], texp_s=60.0, solver=solver)
    np.testing.assert_array_equal(kb.coef, got.take(np.arange(5)).coef)


def test_device_bandpass_amplitude(solver):
    """gf_bandpass_amplitude (Morris+ 2020 Eqn 11, reference gadfly/scale.py:635-729) against the
    host quadrature on the same 10^4-point wavelength grid, and the feeder with a tabulated bandpass."""
    from gadfly_b200 import feeder, scale

    class Band:
        wavelength = np.linspace(0.4, 0.9, 200)
        transmittance = np.exp(-0.5 * ((np.linspace(0.4, 0.9, 200) - 0.65) / 0.1) ** 2)

    filt = g.Filter(Band)
    T = np.array([3900.0, 4400.0, 5000.0, 5777.0, 6500.0, 7200.0])
    wl, tr = scale.bandpass_grid(filt)
    got = solver.bandpass_amplitude(T, wl, tr)
    ref = np.array([scale.amplitude_with_wavelength(filt, x) for x in T])
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    M, R, L = np.ones(6), np.ones(6), (T / 5777.0) ** 4
    kd = feeder.kernel_batch_for_stars_device(solver, M, R, T, L, bandpass=filt)
    kh = feeder.kernel_batch_for_stars(M, R, T, L, bandpass=filt)
    assert np.array_equal(kd.j_off, kh.j_off)
    np.testing.assert_allclose(kd.base[:, 0], kh.base[:, 0], rtol=1e-11)


def test_device_sho_to_coefficients(solver):
    """gf_feed_sho: (S0, w0, Q) in CSR layout -> coefficients, against feeder.kernel_batch_from_sho on
    a slice of the cfg4 lattice (solar kernel, J = 172) and on a ragged batch; a', b' within the
    rounding of the closed form (1/|w|^2, see test_device_feeder_matches_host_feeder), Delta-diag on
    the scale of k(0); overdamped terms are refused like on the host."""
    from gadfly_b200 import feeder, workloads
    ref, _ = workloads.lattice_batch(64, 3)
    got, _ = workloads.lattice_batch(64, 3, solver=solver)
    assert np.array_equal(got.j_off, ref.j_off)
    np.testing.assert_allclose(got.base, ref.base, rtol=1e-13)
    star = np.repeat(np.arange(ref.B), np.diff(ref.j_off))
    k0 = np.add.reduceat(np.abs(ref.coef[:, 0]), ref.j_off[:-1])
    mag = np.hypot(ref.coef[:, 0], ref.coef[:, 1])
    w = np.hypot(ref.coef[:, 2], ref.coef[:, 3]) * 6e-5
    err = np.max(np.abs(got.coef[:, :2] - ref.coef[:, :2]), axis=1)
    assert np.all(err <= 1e-13 * mag / np.minimum(w * w, 1.0) + 1e-11 * k0[star])
    np.testing.assert_array_equal(got.coef[:, 2:], got.base[:, 2:])
    assert np.max(np.abs(got.ddiag - ref.ddiag) / k0) < 1e-13 / np.min(w) ** 2
    # ragged widths, per-kernel exposure times
    rng = np.random.default_rng(12)
    widths = rng.integers(1, 100, 37)
    j_off = np.concatenate([[0], np.cumsum(widths)])
    n = int(j_off[-1])
    hpb = feeder.HyperparameterBatch(rng.uniform(0.1, 100, n), rng.uniform(50.0, 3e4, n), rng.uniform(0.5, 900, n), j_off)
    delta = rng.uniform(2e-4, 1e-3, 37)
    r2 = feeder.kernel_batch_from_sho(hpb, delta)
    g2 = feeder.kernel_batch_from_sho(hpb, delta, solver=solver)
    np.testing.assert_allclose(g2.base, r2.base, rtol=1e-13)
    k2 = np.add.reduceat(np.abs(r2.coef[:, 0]), j_off[:-1])
    s2 = np.repeat(np.arange(37), widths)
    assert np.max(np.abs(g2.coef[:, :2] - r2.coef[:, :2]) / k2[s2, None]) < 1e-10
    assert np.max(np.abs(g2.ddiag - r2.ddiag) / k2) < 1e-9
    t = np.arange(1500) * 4e-4
    y = rng.standard_normal((37, 1500)) * np.sqrt(k2)[:, None]
    np.testing.assert_allclose(batch.log_likelihood(g2, t, y, solver=solver),
                               batch.log_likelihood(r2, t, y, solver=solver), rtol=1e-9)
    bad = feeder.HyperparameterBatch([1.0, 1.0], [100.0, 200.0], [0.7, 0.4], [0, 2])
    with pytest.raises(ValueError):
        feeder.kernel_batch_from_sho(bad, 6e-5, solver=solver)
    with pytest.raises(ValueError):
        solver.feed_sho(np.array([0, 2]), [1.0, 1.0], [100.0, 200.0], [0.7, 0.4], 6e-5)
