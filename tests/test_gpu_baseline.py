"""GPU parity at the sizes BASELINE.json names, through the C ABI: slices of the
configurations the CPU oracle can finish in seconds, plus size-independent properties at full length,
and the randomised stress run as a test.  Tolerances: the north star's 1e-9 (logL pieces, samples) and
1e-12 (PSD)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import gadfly_b200 as g
from gadfly_b200 import batch, philox, solver as S, workloads
from gadfly_b200.solver import Geometry, KernelBatch
import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-9


def _oracle_batch(kb, idx, t, y, ddiag):
    """Fused log-likelihood streams of the oracle for the sequences ``idx`` (all host cores)."""
    sub = kb.take(idx)
    n = len(t)
    n_off = np.arange(len(idx) + 1) * n
    yy = np.concatenate([y[i] for i in idx])
    out, _, status = oracle.stream_batch(0, n_off, np.zeros(len(idx), dtype=np.int64), sub.j_off, t, yy, ddiag[idx],
                                         *[np.ascontiguousarray(sub.coef[:, k]) for k in range(4)],
                                         nthreads=max(1, len(os.sched_getaffinity(0))), fast=False)
    return out, status


def test_cfg2_kepler_population_65536_points(solver):
    """BASELINE configs[1]: Kepler-like stars (Huber-2011 table, J = 88 ... 172), 65 536-point 1-min
    cadence, yerr = 50 ppm, batched log-likelihood: 64 stars on the GPU, 8 of them through the oracle."""
    kb, _ = workloads.kepler_like_batch(64, 1)
    kb.ddiag = kb.ddiag + 50.0 ** 2
    B, N = kb.B, 65536
    t = np.arange(N) * 6e-5
    rng = np.random.default_rng(2)
    k0 = np.array([np.sum(kb.coef[kb.j_off[b]:kb.j_off[b + 1], 0]) + kb.ddiag[b] for b in range(B)])
    y = rng.standard_normal((B, N)) * np.sqrt(k0)[:, None]
    ll, logdet, quad, status = batch.log_likelihood(kb, t, y, solver=solver, return_parts=True)
    assert len(set(kb.J.tolist())) > 4                      # a mixed-width batch
    idx = np.array([0, 9, 18, 27, 36, 45, 54, 63])
    out, o_status = _oracle_batch(kb, idx, t, y, kb.ddiag)
    for k, b in enumerate(idx):
        assert (status[b] == 0) == (o_status[k] == 0)
        if status[b] == 0:
            assert logdet[b] == pytest.approx(out[k, 0], rel=RTOL)
            assert quad[b] == pytest.approx(out[k, 1], rel=RTOL)
    # size-independent property on all 64: two stars with the same kernel and data agree bit for bit
    kb2 = kb.take(np.array([5, 5]))
    ll2 = batch.log_likelihood(kb2, t, np.stack([y[5], y[5]]), solver=solver)
    assert ll2[0] == ll2[1] == ll[5]


def test_cfg4_lattice_slice_shared_light_curve(solver):
    """BASELINE configs[3]: hyper-parameter lattice x one 100 000-point light curve passed once
    (GF_FLAG_SHARED_Y): 32 grid points on the GPU, 4 through the oracle; and the lattice point with
    all factors 1 equals the plain solar kernel."""
    kb, _ = workloads.lattice_batch(32, 3)
    N = 100000
    t = np.arange(N) * 6e-5
    y = np.random.default_rng(3).standard_normal(N) * 285.0
    ll, logdet, quad, status = batch.log_likelihood(kb, t, y, solver=solver, return_parts=True,
                                                    flags=S.FLAG_SHARED_Y)
    assert np.all(status == 0) and np.all(np.isfinite(ll))
    idx = np.array([0, 11, 22, 31])
    out, o_status = _oracle_batch(kb, idx, t, [y] * kb.B, kb.ddiag)
    for k, b in enumerate(idx):
        assert o_status[k] == 0
        assert logdet[b] == pytest.approx(out[k, 0], rel=RTOL)
        assert quad[b] == pytest.approx(out[k, 1], rel=RTOL)
    # replicated y (the ordinary layout) gives the same numbers as the shared one
    ll_rep = batch.log_likelihood(kb.take(idx), t, np.tile(y, (len(idx), 1)), solver=solver)
    np.testing.assert_array_equal(ll_rep, ll[idx])


def test_cfg5_psd_million_bins(solver, solar_kernel, giant_kernel):
    """BASELINE configs[4]: kernel PSD on a 10^6-bin grid up to the Nyquist frequency of the 1-min
    cadence, rtol 1e-12 against the oracle (every bin)."""
    F = 1_000_000
    omega = 2 * np.pi * np.linspace(0.01, 8333.0, F)
    kb = KernelBatch([solar_kernel, giant_kernel])
    got = solver.psd(kb, omega)
    for k, kern in enumerate((solar_kernel, giant_kernel)):
        ref = oracle.psd(kern.base_coefficients(), omega, kern.exposure)
        assert np.max(np.abs(got[k] / ref - 1)) <= 1e-12


def test_cfg3_million_point_sample_round_trip(solver, solar_kernel):
    """BASELINE configs[2] at full length: one 2^20-point solar light curve drawn with the fused
    Philox kernel; its log-likelihood pieces are those of the normals it was drawn from
    (z = L^-1 x = sqrt(d) n  =>  quad = sum n^2), and the log-determinants of the two passes agree."""
    N = 1 << 20
    t = np.arange(N) * 6e-5
    x, st = batch.sample([solar_kernel], t, seed=31, seq0=7, solver=solver, subtract_mean=False)
    assert st[0] == 0
    ll, logdet, quad, status = batch.log_likelihood([solar_kernel], t, x, solver=solver, return_parts=True)
    n = philox.normals(31, 7, N)
    assert status[0] == 0
    assert quad[0] == pytest.approx(float(np.sum(n * n)), rel=1e-8)     # round trip through L and L^-1
    kb = KernelBatch([solar_kernel])
    _, logdet2, _ = solver.sample(kb, Geometry.shared_t(1, N), t, seed=1)
    assert logdet2[0] == logdet[0]


def test_randomised_stress_fixed_seed():
    """tools/stress.py with a fixed seed and a fixed number of batches: random widths (J = 2 ... 172),
    lengths, cadence patterns (uniform, jittered, gaps, cadence changes, absolute time stamps), batch
    sizes, modes and kernel paths against the oracle; identical not-positive-definite reports; worst
    deviation within the north star's 1e-9."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress.py"), "150", "11", "0", "60"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("stress ok")][-1]
    assert "in 60 batches" in line, line
    worst = float(line.rsplit(" ", 1)[1])
    assert worst <= RTOL, line


def test_sharded_legs_two_gpus_nccl_gather():
    """bench.py's sharded legs (cfg4 grid, cfg5 PSD) on two GPUs: every rank scans its block, the
    NCCL all-gather of log L / status / checksums runs inside the timed region, and each rank
    re-computes a slice the OTHER rank owns and finds it bit-identical in the gathered vector
    (asserted inside bench.py).  Skips on a single-GPU box."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--steps", "1", "--warmup", "3", "--n-points", "8192", "--no-cpu",
           "--grid", "300", "--psd-stars", "300"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2
    assert "NCCL" in line["cfg4"]["gather"]["collective"] and "bit-identical" in line["cfg4"]["gather"]["verified"]
    assert line["cfg5"]["max_rel_vs_closed_form"] < 1e-6
