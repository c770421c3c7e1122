#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run in the build container,
where /root/reference exists; the GPU box only reads the .npz files).

1. ``ref_sho_psd.npz`` -- outputs of the REFERENCE'S OWN code: the function ``_sho_psd`` is
   extracted (by AST, unmodified) from /root/reference/gadfly/core.py:33-41 and executed on the
   reference's solar hyperparameters (data/hyperparameters.json + the BiSON table mapped with
   the arithmetic of gadfly/sun.py:36-62).  The module itself cannot be imported here (astropy,
   celerite2 and tynt are absent), the function is pure numpy.  This pins the un-convolved
   kernel PSD (A.5) -- the only hot-path quantity the reference states in closed form.
2. ``dense_*.npz`` -- ground truth from the kernel *definition*: dense covariance matrix of the
   exposure-integrated kernel + ``numpy.linalg.cholesky`` in FP64 (log-likelihood, L n, K^-1 y).
   Inputs and outputs are stored, so the oracle and the CUDA path are both checked against the
   same numbers.  (The reference ships no golden vectors for this path: SURVEY.md section 8c.)
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(HERE, "tests", "golden")


def reference_function(path, name):
    """Compile one top-level function of a reference source file, unmodified."""
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def read_ecsv(path):
    rows, names = [], None
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            if names is None:
                names = line.split()
                continue
            rows.append([float(x) for x in line.split()])
    return names, np.array(rows)


def solar_sho_parameters():
    """(S0, w0, Q) of the reference's solar fit: 5 granulation terms + 81 p-modes."""
    with open(os.path.join(REF, "gadfly", "data", "hyperparameters.json")) as fh:
        hp = json.load(fh)
    gran = np.array([[p["hyperparameters"][k] for k in ("S0", "w0", "Q")]
                     for p in hp if p["metadata"]["source"] == "granulation"])
    osc = sorted((p for p in hp if p["metadata"]["source"] == "oscillation"),
                 key=lambda p: p["metadata"]["degree"])
    _, tab = read_ecsv(os.path.join(REF, "gadfly", "data", "broomhall2009_table2_labeled.ecsv"))
    freq, ell = tab[:, 0], tab[:, 1].astype(int)
    S0 = np.array([osc[l]["hyperparameters"]["S0"] for l in ell])
    Q = np.array([osc[l]["hyperparameters"]["Q"] for l in ell])
    pm = np.stack([S0, 2 * np.pi * freq, Q], axis=1)
    return np.concatenate([gran, pm], axis=0)


def main():
    os.makedirs(OUT, exist_ok=True)
    sho_psd = reference_function(os.path.join(REF, "gadfly", "core.py"), "_sho_psd")
    params = solar_sho_parameters()
    freq = np.sort(np.concatenate([np.logspace(-1, 3.5, 400), np.linspace(2000, 4500, 400),
                                   params[5:45, 1] / (2 * np.pi)]))   # incl. exact resonances
    omega = 2 * np.pi * freq
    per_term = sho_psd(omega[None, :], params[:, 0:1], params[:, 1:2], params[:, 2:3])
    np.savez(os.path.join(OUT, "ref_sho_psd.npz"), params=params, omega=omega,
             psd_sum=per_term.sum(0), psd_terms=per_term[[0, 4, 5, 40, 85]])

    # dense ground truth (uses the oracle's numpy term algebra only for the kernel function)
    from oracle import terms_oracle as T, dense
    import gadfly_b200 as g

    rng = np.random.default_rng(20261018)
    cases = {}
    sun = g.Hyperparameters.for_sun()
    sun_sho = [(p["hyperparameters"]["S0"], p["hyperparameters"]["w0"], p["hyperparameters"]["Q"])
               for p in sun]
    giant = g.Hyperparameters.for_star(0.9, 10.0, 4919.0, 52.3, bandpass="SOHO VIRGO", quiet=True)
    giant_sho = [(p["hyperparameters"]["S0"], p["hyperparameters"]["w0"], p["hyperparameters"]["Q"])
                 for p in giant]
    cases["sun_n384"] = (sun_sho, 6e-5, np.arange(384) * 6e-5, 0.0)   # last entry: diag / k(0)
    tj = np.sort(rng.uniform(0, 600 * 8.64e-5, 300))
    tj = tj[np.concatenate([[True], np.diff(tj) > 6.5e-5])]          # ragged, gaps >= exposure
    cases["sun_ragged"] = (sun_sho, 6e-5, tj, 3e-4)
    cases["giant_n512"] = (giant_sho, 6e-5, np.arange(512) * 3 * 6e-5, 1e-4)
    cases["gran_only"] = (sun_sho[:5], 0.0, np.sort(rng.uniform(0, 5.0, 200)), 1e-6)
    for name, (sho, delta, t, dg) in cases.items():
        coeffs = T.sho_sum(sho)
        scan = T.scan_coefficients(coeffs, delta)
        k0 = np.sum(scan[0]) + np.sum(scan[2]) + scan[6]
        diag = np.full(len(t), dg * k0)
        K = dense.covariance_semiseparable(scan, t, diag)
        y = rng.standard_normal(len(t)) * np.sqrt(K[0, 0])
        n = rng.standard_normal(len(t))
        L = np.linalg.cholesky(K)
        z = np.linalg.solve(L, y)
        np.savez(os.path.join(OUT, f"dense_{name}.npz"), sho=np.array(sho), delta=delta, t=t,
                 diag=diag, y=y, normals=n,
                 loglike=-0.5 * z @ z - np.sum(np.log(np.diag(L))) - 0.5 * len(t) * np.log(2 * np.pi),
                 logdet=2 * np.sum(np.log(np.diag(L))), dot_tril=L @ n,
                 apply_inverse=np.linalg.solve(K, y), d0=K[0, 0])
        print(name, "N", len(t), "J", 2 * len(sho), "cond", np.linalg.cond(K))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
