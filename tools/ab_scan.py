#!/usr/bin/env python
"""A/B timing of builds of the scan kernel on one GPU (development tool).

usage: tools/ab_scan.py [--n 65536] [--reps 3] lib_a.so lib_b.so ...

Each library (same C ABI, e.g. built by tools/build_variant.sh with different -D flags) is loaded
in its own process through $GADFLY_B200_LIB; the solar kernel (J = 172) runs 148 light curves of
--n points through the fused log-likelihood and the fused Philox sample kernel.  Prints per library
the kernel times, the fraction of the measured FP64 peak, and the largest relative difference of
logdet / quad / samples against the first library (parity between variants; parity against the
oracle is what tests/ check).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(n, reps, dump):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import gadfly_b200 as g
    from gadfly_b200 import solver as S
    from gadfly_b200.solver import Geometry, KernelBatch, Solver
    dev = torch.device("cuda", 0)
    kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
    if os.environ.get("GADFLY_AB_JC"):     # a narrower kernel: the first Jc terms of the solar one
        kernel = g.StellarOscillatorKernel(terms=list(kernel.term.terms)[:int(os.environ["GADFLY_AB_JC"])],
                                           delta=kernel.delta)
    solver = Solver(0)
    info = solver.device_info(measure=True)
    B = info["sm_count"]
    kb = KernelBatch([kernel] * B)
    geom = Geometry.shared_t(B, n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    k0 = float(np.sum(kb.coef[:len(kernel.term.terms), 0]) + kb.ddiag[0])
    t = torch.arange(n, dtype=torch.float64, device=dev) * 6e-5
    y = torch.randn(B * n, dtype=torch.float64, device=dev, generator=gen) * k0 ** 0.5
    x = torch.empty(B * n, dtype=torch.float64, device=dev)
    logdet = torch.empty(B, dtype=torch.float64, device=dev)
    quad = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    ll, sm = [], []
    for i in range(reps + 1):
        solver.loglike(kb, geom, t, y, logdet=logdet, quad=quad, status=status)
        ll.append(solver.last_kernel_ms)
        solver.sample(kb, geom, t, seed=77, seq0=0, out=x, logdet=logdet, status=status)
        sm.append(solver.last_kernel_ms)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0 or os.environ.get("GADFLY_AB_NOASSERT")
    J = kernel.J
    flops = 4.0 * J * J * B * n
    res = dict(loglike_ms=min(ll[1:]), sample_ms=min(sm[1:]), peak=info["fp64_flops"] / 1e12)
    res["frac_loglike"] = flops / (res["loglike_ms"] * 1e-3) / info["fp64_flops"]
    res["frac_sample"] = flops / (res["sample_ms"] * 1e-3) / info["fp64_flops"]
    res["cycles_per_step"] = res["loglike_ms"] * 1e-3 * 1.965e9 / n
    np.savez(dump, logdet=logdet.cpu().numpy(), quad=quad.cpu().numpy(),
             x=x[:4 * n].cpu().numpy())
    print("RESULT " + json.dumps(res))


def main():
    args = sys.argv[1:]
    n, reps = 65536, 3
    libs = []
    while args:
        a = args.pop(0)
        if a == "--n":
            n = int(args.pop(0))
        elif a == "--reps":
            reps = int(args.pop(0))
        elif a == "--child":
            return child(int(args[0]), int(args[1]), args[2])
        else:
            libs.append(a)
    import numpy as np
    base = None
    for k, lib in enumerate(libs):
        dump = f"/tmp/ab_scan_{k}.npz"
        env = dict(os.environ, GADFLY_B200_LIB=os.path.abspath(lib))
        p = subprocess.run([sys.executable, __file__, "--child", str(n), str(reps), dump],
                           env=env, capture_output=True, text=True, timeout=150)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(f"{lib}: FAILED\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}")
            continue
        r = json.loads(line[0][7:])
        d = np.load(dump)
        if base is None:
            base = d
        rel = {k2: float(np.max(np.abs(d[k2] - base[k2])) / np.max(np.abs(base[k2]))) for k2 in ("logdet", "quad", "x")}
        print(f"{os.path.basename(lib):40s} loglike {r['loglike_ms']:8.3f} ms ({100 * r['frac_loglike']:.1f}%)  "
              f"sample {r['sample_ms']:8.3f} ms ({100 * r['frac_sample']:.1f}%)  "
              f"{r['cycles_per_step']:.0f} cyc/step  peak {r['peak']:.2f}  "
              f"rel vs first: logdet {rel['logdet']:.1e} quad {rel['quad']:.1e} x {rel['x']:.1e}", flush=True)


if __name__ == "__main__":
    main()
