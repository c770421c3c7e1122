#!/usr/bin/env python
"""Where does the end-to-end step lose time against the device-resident one?  Times, for the bench
workload (148 x 2^20, J = 172): the two kernels alone, each call from / to pinned host memory, and the
overlapped step of bench.py.  usage: python tools/e2e_probe.py [n_points]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gadfly_b200 as g
from gadfly_b200 import batch, solver as S
from gadfly_b200.solver import Geometry, KernelBatch, Solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
solver = Solver(0)
dev = torch.device("cuda", 0)
B = 148
kb = KernelBatch([kernel] * B)
geom = Geometry.shared_t(B, N)
t_dev = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
y_dev = torch.randn(B * N, dtype=torch.float64, device=dev) * 285.0
x_dev = torch.empty(B * N, dtype=torch.float64, device=dev)
t_np = t_dev.cpu().pin_memory().numpy()
y_host = torch.empty(B * N, dtype=torch.float64).pin_memory(); y_host.copy_(y_dev); y_np = y_host.numpy()
x_host = torch.empty(B * N, dtype=torch.float64).pin_memory(); x_np = x_host.numpy()


def timed(fn, reps=3):
    fn(); solver.synchronize(); torch.cuda.synchronize()
    best = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(); solver.synchronize(); torch.cuda.synchronize()
        best.append(time.perf_counter() - t0)
    return min(best) * 1e3


k1 = timed(lambda: solver.loglike(kb, geom, t_dev, y_dev))
k2 = timed(lambda: solver.sample(kb, geom, t_dev, seed=1, out=x_dev))
print(f"device-resident: loglike {k1:.1f} ms, sample {k2:.1f} ms, sum {k1 + k2:.1f} ms")
h1 = timed(lambda: solver.loglike(kb, geom, t_np, y_np))
h2 = timed(lambda: solver.sample(kb, geom, t_np, seed=1, out=x_np))
print(f"from / to pinned host, one call at a time: loglike {h1:.1f} ms (+{h1 - k1:.1f}), sample {h2:.1f} ms (+{h2 - k2:.1f})")


def step():
    solver.sample(kb, geom, t_np, seed=2, out=x_np, flags=S.FLAG_ASYNC)
    return batch.log_likelihood(kb, t_np, y_np, solver=solver)


e = timed(step)
print(f"overlapped step (sample async, then log_likelihood): {e:.1f} ms (+{e - k1 - k2:.1f} over the two kernels)")


def step2():
    solver.loglike(kb, geom, t_np, y_np, flags=S.FLAG_ASYNC)
    solver.sample(kb, geom, t_np, seed=2, out=x_np)


e2 = timed(step2)
print(f"other order (log_likelihood async, then sample): {e2:.1f} ms (+{e2 - k1 - k2:.1f})")
# raw copy speeds
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); y_dev.copy_(y_host, non_blocking=True); ev1.record(); torch.cuda.synchronize()
print(f"H2D 1.24 GB: {ev0.elapsed_time(ev1):.1f} ms")
ev0.record(); x_host.copy_(x_dev, non_blocking=True); ev1.record(); torch.cuda.synchronize()
print(f"D2H 1.24 GB: {ev0.elapsed_time(ev1):.1f} ms")


def step_a():   # only the D2H of x has to hide (behind the log-likelihood kernel)
    solver.sample(kb, geom, t_dev, seed=2, out=x_np, flags=S.FLAG_ASYNC)
    solver.loglike(kb, geom, t_dev, y_dev)


def step_b():   # only the H2D of y has to hide (behind the sample kernel)
    solver.sample(kb, geom, t_dev, seed=2, out=x_dev, flags=S.FLAG_ASYNC)
    solver.loglike(kb, geom, t_dev, y_np)


a, b_ = timed(step_a), timed(step_b)
print(f"only D2H(x) to hide: {a:.1f} ms (+{a - k1 - k2:.1f});  only H2D(y) to hide: {b_:.1f} ms (+{b_ - k1 - k2:.1f})")

side = torch.cuda.Stream(device=dev)


def step_c():   # the H2D of y issued by torch on its own stream right after the sample launch
    solver.sample(kb, geom, t_dev, seed=2, out=x_dev, flags=S.FLAG_ASYNC)
    with torch.cuda.stream(side):
        y_dev.copy_(y_host, non_blocking=True)
    side.synchronize()
    solver.loglike(kb, geom, t_dev, y_dev)


c = timed(step_c)
print(f"H2D(y) by torch on a side stream beside the sample kernel: {c:.1f} ms (+{c - k1 - k2:.1f})")

# finer: time stamps of the pieces of step_c on the device clock
solver.synchronize(); torch.cuda.synchronize()
cstream = torch.cuda.ExternalStream(solver.stream, device=dev)
e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
e[0].record(cstream)
solver.sample(kb, geom, t_dev, seed=2, out=x_dev, flags=S.FLAG_ASYNC)
e[1].record(cstream)                       # end of K2
with torch.cuda.stream(side):
    e[2].record(side)
    y_dev.copy_(y_host, non_blocking=True)
    e[3].record(side)
side.synchronize()
th = time.perf_counter()
solver.loglike(kb, geom, t_dev, y_dev, flags=S.FLAG_ASYNC)
e[4].record(cstream)                       # end of K1
solver.synchronize(); torch.cuda.synchronize()
print("device clock [ms] from the start of the step: K2 ends %.1f | H2D %.1f .. %.1f | K1 ends %.1f" % (
    e[0].elapsed_time(e[1]), e[0].elapsed_time(e[2]), e[0].elapsed_time(e[3]), e[0].elapsed_time(e[4])))

# the copy issued BEFORE the kernel launch (both queued back to back)
solver.synchronize(); torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
e[0].record(cstream)
with torch.cuda.stream(side):
    e[2].record(side)
    y_dev.copy_(y_host, non_blocking=True)
    e[3].record(side)
solver.sample(kb, geom, t_dev, seed=2, out=x_dev, flags=S.FLAG_ASYNC)
e[1].record(cstream)
solver.synchronize(); torch.cuda.synchronize()
print("copy issued first: H2D %.1f .. %.1f | K2 ends %.1f" % (e[0].elapsed_time(e[2]), e[0].elapsed_time(e[3]), e[0].elapsed_time(e[1])))
# a second copy issued 100 ms into the kernel, on a FRESH stream, no events on the compute stream in between
fresh = torch.cuda.Stream(device=dev)
solver.synchronize(); torch.cuda.synchronize()
t0 = time.perf_counter()
solver.sample(kb, geom, t_dev, seed=2, out=x_dev, flags=S.FLAG_ASYNC)
time.sleep(0.1)
with torch.cuda.stream(fresh):
    y_dev.copy_(y_host, non_blocking=True)
fresh.synchronize()
t1 = time.perf_counter()
solver.synchronize()
t2 = time.perf_counter()
print("copy issued 100 ms into the kernel on a fresh stream: copy done at %.1f ms, kernel done at %.1f ms (host clock)" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
