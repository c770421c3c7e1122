#!/usr/bin/env python
"""Which ingredient of the Python process stops host<->device copies from running beside a scan kernel?
tools/overlap_probe2.cu (C ABI, no torch) overlaps everything; tools/e2e_probe.py (torch in the process)
does not.  Stages, in ONE process, timing "K2 async (device out) then K1 with y in pinned host memory"
against K1 + K2 after each:
  1. no torch; pinned memory from the bundled shared libcudart (cudaHostAlloc) through ctypes
  2. after `import torch`
  3. after torch.cuda.init() + one device tensor
  4. y in memory pinned by torch (Tensor.pin_memory)
  5. t and the output as torch CUDA tensors (the binding orders the handle's stream against torch's)
usage: python tools/overlap_probe3.py [n_points]"""
import ctypes
import glob
import os
import site
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gadfly_b200 as g
from gadfly_b200 import solver as S
from gadfly_b200.solver import DevicePointer, Geometry, KernelBatch, Solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
B = 148
rt = None
for sp in site.getsitepackages():
    for p in glob.glob(sp + "/nvidia/cuda_runtime/lib/libcudart.so.*"):
        rt = ctypes.CDLL(p)
assert rt is not None, "no shared libcudart"
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]


def pinned(n):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), n * 8, 0) == 0
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(n,))


def device(n):
    p = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(p), n * 8) == 0
    return DevicePointer(p.value, n)


kernel = g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')
solver = Solver(0)
kb = KernelBatch([kernel] * B)
geom = Geometry.shared_t(B, N)
t_np = pinned(N); t_np[:] = np.arange(N) * 6e-5
y_np = pinned(B * N)
x_np = pinned(B * N)
t_d, y_d, x_d = device(N), device(B * N), device(B * N)
assert rt.cudaMemcpy(t_d.address, t_np.ctypes.data, N * 8, 1) == 0
solver.sample(kb, geom, t_d, seed=1, out=y_d)
assert rt.cudaMemcpy(y_np.ctypes.data, y_d.address, B * N * 8, 2) == 0


def timed(fn, reps=3):
    fn(); solver.synchronize()
    best = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(); solver.synchronize()
        best.append(time.perf_counter() - t0)
    return min(best) * 1e3


def report(label, t=t_d, y=y_np, x=x_d):
    k1 = timed(lambda: solver.loglike(kb, geom, t, y_d))
    k2 = timed(lambda: solver.sample(kb, geom, t, seed=2, out=x))

    def step():
        solver.sample(kb, geom, t, seed=2, out=x, flags=S.FLAG_ASYNC)
        solver.loglike(kb, geom, t, y)

    e = timed(step)
    print(f"{label}: K1 {k1:.1f} + K2 {k2:.1f}; K2 async then K1 with host y: {e:.1f} ms (+{e - k1 - k2:.1f})", flush=True)


report("1. no torch, cudaHostAlloc'd y")
import torch  # noqa: E402
report("2. after import torch")
torch.cuda.init()
dev = torch.device("cuda", 0)
z = torch.zeros(16, device=dev); torch.cuda.synchronize()
report("3. after torch.cuda.init + a device tensor")
y_t = torch.empty(B * N, dtype=torch.float64).pin_memory()
y_t.numpy()[:] = y_np
report("4. y pinned by torch", y=y_t.numpy())
t_t = torch.arange(N, dtype=torch.float64, device=dev) * 6e-5
x_t = torch.empty(B * N, dtype=torch.float64, device=dev)
torch.cuda.synchronize()


def report_t(label, y):
    yd = torch.empty(B * N, dtype=torch.float64, device=dev); yd.copy_(torch.from_numpy(y_np)); torch.cuda.synchronize()
    k1 = timed(lambda: solver.loglike(kb, geom, t_t, yd))
    k2 = timed(lambda: solver.sample(kb, geom, t_t, seed=2, out=x_t))

    def step():
        solver.sample(kb, geom, t_t, seed=2, out=x_t, flags=S.FLAG_ASYNC)
        solver.loglike(kb, geom, t_t, y)

    e = timed(step)
    print(f"{label}: K1 {k1:.1f} + K2 {k2:.1f}; K2 async then K1 with host y: {e:.1f} ms (+{e - k1 - k2:.1f})", flush=True)


report_t("5. t / out torch CUDA tensors, y cudaHostAlloc'd", y_np)
report_t("6. t / out torch CUDA tensors, y pinned by torch", y_t.numpy())
report("7. raw device pointers again, cudaHostAlloc'd y")
