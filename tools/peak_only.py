#!/usr/bin/env python
"""Run only the library's DFMA peak microbenchmark (gf_device_info(measure=1)) -- the target of the
ncu capture that shows the FP64 pipe saturated (profiles/r2_dfma_peak_ncu.txt)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadfly_b200.solver import Solver
info = Solver(0).device_info(measure=True)
print(f"sm_count {info['sm_count']}  fp64 peak {info['fp64_flops'] / 1e12:.3f} TFLOP/s")
