import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """The in-tree CUDA library and the C oracle, compiled if stale (nvcc cross-compiles)."""
    import __graft_entry__ as entry
    entry.build()
    return entry.LIB


@pytest.fixture(scope="session")
def solar_kernel():
    import gadfly_b200 as g
    return g.SolarOscillatorKernel(texp=1 * g.units.min, bandpass='SOHO VIRGO')


@pytest.fixture(scope="session")
def giant_kernel():
    import gadfly_b200 as g
    hp = g.Hyperparameters.for_star(0.9, 10.0, 4919.0, 52.3, bandpass='SOHO VIRGO', quiet=True)
    return g.StellarOscillatorKernel(hp, texp=1 * g.units.min)


@pytest.fixture(scope="session")
def solver(built):
    from gadfly_b200.solver import Solver
    return Solver(0)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
