"""gadfly_b200: B200-native (sm_100a) GP hot path behind gadfly's Python API."""
__version__ = "0.1.0"

from . import units  # noqa: F401
from .core import *  # noqa: F401,F403
from .terms import *  # noqa: F401,F403
from .gp import *  # noqa: F401,F403
from .psd import *  # noqa: F401,F403
from .solver import Solver, KernelBatch, LinAlgError, SolverUnavailable  # noqa: F401
from . import batch  # noqa: F401
from . import feeder  # noqa: F401,E402
