"""B200-native statevector -> reduced qumode rho -> Wigner grid path of C2QA/bosonic-qiskit.

Drop-in for the reference's ``wigner`` / ``util`` analysis entry points and the ``animate`` frame
loops; ``CVCircuit``, ``CVOperators``, ``QumodeRegister`` and the Aer-driven ``simulate()`` stay the
reference's own.  All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/bq_b200.h`` (``libbq_b200.so``); there is no CPU fallback.
"""

from . import animate, util, wigner  # noqa: F401
from ._lib import BQError  # noqa: F401
from .engine import Engine, MultiEngine, device_count, get_engine, get_multi_engine  # noqa: F401

__version__ = "0.2.0"
__all__ = ["wigner", "util", "animate", "Engine", "MultiEngine", "get_engine", "get_multi_engine", "device_count", "BQError",
           "__version__"]
