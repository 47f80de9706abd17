"""CPU oracle for the statevector -> reduced rho -> Wigner grid path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bosonic_qiskit_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as
the checker / the reported CPU baseline -- never as the thing shipped.

Parity status: **parity unpinned by the reference's own tests** -- the
arithmetic of this path lives in two third-party libraries that are absent
from ``/root/reference`` and from this image (qutip 5.2.2 ``qutip/wigner.py``
and qiskit 2.1.2 ``quantum_info/states/utils.py::partial_trace``, pinned in
the reference's ``uv.lock:2935-2936`` / ``uv.lock:2841-2842``), and the
reference's tests assert no Wigner value (``tests/test_wigner.py:102,113,150,192``).
The restatement here is therefore pinned against (i) analytic closed forms
(Fock, coherent, cat, single |m><n| elements), (ii) cross-agreement of the
three independent QuTiP algorithms (clenshaw / iterative / laguerre) and
(iii) the one printed golden pair in the reference tree
(``tutorials/wigner-function/wigner-function.ipynb:309-313,353-361``).
"""

from .ptrace_np import partial_trace_sv, partial_trace_dm, partial_trace  # noqa: F401
from .wigner_np import (  # noqa: F401
    wigner,
    wigner_clenshaw,
    wigner_iterative,
    wigner_laguerre,
    wigner_fft,
    osc_eigen,
    flops_per_point,
)
