"""Pins the CPU oracle AND the CUDA engine against the real third-party code the reference calls, whenever it imports:

  * ``qutip.wigner(qutip.Qobj(rho), xvec, yvec, g, method)``       -- src/bosonic_qiskit/wigner.py:190 (qutip 5.2.2, uv.lock:2935-2936)
  * ``qiskit.quantum_info.partial_trace(Statevector, qargs)``      -- src/bosonic_qiskit/util.py:580,596,619,693 (qiskit 2.1.2,
                                                                       uv.lock:2841-2842)

Neither library is in this image (no network, no wheel), so here every test of this file SKIPS and the oracle stays pinned by
closed forms, QuTiP's own cross-method criterion and the reference's notebook values (tests/test_oracle.py) -- "parity
unpinned" in the sense of the oracle's header.  On a box that has them:

    pip install "qutip>=5.0.4" "qiskit>=2.1,<2.2"
    python -m pytest tests/test_thirdparty_pin.py -q                # oracle vs the libraries (CPU)
    python -m pytest tests/test_thirdparty_pin.py -q -m gpu         # CUDA engine vs the libraries (B200)
    python tests/golden/make_golden_thirdparty.py                   # writes reference-held goldens next to the others

Tolerance (north_star): max|dW| / max|W| <= 1e-10, max|drho| / max|rho| <= 1e-10.
"""

import numpy as np
import pytest

import workloads
from oracle import partial_trace_dm, partial_trace_sv, wigner, wigner_fft

TOL = 1e-10
G = float(np.sqrt(2))


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def small_cases():
    """(name, psi, traced, xvec): BASELINE configs c1-c4 at grid sizes the CPU libraries finish in seconds."""
    c1 = workloads.config1(steps=60)
    c2 = workloads.config2(steps=24)
    c3 = workloads.config3(frames=3, steps=40)
    c4 = workloads.config4(steps=40)
    return [("c1", c1["psi"][0], c1["traced"], c1["xvec"]), ("c2", c2["psi"][0], c2["traced"], c2["xvec"]),
            ("c3-frame2", c3["psi"][2], c3["traced"], c3["xvec"]), ("c4-site1", c4["psi"][0], c4["traced_sets"][1], c4["xvec"])]


@pytest.fixture(scope="module")
def qutip():
    return pytest.importorskip("qutip", reason="qutip is not installed in this image (see this file's docstring)")


@pytest.fixture(scope="module")
def qi():
    pytest.importorskip("qiskit", reason="qiskit is not installed in this image (see this file's docstring)")
    import qiskit.quantum_info as qi

    return qi


# ---------------------------------------------------------------------------------------------- oracle vs the libraries
@pytest.mark.parametrize("case", range(4))
def test_oracle_partial_trace_equals_qiskit(qi, case):
    name, psi, traced, _ = small_cases()[case]
    ref = np.asarray(qi.partial_trace(qi.Statevector(psi), list(traced)).data)
    assert rel(partial_trace_sv(psi, traced), ref) < TOL, name
    dm = qi.DensityMatrix(qi.Statevector(psi[:64] / np.linalg.norm(psi[:64])))
    assert rel(partial_trace_dm(np.asarray(dm.data), [0, 3]), np.asarray(qi.partial_trace(dm, [0, 3]).data)) < TOL


@pytest.mark.parametrize("method", ["clenshaw", "iterative", "laguerre"])
@pytest.mark.parametrize("case", range(4))
def test_oracle_wigner_equals_qutip(qutip, case, method):
    name, psi, traced, x = small_cases()[case]
    rho = partial_trace_sv(psi, traced)
    if method == "laguerre" and rho.shape[0] > 32:
        pytest.skip("qutip's laguerre method builds scipy polynomial objects per element: too slow beyond cutoff 32")
    ref = qutip.wigner(qutip.Qobj(rho), x, x, g=G, method=method)
    assert rel(wigner(rho, x, method=method), ref) < TOL, (name, method)


def test_oracle_wigner_fft_equals_qutip(qutip):
    _, psi, traced, x = small_cases()[0]
    rho = partial_trace_sv(psi, traced)
    W_ref, y_ref = qutip.wigner(qutip.Qobj(rho), x, x, g=G, method="fft")
    W, y = wigner_fft(rho, x)
    assert rel(y, y_ref) < 1e-12 and rel(W, W_ref) < TOL


# ---------------------------------------------------------------------------------------------- CUDA engine vs the libraries
@pytest.mark.gpu
@pytest.mark.parametrize("case", range(4))
def test_engine_equals_qiskit_and_qutip(engine, qi, qutip, case):
    name, psi, traced, x = small_cases()[case]
    W, rho = engine.state_to_wigner(psi, traced, x, return_rho=True)
    rho_ref = np.asarray(qi.partial_trace(qi.Statevector(psi), list(traced)).data)
    assert rel(rho, rho_ref) < TOL, name
    for method in ("clenshaw", "iterative"):
        W_ref = qutip.wigner(qutip.Qobj(rho_ref), x, x, g=G, method=method)
        assert rel(engine.wigner(rho_ref, x, method=method), W_ref) < TOL, (name, method)
    assert rel(W, qutip.wigner(qutip.Qobj(rho_ref), x, x, g=G)) < TOL, name
    engine.set_grid_sharing(False)  # the point-by-point kernels as well
    try:
        assert rel(engine.wigner(rho_ref, x), qutip.wigner(qutip.Qobj(rho_ref), x, x, g=G)) < TOL, name
    finally:
        engine.set_grid_sharing(True)


@pytest.mark.gpu
def test_engine_fft_equals_qutip(engine, qutip):
    _, psi, traced, x = small_cases()[0]
    rho = partial_trace_sv(psi, traced)
    W_ref, y_ref = qutip.wigner(qutip.Qobj(rho), x, x, g=G, method="fft")
    W, y = engine.wigner(rho, x, method="fft")
    assert rel(y, y_ref) < 1e-12 and rel(W, W_ref) < TOL


@pytest.mark.gpu
def test_engine_cutoff_1024_rows_equal_qutip(engine, qutip):
    """BASELINE config 5 at full cutoff on a few grid rows (qutip needs ~20 s per 4096 points at cutoff 1024)."""
    wl = workloads.config5(steps=4096)
    rho = np.outer(wl["psi"][0], wl["psi"][0].conj())
    x = wl["xvec"]
    rows = np.array([300, 2047, 2048, 3900])
    cols = np.arange(0, 4096, 64)
    W = np.asarray(engine.wigner(rho, x))
    ref = qutip.wigner(qutip.Qobj(rho), x[cols], x[rows], g=G)
    assert np.abs(W[np.ix_(rows, cols)] - ref).max() / np.abs(W).max() < TOL
