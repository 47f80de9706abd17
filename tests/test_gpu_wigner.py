"""Parity of the CUDA Wigner kernel (K3) with the CPU oracle, through the C ABI.

Tolerance (north_star): max|dW| / max|W| <= 1e-10 in fp64.
"""

import os

import numpy as np
import pytest

import workloads
from oracle import analytic as an
from oracle import c_oracle, wigner_clenshaw, wigner_iterative

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 8, 16, 31, 32, 64])
def test_resident_random_rho(engine, N):
    rho = workloads.random_rho(N, seed=N)
    x = np.linspace(-6, 6, 52)
    W = engine.wigner(rho, x)
    assert W.shape == (52, 52) and W.dtype == np.float64
    assert rel_err(W, wigner_clenshaw(rho, x)) < TOL


@pytest.mark.parametrize("N", [65, 66, 96, 128, 255, 256])
def test_stream_random_rho(engine, N):
    rho = workloads.random_rho(N, seed=N, rank=5)
    x = np.linspace(-6, 6, 40)
    W = engine.wigner(rho, x)
    assert rel_err(W, c_oracle.wigner_clenshaw(rho, x)) < TOL


def test_stream_cutoff_1024(engine):
    rho = workloads.random_rho(1024, seed=1, rank=2)
    x = np.linspace(-6, 6, 24)
    W = engine.wigner(rho, x)
    ref = c_oracle.wigner_clenshaw(rho, x, long_double=True)
    assert rel_err(W, ref) < TOL


@pytest.mark.parametrize("nx,ny", [(1, 1), (7, 3), (33, 65), (130, 4), (201, 199), (256, 128)])
def test_ragged_grids(engine, nx, ny):
    rho = workloads.random_rho(12, seed=3)
    x = np.linspace(-4, 5, nx)
    y = np.linspace(-3, 6, ny)
    W = engine.wigner(rho, x, y)
    assert W.shape == (ny, nx)
    assert rel_err(W, wigner_clenshaw(rho, x, y)) < TOL


def test_large_grid_full_size_shapes(engine):
    # big-tile (256 threads x 4 points) variants of both kernels
    x = np.linspace(-6, 6, 640)
    rho = workloads.random_rho(8, seed=9)
    W = engine.wigner(rho, x)
    assert rel_err(W, c_oracle.wigner_clenshaw(rho, x)) < TOL
    rho = workloads.random_rho(70, seed=9, rank=3)
    W = engine.wigner(rho, x)
    assert rel_err(W, c_oracle.wigner_clenshaw(rho, x)) < TOL


def test_empty(engine):
    rho = workloads.random_rho(4, seed=0)
    assert engine.wigner(rho, np.zeros(0)).shape == (0, 0)
    assert engine.wigner(np.zeros((0, 4, 4), dtype=complex), np.linspace(-1, 1, 5)).shape == (0, 5, 5)


@pytest.mark.parametrize("g", [np.sqrt(2), 2.0, 1.0, 0.5])
def test_g_scaling(engine, g):
    rho = workloads.random_rho(10, seed=2)
    x = np.linspace(-5, 5, 41)
    assert rel_err(engine.wigner(rho, x, g=g), wigner_clenshaw(rho, x, g=g)) < TOL


def test_batched_frames_match_single(engine):
    rhos = np.stack([workloads.random_rho(20, seed=s) for s in range(7)])
    x = np.linspace(-6, 6, 64)
    W = engine.wigner(rhos, x)
    assert W.shape == (7, 64, 64)
    for b in range(7):
        assert rel_err(W[b], wigner_clenshaw(rhos[b], x)) < TOL
    rhos = np.stack([workloads.random_rho(80, seed=s, rank=2) for s in range(3)])
    W = engine.wigner(rhos, x)
    for b in range(3):
        assert rel_err(W[b], c_oracle.wigner_clenshaw(rhos[b], x)) < TOL


@pytest.mark.parametrize("method", ["clenshaw", "iterative", "laguerre"])
def test_method_names(engine, method):
    rho = workloads.random_rho(9, seed=4)
    x = np.linspace(-4, 4, 33)
    assert rel_err(engine.wigner(rho, x, method=method), wigner_iterative(rho, x)) < TOL


def test_bad_method(engine):
    rho = workloads.random_rho(4, seed=0)
    with pytest.raises(TypeError):
        engine.wigner(rho, np.linspace(-1, 1, 4), method="nope")
    import ctypes

    from bosonic_qiskit_b200._lib import BQ_ERR_UNSUPPORTED

    # the grid entry point itself refuses method code 3 ("fft" has its own entry point, tests/test_gpu_fft.py)
    x = np.linspace(-1, 1, 4)
    out = np.empty((1, 4, 4))
    rc = engine._lib.bq_wigner_host(engine._h, ctypes.c_void_p(rho.ctypes.data), 4, 4, 16, 1, ctypes.c_void_p(x.ctypes.data), 4,
                                    ctypes.c_void_p(x.ctypes.data), 4, 2 ** 0.5, 3, ctypes.c_void_p(out.ctypes.data))
    assert rc == BQ_ERR_UNSUPPORTED


# ---- analytic known answers (the reference pins no Wigner value) --------------------
@pytest.mark.parametrize("n", [0, 1, 5, 15])
def test_fock(engine, n):
    rho = np.zeros((16, 16), dtype=complex)
    rho[n, n] = 1
    x = np.linspace(-6, 6, 101)
    assert np.abs(engine.wigner(rho, x) - an.wigner_fock(n, x)).max() < 1e-13


def test_coherent_and_layout(engine):
    c = an.coherent_amplitudes(1 - 1j, 64)
    x = np.linspace(-6, 6, 121)
    W = engine.wigner(np.outer(c, c.conj()), x)
    assert np.abs(W - an.wigner_coherent(1 - 1j, x)).max() < 1e-13
    iy, ix = np.unravel_index(W.argmax(), W.shape)
    assert x[ix] > 1.3 and x[iy] < -1.3  # W[iy, ix]: centre at (x, p) = (+sqrt2, -sqrt2)


def test_reference_tutorial_inputs(engine):
    """Inputs of the reference's own notebooks / tests, which assert no Wigner value themselves:
    coherent alpha = 1 - 1j at cutoff 128 (tutorials/bosonic-qiskit-textbook/4. wigner_distribution.ipynb, cells
    9-18: ring kernels), the binomial code words (|0> + |4>)/sqrt2 and |2> (tests/test_wigner.py cat / code-word
    plots), cat alpha = 1, each on the default 200 x 200 grid of wigner.wigner -- against closed forms."""
    x = np.linspace(-6, 6, 200)
    c = an.coherent_amplitudes(1 - 1j, 128)
    assert np.abs(engine.wigner(np.outer(c, c.conj()), x) - an.wigner_coherent(1 - 1j, x)).max() < 1e-13
    psi = np.zeros(16, dtype=complex)
    psi[0] = psi[4] = 1 / np.sqrt(2)
    ref = 0.5 * (an.wigner_element(0, 0, x) + an.wigner_element(4, 4, x) + an.wigner_element(0, 4, x))  # (0, 4) = the Hermitian pair
    assert np.abs(engine.wigner(np.outer(psi, psi.conj()), x) - ref).max() < 1e-13
    rho = np.zeros((16, 16), dtype=complex)
    rho[2, 2] = 1
    assert np.abs(engine.wigner(rho, x) - an.wigner_fock(2, x)).max() < 1e-13
    c = an.cat_amplitudes(1.0, 16, +1)
    assert np.abs(engine.wigner(np.outer(c, c.conj()), x) - an.wigner_cat(1.0, +1, x)).max() < 1e-6  # cutoff 16 truncates the cat (4e-8)


def test_cat(engine):
    c = an.cat_amplitudes(2.0, 64, +1)
    x = np.linspace(-6, 6, 99)
    assert np.abs(engine.wigner(np.outer(c, c.conj()), x) - an.wigner_cat(2.0, +1, x)).max() < 1e-13


@pytest.mark.parametrize("m,n", [(0, 3), (3, 7), (2, 2), (10, 15)])
def test_single_element(engine, m, n):
    rho = np.zeros((16, 16), dtype=complex)
    rho[m, n] = 1
    rho[n, m] = 1
    x = np.linspace(-5, 5, 77)
    ref = an.wigner_element(m, n, x)
    assert np.abs(engine.wigner(rho, x) - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_linearity_and_normalisation(engine):
    a, b = workloads.random_rho(24, seed=1), workloads.random_rho(24, seed=2)
    x = np.linspace(-9, 9, 301)
    Wa, Wb, Wab = engine.wigner(a, x), engine.wigner(b, x), engine.wigner(0.3 * a + 0.7 * b, x)
    assert rel_err(Wab, 0.3 * Wa + 0.7 * Wb) < 1e-12
    dx = x[1] - x[0]
    c = an.coherent_amplitudes(0.7 + 0.2j, 24)
    Wc = engine.wigner(np.outer(c, c.conj()), x)
    assert abs(Wc.sum() * dx * dx - 1.0) < 1e-9


def test_only_upper_triangle_is_read(engine):
    rho = workloads.random_rho(16, seed=6)
    junk = rho.copy()
    junk[np.tril_indices(16, -1)] = 123.0
    x = np.linspace(-5, 5, 50)
    assert np.array_equal(engine.wigner(rho, x), engine.wigner(junk, x))


# ---- BASELINE.json configs ------------------------------------------------------------------
def test_config1_golden(engine):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_cat.npz"))
    x = np.linspace(-6, 6, 200)
    W, rho = engine.state_to_wigner(g["psi"], list(g["traced"]), x, return_rho=True)
    assert rel_err(rho, g["rho"]) < TOL
    assert rel_err(W, g["W"]) < TOL


def test_config2_full_size(engine):
    wl = workloads.config2()
    W, rho = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], return_rho=True)
    assert W.shape == (1, 1024, 1024)
    # oracle on a row subsample (the C oracle; the NumPy one needs ~30 s for the full grid)
    rows = np.arange(0, 1024, 37)
    ref = c_oracle.wigner_clenshaw(rho[0], wl["xvec"], wl["xvec"][rows])
    assert rel_err(W[0][rows], ref) < TOL
    # size-independent properties at full size: normalisation and symmetry under rho -> rho^T (p -> -p)
    dx = wl["xvec"][1] - wl["xvec"][0]
    assert abs(W.sum() * dx * dx - 1.0) < 1e-6
    Wt = engine.wigner(np.ascontiguousarray(rho[0].T), wl["xvec"])
    assert rel_err(Wt, W[0][::-1, :]) < 1e-12


def test_config3_frames(engine):
    wl = workloads.config3(frames=24, steps=400)
    W = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"])
    assert W.shape == (24, 400, 400)
    from oracle import partial_trace_sv

    for k in (0, 11, 23):
        ref = c_oracle.wigner_clenshaw(partial_trace_sv(wl["psi"][k], wl["traced"]), wl["xvec"])
        assert rel_err(W[k], ref) < TOL
    # a coherent state only rotates: the peak height is frame independent
    assert np.ptp(W.max(axis=(1, 2))) < 1e-3


def test_config5_slab(engine):
    wl = workloads.config5(steps=4096)
    rho = np.outer(wl["psi"][0], wl["psi"][0].conj())
    rows = np.array([0, 1000, 2047, 2048, 4095])
    W = engine.wigner(rho, wl["xvec"], wl["xvec"][rows])
    cols = np.arange(0, 4096, 64)
    ref = c_oracle.wigner_clenshaw(rho, wl["xvec"][cols], wl["xvec"][rows], long_double=True)
    assert rel_err(W[:, cols], ref) < TOL


def test_engine_against_thirdparty_golden(engine):
    """CUDA path against goldens produced by qutip + qiskit themselves (tests/golden/make_golden_thirdparty.py);
    skips until that file has been generated on a machine that has the two libraries."""
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thirdparty_c1_c4.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/thirdparty_c1_c4.npz not generated yet (needs qutip + qiskit)")
    z = np.load(path)
    for name in ("c1", "c2", "c3", "c4"):
        W, rho = engine.state_to_wigner(z[f"{name}_psi"], list(z[f"{name}_traced"]), z[f"{name}_xvec"], return_rho=True)
        assert np.abs(rho - z[f"{name}_rho"]).max() / np.abs(z[f"{name}_rho"]).max() < 1e-10
        ref = z[f"{name}_W_clenshaw"]
        assert np.abs(np.asarray(W) - ref).max() / np.abs(ref).max() < 1e-10
