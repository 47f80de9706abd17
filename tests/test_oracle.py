"""Pins the CPU oracle (oracle/) before anything is compared against it.

The reference's own tests hold no Wigner value and only basis-state partial traces, and
qutip / qiskit are not installable here, so the oracle is pinned against: closed forms, agreement of
the three independently restated QuTiP algorithms, the values printed in the reference's notebook,
and the reference's own partial-trace test inputs.
"""

import os

import numpy as np
import pytest

from oracle import analytic as an
from oracle import (c_oracle, flops_per_point, partial_trace, partial_trace_dm, partial_trace_sv, wigner,
                    wigner_clenshaw, wigner_iterative, wigner_laguerre)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
X = np.linspace(-6, 6, 61)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 11])
def test_fock_closed_form(n):
    rho = np.zeros((16, 16), dtype=complex)
    rho[n, n] = 1
    assert np.abs(wigner_clenshaw(rho, X) - an.wigner_fock(n, X)).max() < 1e-14


@pytest.mark.parametrize("alpha", [0.0, 1 - 1j, 0.5 + 2j])
def test_coherent_closed_form_and_layout(alpha):
    c = an.coherent_amplitudes(alpha, 80)
    W = wigner_clenshaw(np.outer(c, c.conj()), X)
    assert np.abs(W - an.wigner_coherent(alpha, X)).max() < 1e-13
    iy, ix = np.unravel_index(W.argmax(), W.shape)
    assert abs(X[ix] - np.sqrt(2) * alpha.real) < 0.11 and abs(X[iy] - np.sqrt(2) * alpha.imag) < 0.11


@pytest.mark.parametrize("parity", [+1, -1])
def test_cat_closed_form(parity):
    c = an.cat_amplitudes(2.0, 64, parity)
    assert np.abs(wigner_clenshaw(np.outer(c, c.conj()), X) - an.wigner_cat(2.0, parity, X)).max() < 1e-13


@pytest.mark.parametrize("m,n", [(0, 1), (3, 7), (4, 4), (0, 15)])
def test_single_element(m, n):
    rho = np.zeros((16, 16), dtype=complex)
    rho[m, n] = rho[n, m] = 1
    ref = an.wigner_element(m, n, X)
    assert np.abs(wigner_clenshaw(rho, X) - ref).max() < 1e-12 * max(1, np.abs(ref).max())


@pytest.mark.parametrize("g", [np.sqrt(2), 2.0, 1.0])
def test_methods_agree(g):
    rho = an.random_density_matrix(12, 0)
    a, b, c = wigner_clenshaw(rho, X, g=g), wigner_iterative(rho, X, g=g), wigner_laguerre(rho, X, g=g)
    scale = np.abs(a).max()
    assert np.abs(a - b).max() < 1e-14 * max(1, scale / scale) and np.abs(a - c).max() < 1e-14
    assert np.abs(wigner(rho, X, g=g, method="iterative") - b).max() == 0
    with pytest.raises(TypeError):
        wigner(rho, X, method="nope")


def test_normalisation_marginal_and_realness():
    x = np.linspace(-9, 9, 241)
    rho = an.random_density_matrix(10, 3)
    W = wigner_clenshaw(rho, x)
    dx = x[1] - x[0]
    assert abs(W.sum() * dx * dx - 1) < 1e-9
    # x-marginal of a Fock state |1>: |psi_1(x)|^2 = 2 x^2 exp(-x^2)/sqrt(pi)
    r1 = np.zeros((4, 4), dtype=complex)
    r1[1, 1] = 1
    marg = wigner_clenshaw(r1, x).sum(axis=0) * dx
    assert np.abs(marg - 2 * x**2 * np.exp(-x**2) / np.sqrt(np.pi)).max() < 1e-9


def test_rectangular_grid_and_ket_input():
    rho = an.random_density_matrix(6, 1)
    y = np.linspace(-2, 3, 17)
    W = wigner_clenshaw(rho, X, y)
    assert W.shape == (17, 61)
    sub = wigner_clenshaw(rho, X, y[5:6])
    assert np.abs(W[5:6] - sub).max() == 0
    v = an.random_state(6, 4)
    assert np.abs(wigner_clenshaw(v, X) - wigner_clenshaw(np.outer(v, v.conj()), X)).max() == 0


def test_c_oracle_matches_numpy_oracle():
    rho = an.random_density_matrix(40, 2)
    a = wigner_clenshaw(rho, X)
    assert np.abs(c_oracle.wigner_clenshaw(rho, X) - a).max() < 1e-13 * np.abs(a).max() * 100
    assert np.abs(c_oracle.wigner_clenshaw(rho, X, long_double=True) - a).max() < 1e-13 * np.abs(a).max() * 100
    p = an.random_state(1 << 9, 3)
    assert np.abs(c_oracle.partial_trace_sv(p, [0, 4, 8]) - partial_trace_sv(p, [0, 4, 8])).max() < 1e-15


def test_flops_per_point_table():
    # SURVEY 8d / BASELINE.md
    assert [flops_per_point(n) for n in (16, 32, 64, 256, 1024)] == [1320, 5176, 20568, 327960, 5243928]


# ---- partial trace ------------------------------------------------------------------------------
def test_notebook_golden_pair():
    g = np.load(os.path.join(GOLDEN, "notebook_cutoff4.npz"))
    assert np.abs(partial_trace_sv(g["psi"], list(g["traced"])) - g["rho"]).max() < 2e-8
    # and the printed statevector itself is what the reference's displacement construction gives
    psi4 = an.displacement_matrix(1j * np.pi / 2, 4)[:, 0]
    assert np.abs(psi4 - g["psi"][:4]).max() < 1e-8


def test_config1_golden_is_reproducible():
    g = np.load(os.path.join(GOLDEN, "config1_cat.npz"))
    rho = partial_trace_sv(g["psi"], list(g["traced"]))
    assert np.abs(rho - g["rho"]).max() < 1e-15
    W = wigner_clenshaw(rho, np.linspace(-6, 6, 200))
    assert np.abs(W - g["W"]).max() < 1e-15


def test_reference_basis_state_cases():
    # tests/test_util.py:11-70 (cutoff 4 + 1 qubit): index = q*4 + n
    for idx, key in ((4, 0), (1, 1)):
        psi = np.zeros(8, dtype=complex)
        psi[idx] = 1
        rho = partial_trace_sv(psi, [2])
        assert rho.shape == (4, 4) and abs(rho[key, key] - 1) < 1e-15


def test_partial_trace_matrix_form_and_dm_branch():
    rng = np.random.default_rng(0)
    for n, q in ((5, [4]), (6, [0, 3]), (7, [1, 2, 6]), (4, [0, 1, 2, 3]), (3, [])):
        psi = an.random_state(1 << n, n)
        keep = [k for k in range(n) if k not in q]
        # explicit matrix form: rows = kept bits (original significance order), columns = traced bits
        D, T = 1 << len(keep), 1 << len(q)
        Psi = np.zeros((D, T), dtype=complex)
        for i in range(1 << n):
            r = sum(((i >> b) & 1) << k for k, b in enumerate(keep))
            c = sum(((i >> b) & 1) << k for k, b in enumerate(sorted(q)))
            Psi[r, c] = psi[i]
        rho = partial_trace_sv(psi, q)
        assert np.abs(rho - Psi @ Psi.conj().T).max() < 1e-15
        assert np.abs(partial_trace_dm(np.outer(psi, psi.conj()), q) - rho).max() < 1e-15
        assert np.abs(partial_trace(psi, q) - rho).max() == 0
    with pytest.raises(ValueError):
        partial_trace_sv(np.ones(8), [3])
    with pytest.raises(ValueError):
        partial_trace_sv(np.ones(8), [1, 1])
    with pytest.raises(ValueError):
        partial_trace_sv(np.ones(6), [0])


def test_fft_method_restatement():
    """qutip/wigner.py::_wigner_fourier restated (oracle.wigner_fft): qutip's own test criterion (fft against
    iterative on the fft's p axis, 20 levels, linspace(-10, 10, 128): sum|diff| < 1e-7), the tuple it returns, and
    the closed form the GPU path evaluates (sum over eigenvectors taken before the transform)."""
    import numpy as np

    from oracle import osc_eigen, wigner_fft, wigner_iterative

    rng = np.random.default_rng(0)
    N, x = 20, np.linspace(-10, 10, 128)
    psi = rng.normal(size=N) + 1j * rng.normal(size=N)
    psi /= np.linalg.norm(psi)
    psi2 = rng.normal(size=N) + 1j * rng.normal(size=N)
    psi2 /= np.linalg.norm(psi2)
    rho = 0.6 * np.outer(psi, psi.conj()) + 0.4 * np.outer(psi2, psi2.conj())
    for state in (psi, rho):
        Wf, y = wigner_fft(state, x)
        dm = state if state.ndim == 2 else np.outer(state, state.conj())
        assert Wf.shape == (128, 128) and y.shape == (128,)
        assert np.abs(wigner_iterative(dm, x, y) - Wf).sum() < 1e-7
    for S in (52, 53):
        xs = np.linspace(-7, 7, S)
        Wf, y = wigner_fft(rho, xs)
        A = osc_eigen(N, xs)
        R = A.T @ rho @ A
        n = 2 * S
        ks = np.concatenate((np.arange(3 * n // 4, n), np.arange(0, n // 4)))
        ref = np.zeros((S, S))
        for i in range(S):
            m = min(i, S - 1 - i)
            d = np.arange(-m, m + 1)
            w = R[i + d, i - d]
            th = 2 * np.pi * np.outer(ks, d) / n
            ref[i] = (w.real * np.cos(th) + w.imag * np.sin(th)).sum(axis=1)
        p = np.arange(-n / 4, n / 4) * np.pi / (n * (xs[1] - xs[0]))
        ref = (ref / (p[1] - p[0]) / n).T
        assert np.abs(Wf - ref).max() < 1e-12 * np.abs(Wf).max()
        np.testing.assert_allclose(y, p, rtol=1e-13)  # (g = sqrt(2): the axis scalings cancel to rounding)


def test_rdbu_lookup_table_of_the_library_equals_the_restated_matplotlib_one():
    """Host-only entry point (no device): the 256-entry RdBu table behind bq_frame_rgba_dev, bit for bit."""
    import ctypes

    import numpy as np

    from bosonic_qiskit_b200 import _lib
    from oracle.colormap_np import frame_rgba, rdbu_lut

    buf = (ctypes.c_ubyte * 1024)()
    assert _lib.load().bq_rdbu_lut(buf) == 0
    lut = rdbu_lut()
    assert np.array_equal(np.frombuffer(buf, dtype=np.uint8).reshape(256, 4), lut)
    # ColorBrewer anchors survive at both ends and in the middle (white), red = negative values
    assert tuple(lut[0][:3]) == (103, 0, 31) and tuple(lut[255][:3]) == (5, 48, 97)
    img = frame_rgba(np.array([[-1.0, 0.0], [0.5, 1.0]]))
    # band j takes cmap((j + 0.5) / 99): the outermost bands sit one table entry inside the ends
    assert np.array_equal(img[0, 0], lut[int(0.5 / 99 * 256)]) and np.array_equal(img[1, 1], lut[int(98.5 / 99 * 256)])
    assert img[0, 0][0] > 3 * img[0, 0][2] and img[1, 1][2] > 3 * img[1, 1][0] and img[0, 1][:3].min() > 240
    assert (frame_rgba(np.zeros((3, 3))) == frame_rgba(np.zeros((3, 3)))[0, 0]).all()  # abs_max == 0 -> 5: mid band


def test_thirdparty_golden():
    """Reference-held goldens written by tests/golden/make_golden_thirdparty.py on a machine with qutip + qiskit
    (absent from this image: the test skips until the file is committed)."""
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thirdparty_c1_c4.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/thirdparty_c1_c4.npz not generated yet (needs qutip + qiskit: make_golden_thirdparty.py)")
    from oracle import partial_trace_sv, wigner

    z = np.load(path)
    for name in ("c1", "c2", "c3", "c4"):
        rho = partial_trace_sv(z[f"{name}_psi"], list(z[f"{name}_traced"]))
        assert np.abs(rho - z[f"{name}_rho"]).max() / np.abs(z[f"{name}_rho"]).max() < 1e-10
        for method in ("clenshaw", "iterative"):
            W = wigner(z[f"{name}_rho"], z[f"{name}_xvec"], method=method)
            ref = z[f"{name}_W_{method}"]
            assert np.abs(W - ref).max() / np.abs(ref).max() < 1e-10
