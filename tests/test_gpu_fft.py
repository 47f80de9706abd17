"""method="fft" of qutip.wigner on the GPU (bq_wigner_fft_host / _dev) against the CPU oracle's restatement of
qutip/wigner.py::_wigner_fourier (eigendecomposition + Toeplitz product + FFT).  The GPU path sums over the
eigenvectors in closed form and evaluates a banded DFT; both compute the same discrete transform, so the tolerance
is the fp64 one of north_star: max|dW| / max|W| <= 1e-10.  The p axis must match to rounding."""

import numpy as np
import pytest

import workloads
from oracle import wigner_clenshaw, wigner_fft

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("N", [1, 2, 5, 16, 33, 64])
@pytest.mark.parametrize("S", [2, 3, 40, 41, 128, 200])
def test_fft_vs_oracle(engine, N, S):
    rho = workloads.random_rho(N, seed=N + S, rank=min(N, 3))
    x = np.linspace(-6, 6, S)
    W, y = engine.wigner_fft(rho, x)
    Wo, yo = wigner_fft(rho, x)
    assert W.shape == (S, S) and y.shape == (S,)
    np.testing.assert_allclose(y, yo, rtol=1e-14, atol=1e-300)
    assert rel_err(W, Wo) < TOL


def test_fft_ring_sized_cutoff_and_other_g(engine):
    rho = workloads.random_rho(150, seed=9, rank=4)
    x = np.linspace(-9, 9, 257)
    for g in (np.sqrt(2), 2.0, 0.7):
        W, y = engine.wigner_fft(rho, x, g=g)
        Wo, yo = wigner_fft(rho, x, g=g)
        np.testing.assert_allclose(y, yo, rtol=1e-14)
        assert rel_err(W, Wo) < TOL


def test_fft_is_the_wigner_function_where_the_method_is_accurate(engine):
    """qutip's own test (qutip/tests/test_wigner.py: fft against iterative on the fft's p axis, random state of 20
    levels on linspace(-10, 10, 128), sum|diff| < 1e-7) with the Clenshaw kernel as the other side."""
    rho = workloads.random_rho(20, seed=1, rank=2)
    x = np.linspace(-10, 10, 128)
    Wf, y = engine.wigner_fft(rho, x)
    Wc = engine.wigner(rho, x, y)
    assert np.abs(Wf - Wc).sum() < 1e-7
    assert np.abs(Wf - wigner_clenshaw(rho, x, y)).sum() < 1e-7


def test_fft_batch_method_routing_and_non_hermitian_input(engine):
    rhos = np.stack([workloads.random_rho(12, seed=s) for s in range(3)])
    psi = workloads.random_rho(12, seed=7, rank=1)[:, 0].copy()
    rhos[2] = np.outer(psi, psi)  # the reference's no-conjugate quirk (wigner.py:130-131): not Hermitian
    x = np.linspace(-5, 5, 60)
    W, y = engine.wigner(rhos, x, method="fft")  # same routing as qutip.wigner(method="fft"): a tuple
    assert W.shape == (3, 60, 60)
    for k in range(3):
        # the oracle follows qutip (eigh of the matrix as given); for the non-Hermitian quirk input eigh reads one
        # triangle only, so compare that frame with the closed form the GPU evaluates instead
        if k < 2:
            assert rel_err(W[k], wigner_fft(rhos[k], x)[0]) < TOL
    from oracle import osc_eigen

    A = osc_eigen(12, x)
    R = A.T @ rhos[2] @ A
    S, n = 60, 120
    ks = np.concatenate((np.arange(3 * n // 4, n), np.arange(0, n // 4)))
    ref = np.zeros((S, S))
    for i in range(S):
        m = min(i, S - 1 - i)
        d = np.arange(-m, m + 1)
        w = R[i + d, i - d]
        th = 2 * np.pi * np.outer(ks, d) / n
        ref[i] = (w.real * np.cos(th) + w.imag * np.sin(th)).sum(axis=1)
    p = np.arange(-n / 4, n / 4) * np.pi / (n * (x[1] - x[0]))
    ref = (ref / (p[1] - p[0]) / n).T
    assert rel_err(W[2], ref) < TOL


def test_fft_device_entry_point_and_fused_frames(engine):
    import torch

    dev = torch.device("cuda", engine.device)
    rho = workloads.random_rho(24, seed=3, rank=2)
    x = np.linspace(-7, 7, 96)
    W_t, y = engine.wigner_fft_t(torch.from_numpy(rho).to(dev), x)
    Wo, yo = wigner_fft(rho, x)
    assert rel_err(W_t.cpu().numpy(), Wo) < TOL
    np.testing.assert_allclose(y, yo, rtol=1e-14)
    wl = workloads.config3(frames=3, steps=48)
    (W, yv), rho_f = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], method="fft", return_rho=True)
    for k in range(3):
        assert rel_err(W[k], wigner_fft(rho_f[k], wl["xvec"])[0]) < TOL


def test_fft_argument_errors(engine):
    from bosonic_qiskit_b200 import BQError

    rho = workloads.random_rho(4, seed=0)
    with pytest.raises(BQError):
        engine.wigner_fft(rho, np.array([0.5]))  # the p axis needs xvec[1] - xvec[0]
    with pytest.raises(BQError):
        engine.wigner_fft(rho, np.array([1.0, 1.0, 1.0]))
    with pytest.raises(TypeError):
        engine.wigner_fft(np.zeros((3, 4), dtype=complex), np.linspace(-1, 1, 8))
