"""Parity of the CUDA partial-trace kernels (K1 / K2 / K5) with the CPU oracle, through the C ABI."""

import os

import numpy as np
import pytest

import workloads
from oracle import partial_trace_dm, partial_trace_sv

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize(
    "n,traced",
    [
        (1, [0]), (1, []), (3, [2]), (3, [0, 1, 2]), (5, [4]), (5, [0, 2]), (6, [5]), (8, [1, 6, 3]),
        (10, [9, 0]), (12, list(range(6, 12))), (13, list(range(6, 13))), (13, list(range(0, 12))),
        (16, list(range(4, 16))), (16, [0, 1, 2, 3] + list(range(8, 16))),
    ],
)
def test_sv_vs_oracle(engine, n, traced):
    psi = workloads.random_state(1 << n, seed=n)
    rho = engine.partial_trace_sv(psi, traced)
    ref = partial_trace_sv(psi, traced)
    assert rho.shape == ref.shape
    assert rel_err(rho, ref) < TOL


def test_notebook_golden(engine):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_cutoff4.npz"))
    rho = engine.partial_trace_sv(g["psi"], list(g["traced"]))
    assert np.abs(rho - g["rho"]).max() < 2e-8  # the notebook prints 8 digits


def test_batched_and_strided(engine):
    psi = np.stack([workloads.random_state(64, seed=s) for s in range(9)])
    rho = engine.partial_trace_sv(psi, [5])
    for b in range(9):
        assert rel_err(rho[b], partial_trace_sv(psi[b], [5])) < TOL


def test_multi_sets_config4(engine):
    wl = workloads.config4()
    rhos = engine.partial_trace_sv_multi(wl["psi"][0], wl["traced_sets"])
    assert rhos.shape == (4, 16, 16)
    for j, traced in enumerate(wl["traced_sets"]):
        assert rel_err(rhos[j], partial_trace_sv(wl["psi"][0], traced)) < TOL
    assert np.allclose(np.trace(rhos, axis1=1, axis2=2), 1.0, atol=1e-12)


def test_hermitian_psd_trace(engine):
    psi = workloads.random_state(1 << 12, seed=7)
    rho = engine.partial_trace_sv(psi, [0, 3, 4, 7, 8, 11])
    assert np.abs(rho - rho.conj().T).max() < 1e-15
    assert abs(np.trace(rho) - 1) < 1e-12
    assert np.linalg.eigvalsh(rho).min() > -1e-14


@pytest.mark.parametrize("n,traced", [(2, [0]), (4, [1, 3]), (5, [0, 1, 2, 3, 4]), (6, [5]), (7, [0, 6])])
def test_dm_vs_oracle(engine, n, traced):
    rho = workloads.random_rho(1 << n, seed=n)
    out = engine.partial_trace_dm(rho, traced)
    assert rel_err(out, partial_trace_dm(rho, traced)) < TOL


def test_photon_number(engine):
    rhos = np.stack([workloads.random_rho(16, seed=s) for s in range(5)])
    got = engine.photon_number(rhos)
    ref = np.array([np.trace(np.diag(np.arange(16)) @ r) / np.trace(r) for r in rhos])
    assert np.abs(got - ref).max() < 1e-12


def test_bad_arguments(engine):
    from bosonic_qiskit_b200 import BQError

    psi = workloads.random_state(8, seed=0)
    with pytest.raises(BQError):
        engine.partial_trace_sv(psi, [3])
    with pytest.raises(BQError):
        engine.partial_trace_sv(psi, [1, 1])
    with pytest.raises(ValueError):
        engine.partial_trace_sv(np.ones(6, dtype=complex), [0])
