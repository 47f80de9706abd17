"""The drop-in API cases (tests/dropin_cases.py) against the real CUDA engine."""

import inspect

import pytest

import dropin_cases as cases

pytestmark = pytest.mark.gpu
CASES = [getattr(cases, n) for n in sorted(dir(cases)) if n.startswith("case_")]


@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__)
def test_dropin_on_gpu(case, monkeypatch, engine):
    before = engine.launches
    if "monkeypatch" in inspect.signature(case).parameters:
        case(monkeypatch)
    else:
        case()
    if case.__name__ != "case_get_qubit_indices":
        assert engine.launches > before, "the CUDA library did not run: silent fallback?"


def test_many_frames_host_path_copies_blocks_under_compute(engine):
    """>= 8 frames and >= 8 MiB of output on a mirror grid: the host entry point evaluates the frames in blocks and
    copies each finished block to the host under the next block's kernel.  Same values as frame-by-frame calls."""
    import numpy as np

    from oracle import partial_trace_sv, wigner_clenshaw

    rng = np.random.default_rng(5)
    F, S = 36, 400
    psi = rng.normal(size=(F, 16)) + 1j * rng.normal(size=(F, 16))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    x = np.linspace(-5, 5, S)
    W, rho = engine.state_to_wigner(psi, [3], x, return_rho=True)
    assert W.shape == (F, S, S)
    for k in (0, 8, 9, 17, 18, 26, 27, 35):  # first / last frames of every block
        ref_rho = partial_trace_sv(psi[k], [3])
        assert np.abs(rho[k] - ref_rho).max() < 1e-14
        ref = wigner_clenshaw(ref_rho, x)
        assert np.abs(W[k] - ref).max() <= 1e-10 * np.abs(ref).max()
    one = engine.state_to_wigner(psi[20], [3], x)
    assert np.array_equal(one, W[20])
