"""The re-scheduled SASS of the Wigner kernels must be bit-identical to nvcc's own schedule.

The patched image runs the same DFMAs on the same values in a different order with other registers
(bosonic_qiskit_b200/_sass_resched.py), so equality is exact, not a tolerance."""

import numpy as np
import pytest

import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture()
def both_modes(engine):
    have = engine.set_sass_mode(True)
    yield have
    engine.set_sass_mode(True)


def run_both(engine, fn):
    engine.set_sass_mode(False)
    plain = fn()
    assert engine.set_sass_mode(True), "patched image missing from this build"
    patched = fn()
    return plain, patched


def test_patched_image_is_present(engine, both_modes):
    assert both_modes, "libbq_b200.so was built without the re-scheduled cubin"


@pytest.mark.parametrize("N", list(range(1, 20)) + [23, 24, 31, 32, 33, 47, 48, 63, 64])
def test_resident_bit_identical(engine, both_modes, N):
    rho = workloads.random_rho(N, seed=100 + N)
    # three grid sizes -> the three tile shapes (1, 2 and 4 points per thread)
    for S in (40, 300, 1100):
        x = np.linspace(-6, 6, S)
        a, b = run_both(engine, lambda: engine.wigner(rho, x).copy())
        assert np.array_equal(a, b), (N, S, np.abs(a - b).max())


@pytest.mark.parametrize("N", [65, 100, 256])
def test_stream_bit_identical(engine, both_modes, N):
    rho = workloads.random_rho(N, seed=N, rank=4)
    for S in (64, 700):
        x = np.linspace(-6, 6, S)
        a, b = run_both(engine, lambda: engine.wigner(rho, x).copy())
        assert np.array_equal(a, b)


def test_batched_frames_bit_identical(engine, both_modes):
    wl = workloads.config3(frames=9, steps=400)
    a, b = run_both(engine, lambda: engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"]).copy())
    assert np.array_equal(a, b)
