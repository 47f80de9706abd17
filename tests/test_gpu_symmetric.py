"""Mirror-symmetric grids: the recurrence-sharing kernels (wigner_sym_*) against the CPU oracle and against the
point-by-point kernels, through the C ABI (host and device entry points).

The reference only ever builds `xvec = linspace(-a, a, S)`, `yvec = xvec` (src/bosonic_qiskit/wigner.py:135,187-188),
so this is the path every reference call takes.  Tolerance (north_star): max|dW| / max|W| <= 1e-10 in fp64.
"""

import numpy as np
import pytest

import workloads
from oracle import c_oracle, partial_trace_sv, wigner_clenshaw

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture()
def general(engine):
    """The same engine with recurrence sharing switched off (always the point-by-point kernels)."""
    engine.set_grid_sharing(False)
    yield engine
    engine.set_grid_sharing(True)


def test_library_grid_test(engine):
    assert engine.grid_is_mirror(np.linspace(-6, 6, 200))
    assert engine.grid_is_mirror(np.linspace(-6, 6, 201))
    assert engine.grid_is_mirror(np.linspace(-0.37, 0.37, 17))
    assert not engine.grid_is_mirror(np.linspace(-6, 5, 200))
    assert not engine.grid_is_mirror(np.linspace(-6, 6, 200), np.linspace(-6, 6, 199))
    x = np.linspace(-6, 6, 64)
    y = x.copy()
    y[5] += 1e-9
    assert not engine.grid_is_mirror(x, y)
    x2 = x.copy()
    x2[3] *= 1 + 1e-12  # not a mirror image any more
    assert not engine.grid_is_mirror(x2)
    assert not engine.grid_is_mirror(np.array([1.0]))


@pytest.mark.parametrize("N", [2, 3, 4, 5, 8, 16, 31, 32, 33, 64])
@pytest.mark.parametrize("S", [52, 53])
def test_resident_symmetric_vs_oracle(engine, N, S):
    rho = workloads.random_rho(N, seed=N)
    x = np.linspace(-6, 6, S)
    W = engine.wigner(rho, x)
    assert W.shape == (S, S)
    assert rel_err(W, wigner_clenshaw(rho, x)) < TOL


@pytest.mark.parametrize("S", [2, 3, 4, 7, 31, 32, 96, 97, 200, 401, 640])
def test_grid_sizes_vs_general_kernel(engine, S):
    """every grid size class: tiny, odd (a middle line), the small-tile and the big-tile variants"""
    rho = workloads.random_rho(12, seed=S)
    x = np.linspace(-5, 5, S)
    W = np.array(engine.wigner(rho, x))
    engine.set_grid_sharing(False)
    try:
        Wg = np.array(engine.wigner(rho, x))
    finally:
        engine.set_grid_sharing(True)
    assert rel_err(W, Wg) < 1e-13
    assert rel_err(W, wigner_clenshaw(rho, x)) < TOL


@pytest.mark.parametrize("N,S", [(65, 416), (96, 420), (128, 417), (256, 416)])
def test_ring_symmetric_vs_oracle(engine, N, S):
    """N > 64 and enough units for the ring kernel (wigner_sym_diag_kernel)"""
    rho = workloads.random_rho(N, seed=N, rank=4)
    x = np.linspace(-6, 6, S)
    W = engine.wigner(rho, x)
    ref = c_oracle.wigner_clenshaw(rho, x[: S // 2 + 1], x)  # left half (+ middle) of the grid: enough to pin it
    assert rel_err(W[:, : S // 2 + 1], ref) < TOL
    # the other half by symmetry against the point-by-point kernels
    engine.set_grid_sharing(False)
    try:
        Wg = np.array(engine.wigner(rho, x))
    finally:
        engine.set_grid_sharing(True)
    assert rel_err(W, Wg) < 1e-12


def test_ring_small_grid_falls_back(engine):
    """N > 64 with too few units for the ring kernel: point-by-point kernels, same answer"""
    rho = workloads.random_rho(96, seed=5, rank=3)
    x = np.linspace(-5, 5, 48)
    W = engine.wigner(rho, x)
    assert rel_err(W, wigner_clenshaw(rho, x)) < TOL


def test_cutoff_1024_symmetric(engine):
    rho = workloads.random_rho(1024, seed=1, rank=2)
    x = np.linspace(-6, 6, 416)
    W = engine.wigner(rho, x)
    cols = np.array([0, 1, 57, 207, 208, 300, 415])
    ref = c_oracle.wigner_clenshaw(rho, x[cols], x, long_double=True)
    assert rel_err(W[:, cols], ref) < TOL


def test_batched_frames_symmetric(engine):
    wl = workloads.config3(frames=5, steps=120)
    W, rho = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], return_rho=True)
    assert W.shape == (5, 120, 120)
    for f in range(5):
        r = partial_trace_sv(wl["psi"][f], wl["traced"])
        assert rel_err(rho[f], r) < TOL
        assert rel_err(W[f], wigner_clenshaw(r, wl["xvec"])) < TOL


def test_non_mirror_grids_take_the_general_kernels(engine, general):
    rho = workloads.random_rho(20, seed=2)
    x = np.linspace(-4, 5, 130)  # not symmetric
    Wg = np.array(general.wigner(rho, x))
    general.set_grid_sharing(True)
    W = np.array(engine.wigner(rho, x))
    assert np.array_equal(W, Wg)  # same kernel, bit-identical
    assert rel_err(W, wigner_clenshaw(rho, x)) < TOL


def test_device_entry_points_symmetric(engine):
    import torch

    dev = torch.device("cuda", 0)
    rho = workloads.random_rho(24, seed=4)
    x = np.linspace(-6, 6, 300)
    rho_t = torch.from_numpy(rho).to(dev)
    x_t = torch.from_numpy(x).to(dev)
    before = engine.launches
    W = engine.wigner_t(rho_t, x_t).cpu().numpy()
    assert engine.launches > before
    assert rel_err(W, wigner_clenshaw(rho, x)) < TOL
    # in-place change of the grid tensor: torch's version counter invalidates the cached verdict
    x_t[:150] -= 0.25
    W2 = engine.wigner_t(rho_t, x_t).cpu().numpy()
    assert rel_err(W2, wigner_clenshaw(rho, x_t.cpu().numpy())) < TOL
    # fused frame loop on the device
    wl = workloads.config1(steps=96)
    psi_t = torch.from_numpy(wl["psi"]).to(dev)
    xs = torch.from_numpy(wl["xvec"]).to(dev)
    Wf = engine.state_to_wigner_t(psi_t, wl["traced"], xs).cpu().numpy()
    r = partial_trace_sv(wl["psi"][0], wl["traced"])
    assert rel_err(Wf[0], wigner_clenshaw(r, wl["xvec"])) < TOL


def test_symmetry_of_the_result_is_exact(engine):
    """the eight points of a unit come from one recurrence: W is exactly mirror-symmetric where rho makes it so"""
    N = 10
    rho = np.diag(np.arange(1.0, N + 1)).astype(np.complex128)  # diagonal rho: W depends on x^2 + p^2 only
    rho /= np.trace(rho)
    x = np.linspace(-4, 4, 64)
    W = np.array(engine.wigner(rho, x))
    assert np.array_equal(W, W[::-1, :]) and np.array_equal(W, W[:, ::-1]) and np.array_equal(W, W.T)


def test_grid_verdict_follows_the_tensor_not_its_address(engine):
    """The device entry points test each grid tensor once.  The caching allocator hands a freed grid's address to the
    next grid of the same size: a mirror verdict must not survive that (it would run the symmetric kernels on a
    grid that is not mirror-symmetric)."""
    import torch

    dev = torch.device("cuda", engine.device)
    rho = workloads.random_rho(12, seed=5)
    rho_t = torch.from_numpy(rho).to(dev)
    grids = [np.linspace(-5, 5, 64), np.linspace(-4, 5, 64), np.linspace(-5, 5, 64), np.linspace(-5, 4.5, 64)]
    seen = set()
    for x in grids:
        x_t = torch.from_numpy(x).to(dev)
        seen.add(x_t.data_ptr())
        W = engine.wigner_t(rho_t, x_t).cpu().numpy()
        assert rel_err(W, wigner_clenshaw(rho, x)) < TOL
        del x_t
    # in-place change of a grid that was judged mirror-symmetric
    x_t = torch.from_numpy(grids[0]).to(dev)
    engine.wigner_t(rho_t, x_t)
    x_t.copy_(torch.from_numpy(grids[1]).to(dev))
    W = engine.wigner_t(rho_t, x_t).cpu().numpy()
    assert rel_err(W, wigner_clenshaw(rho, grids[1])) < TOL
    assert len(seen) < len(grids)  # the allocator did recycle an address (otherwise this test shows nothing)


RESIDENT_VARIANTS = [8, 9, 11, 12, 13, 17, 22]
RING_VARIANTS = [10, 14, 15, 16, 18, 19, 20, 21]


@pytest.mark.parametrize("variant", RESIDENT_VARIANTS + RING_VARIANTS)
def test_every_symmetric_kernel_variant(engine, variant):
    """The library picks a variant by predicted duration, so some run only on particular grids: force each one
    (bq_set_sym_variant) on an even and an odd grid and check it against the oracle and the point-by-point kernels."""
    N = 24 if variant in RESIDENT_VARIANTS else 80
    rho = workloads.random_rho(N, seed=variant, rank=3)
    try:
        engine.set_sym_variant(variant)
        for S in (577, 640):
            x = np.linspace(-5.5, 5.5, S)
            before = engine.launches
            W = np.array(engine.wigner(rho, x))
            assert engine.launches > before
            rows = np.arange(0, S, 37)
            ref = wigner_clenshaw(rho, x, x[rows])
            assert rel_err(W[rows], ref) < TOL
            engine.set_grid_sharing(False)
            Wp = np.array(engine.wigner(rho, x))
            engine.set_grid_sharing(True)
            assert rel_err(W, Wp) < 1e-12
    finally:
        engine.set_grid_sharing(True)
        engine.set_sym_variant(None)


@pytest.mark.parametrize("variant", [19, 21])
def test_parked_ring_kernel_with_tiles_taken_in_runs(engine, variant):
    """Enough one-tile frames that the persistent CTAs take tiles from the counter in RUNS (grab >= 2): a CTA then moves to
    its next tile without the block-wide barrier of the counter path, and the ring refills at the start of a tile must not
    overtake warps still reading the previous tile's last groups (they once did: NaNs on a 4096 x 4096 grid)."""
    N, S, frames = 80, 64, 2500
    rng = np.random.default_rng(3)
    base = workloads.random_rho(N, seed=9, rank=3)
    rho = np.stack([base * (1.0 + 0.001 * k) for k in range(8)])
    rho = np.concatenate([rho] * (frames // 8 + 1))[:frames]
    x = np.linspace(-5, 5, S)
    try:
        engine.set_sym_variant(variant)
        W = np.asarray(engine.wigner(rho, x)).copy()
    finally:
        engine.set_sym_variant(None)
    assert np.isfinite(W).all()
    for k in range(8):
        ref = wigner_clenshaw(rho[k], x)
        for f in range(k, frames, 8 * 37):  # the same rho recurs every 8 frames: all copies must agree with the oracle
            assert np.abs(W[f] - ref).max() / np.abs(ref).max() < 1e-10, (k, f)
    # ... and every copy of a frame bit for bit with its first occurrence
    for k in range(8):
        assert all(np.array_equal(W[f], W[k]) for f in range(k, frames, 8)), k
