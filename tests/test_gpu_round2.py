"""Round-2 entry points through the C ABI against the CPU oracle: the fused prologue (trace + pack inside the Wigner
launch), several traces of one statevector -> Wigner, the shot-averaged partial trace, the single-process multi-GPU
entry points (bq_multi_*) and the load-time self-check of the re-scheduled SASS.

Tolerance (north_star): max|dW| / max|W| <= 1e-10 and max|drho| / max|rho| <= 1e-10 in fp64.
"""

import os

import numpy as np
import pytest

import workloads
from oracle import partial_trace_sv, wigner_clenshaw

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def random_psi(n_qubits, batch, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=(batch, 1 << n_qubits)) + 1j * rng.normal(size=(batch, 1 << n_qubits))
    return psi / np.linalg.norm(psi, axis=1, keepdims=True)


# ------------------------------------------------------------------------------------------------ fused prologue
def test_sass_self_check_passed(engine):
    bad, checked = engine.sass_self_check()
    assert bad == 0, f"re-scheduled SASS of variants {bad:#x} is not bit-identical to nvcc's schedule"
    if engine.set_sass_mode(True):
        assert checked >= 18


@pytest.mark.parametrize("n_qubits,traced,S,batch", [
    (5, [4], 64, 1),        # c1 shape: one thread per element (2 traced indices)
    (5, [0], 33, 3),        # traced bit below the kept ones, odd grid
    (6, [1, 4], 48, 2),     # two traced bits in the middle
    (9, [0, 1, 2, 3, 4, 5], 40, 2),   # 64 traced indices: one warp per element
    (10, [3, 4, 5, 6, 7, 8, 9], 56, 1),  # 128 traced indices, cutoff 8
    (7, [], 40, 2),         # nothing traced: rho = |psi><psi| at cutoff 128 (ring kernels)
    (4, [0, 1, 2, 3], 16, 2),  # everything traced: cutoff 1
    # few kept, >= 512 traced indices: slices of Psi staged in shared memory (pro_mode 4)
    (13, list(range(9)), 40, 2),            # cutoff 16 kept in the top bits (lanes along the traced slice), 512 traced
    (13, list(range(4, 13)), 33, 3),        # cutoff 16 kept in the lowest bits (lanes along the kept index)
    (12, [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11], 24, 2),   # cutoff 2, 2048 traced
    (14, list(range(2, 13)), 32, 1),        # cutoff 8 from bits 0, 1 and 13; 2048 traced
    (15, list(range(1, 14)), 24, 1),        # cutoff 4, 8192 traced: the largest this mode takes
    (18, list(range(1, 17)), 24, 1),        # cutoff 4, 65536 traced: the separate trace launch
    (15, list(range(3, 13)), 40, 2),        # cutoff 32 from bits 0-2 and 13-14, 1024 traced (slices of 128)
])
def test_fused_trace_vs_oracle(engine, n_qubits, traced, S, batch):
    psi = random_psi(n_qubits, batch, seed=n_qubits * 100 + len(traced))
    x = np.linspace(-4, 4, S)
    W, rho = engine.state_to_wigner(psi, traced, x, return_rho=True)
    for b in range(batch):
        rho_ref = partial_trace_sv(psi[b], traced)
        assert rel_err(rho[b], rho_ref) < TOL
        assert rel_err(W[b], wigner_clenshaw(rho_ref, x)) < TOL
    # without the rho output, and on a grid that is not mirror-symmetric (point-by-point kernels)
    W2 = engine.state_to_wigner(psi, traced, x)
    assert np.array_equal(np.asarray(W2), np.asarray(W))
    xa, ya = np.linspace(-4, 3, S), np.linspace(-2, 4, S // 2 + 1)
    W3 = engine.state_to_wigner(psi, traced, xa, ya)
    assert rel_err(W3[0], wigner_clenshaw(partial_trace_sv(psi[0], traced), xa, ya)) < TOL


def test_fused_equals_separate_launches(engine):
    """BQ_B200_FUSED=0 runs trace (DMMA) + split reduction + pack + Wigner as separate launches: same results to
    rounding of rho (different summation order of the trace), and fewer launches with the prologue."""
    from bosonic_qiskit_b200 import Engine

    wl = workloads.config2(steps=96)
    os.environ["BQ_B200_FUSED"] = "0"
    try:
        plain = Engine(0)
    finally:
        del os.environ["BQ_B200_FUSED"]
    n0 = plain.launches
    Wp, rp = plain.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], return_rho=True)
    sep = plain.launches - n0
    n1 = engine.launches
    Wf, rf = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], return_rho=True)
    fused = engine.launches - n1
    plain.close()
    assert rel_err(rf, rp) < 1e-13 and rel_err(Wf, Wp) < 1e-12
    assert fused < sep and fused <= 2  # (<= 2: the cutoff's constant table may be built by the first call)


def test_wigner_from_rho_is_one_launch(engine):
    rho = workloads.random_rho(32, seed=3)
    x = np.linspace(-5, 5, 64)
    engine.wigner(rho, x)
    n0 = engine.launches
    W = engine.wigner(np.stack([rho, rho.conj().T]), x)
    assert engine.launches - n0 == 1
    assert rel_err(W[0], wigner_clenshaw(rho, x)) < TOL


def test_fused_device_entry_point(engine):
    import torch

    wl = workloads.config3(frames=6, steps=64)
    dev = torch.device("cuda", engine.device)
    psi_t = torch.from_numpy(wl["psi"]).to(dev)
    x_t = torch.from_numpy(wl["xvec"]).to(dev)
    rho_t = torch.empty((6, wl["N"], wl["N"]), dtype=torch.complex128, device=dev)
    W = engine.state_to_wigner_t(psi_t, wl["traced"], x_t, rho_out=rho_t).cpu().numpy()
    Wh = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"])
    assert np.array_equal(W, np.asarray(Wh))
    for b in (0, 5):
        rho_ref = partial_trace_sv(wl["psi"][b], wl["traced"])
        assert rel_err(rho_t[b].cpu().numpy(), rho_ref) < TOL
        assert rel_err(W[b], wigner_clenshaw(rho_ref, wl["xvec"])) < TOL


# ------------------------------------------------------------------------------------------------ several traces -> Wigner
def test_state_to_wigner_multi_vs_oracle(engine):
    wl = workloads.config4(steps=48)
    psi = wl["psi"][0]
    W, rho = engine.state_to_wigner_multi(psi, wl["traced_sets"], wl["xvec"], return_rho=True)
    assert W.shape == (4, 48, 48) and rho.shape == (4, 16, 16)
    for k, traced in enumerate(wl["traced_sets"]):
        rho_ref = partial_trace_sv(psi, traced)
        assert rel_err(rho[k], rho_ref) < TOL
        assert rel_err(W[k], wigner_clenshaw(rho_ref, wl["xvec"])) < TOL
    # device entry point, one launch
    import torch

    dev = torch.device("cuda", engine.device)
    psi_t, x_t = torch.from_numpy(psi).to(dev), torch.from_numpy(wl["xvec"]).to(dev)
    engine.state_to_wigner_multi_t(psi_t, wl["traced_sets"], x_t)
    n0 = engine.launches
    Wt = engine.state_to_wigner_multi_t(psi_t, wl["traced_sets"], x_t)
    assert engine.launches - n0 == 1
    assert np.array_equal(Wt.cpu().numpy(), np.asarray(W))


# ------------------------------------------------------------------------------------------------ shot-averaged rho
@pytest.mark.parametrize("n_qubits,traced,shots", [(5, [4], 1), (5, [4], 7), (8, [0, 7], 33), (13, [6, 7, 8, 9, 10, 11, 12], 5)])
def test_partial_trace_mean_vs_oracle(engine, n_qubits, traced, shots):
    psi = random_psi(n_qubits, shots, seed=shots)
    ref = np.mean([partial_trace_sv(p, traced) for p in psi], axis=0)
    assert rel_err(engine.partial_trace_sv_mean(psi, traced), ref) < TOL
    w = np.random.default_rng(1).random(shots)
    ref_w = sum(wk * partial_trace_sv(p, traced) for wk, p in zip(w, psi))
    assert rel_err(engine.partial_trace_sv_mean(psi, traced, weights=w), ref_w) < TOL
    import torch

    dev = torch.device("cuda", engine.device)
    out = engine.partial_trace_sv_mean_t(torch.from_numpy(psi).to(dev), traced, weights=torch.from_numpy(w).to(dev))
    assert rel_err(out.cpu().numpy(), ref_w) < TOL


def test_mean_trace_out_qubits_dropin(engine):
    from bosonic_qiskit_b200 import util, wigner
    from fakes import FakeCVCircuit, FakeQuantumRegister, FakeQumodeRegister

    circuit = FakeCVCircuit(FakeQumodeRegister(1, 3), FakeQuantumRegister(1))
    psi = random_psi(4, 9, seed=2)
    dm = util.mean_trace_out_qubits(circuit, list(psi))
    ref = np.mean([partial_trace_sv(p, [3]) for p in psi], axis=0)
    assert rel_err(np.asarray(dm.data), ref) < TOL
    W = wigner.wigner(dm, axes_steps=40)
    assert rel_err(W, wigner_clenshaw(ref, np.linspace(-6, 6, 40))) < TOL


# ------------------------------------------------------------------------------------------------ several GPUs, one process
def _device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]  # a device listed more than once takes that many shares: the sharding logic on one GPU
    if n >= 2:
        lists.append(list(range(n)))
    return lists


@pytest.mark.parametrize("frames", [1, 2, 5, 12])
def test_multi_engine_frames_bit_identical(engine, frames):
    from bosonic_qiskit_b200 import MultiEngine

    wl = workloads.config3(frames=frames, steps=72)
    W1, rho1 = engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], return_rho=True)
    for devs in _device_lists():
        me = MultiEngine(devs)
        try:
            Wm, rhom = me.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"], return_rho=True)
            assert np.array_equal(np.asarray(Wm), np.asarray(W1)), devs
            assert np.array_equal(rhom, rho1), devs
            Wr = me.wigner(rho1, wl["xvec"])
            assert np.array_equal(np.asarray(Wr), np.asarray(engine.wigner(rho1, wl["xvec"]))), devs
        finally:
            me.close()
    assert rel_err(W1[0], wigner_clenshaw(partial_trace_sv(wl["psi"][0], wl["traced"]), wl["xvec"])) < TOL


@pytest.mark.parametrize("N,S", [(24, 96), (24, 97), (96, 64)])
def test_multi_engine_one_grid_bit_identical(engine, N, S):
    """BQ_SHARD_GRID: unit ranges stored into one frame buffer (mirror grid) or row slabs (any other grid)."""
    from bosonic_qiskit_b200 import MultiEngine

    rho = workloads.random_rho(N, seed=N + S)
    x = np.linspace(-5, 5, S)
    xa, ya = np.linspace(-5, 4, S), np.linspace(-3, 5, S - 3)
    W1 = np.asarray(engine.wigner(rho, x))
    engine.set_grid_sharing(False)
    try:
        W1a = np.asarray(engine.wigner(rho, xa, ya))
    finally:
        engine.set_grid_sharing(True)
    for devs in _device_lists():
        me = MultiEngine(devs)
        try:
            assert np.array_equal(np.asarray(me.wigner(rho, x, shard="grid")), W1), devs
            assert np.array_equal(np.asarray(me.wigner(rho, xa, ya, shard="grid")), W1a), devs
        finally:
            me.close()
    assert rel_err(W1, wigner_clenshaw(rho, x)) < TOL


def test_multi_engine_one_grid_from_psi(engine):
    from bosonic_qiskit_b200 import MultiEngine

    psi = random_psi(6, 1, seed=11)[0]
    x = np.linspace(-4, 4, 80)
    W1, rho1 = engine.state_to_wigner(psi, [5], x, return_rho=True)
    me = MultiEngine([0, 0])
    try:
        Wm, rhom = me.state_to_wigner(psi, [5], x, shard="grid", return_rho=True)
    finally:
        me.close()
    assert np.array_equal(np.asarray(Wm), np.asarray(W1)) and np.array_equal(rhom, rho1)


def test_multi_engine_errors(engine):
    from bosonic_qiskit_b200 import BQError, MultiEngine

    with pytest.raises(BQError):
        MultiEngine([99])
    me = MultiEngine([0, 0])
    try:
        with pytest.raises(BQError):  # one grid means one frame
            me.wigner(np.stack([np.eye(4) / 4] * 2), np.linspace(-3, 3, 16), shard="grid")
        with pytest.raises(BQError):  # a failing worker's message reaches the caller
            me.state_to_wigner(np.ones((2, 8), dtype=complex), [7], np.linspace(-3, 3, 16))
        with pytest.raises(ValueError):
            me.wigner(np.eye(4) / 4, np.linspace(-3, 3, 16), shard="rows")
    finally:
        me.close()


def test_frames_to_wigner_devices_dropin(engine):
    """The drop-in frame loop on several devices from one process (what animate_wigner calls)."""
    from bosonic_qiskit_b200 import wigner

    wl = workloads.config3(frames=9, steps=60)
    frames = list(wl["psi"])
    W1 = np.asarray(wigner.frames_to_wigner(frames, wl["traced"], wl["xvec"], devices=[0]))
    W2 = np.asarray(wigner.frames_to_wigner(frames, wl["traced"], wl["xvec"], devices=[0, 0]))
    Wd = np.asarray(wigner.frames_to_wigner(frames, wl["traced"], wl["xvec"]))  # devices=None: every visible GPU when it pays
    assert np.array_equal(W1, W2) and np.array_equal(W1, Wd)


def test_multi_gpu_all_devices(engine):
    """>= 2 GPUs: every device of the box, frames and one grid, bit-identical to one GPU."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs (the sharding logic itself runs above with a device listed twice)")
    from bosonic_qiskit_b200 import get_multi_engine

    me = get_multi_engine()
    assert me.n_devices == torch.cuda.device_count()
    wl = workloads.config3(frames=4 * me.n_devices + 1, steps=100)
    W1 = np.asarray(engine.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"]))
    assert np.array_equal(np.asarray(me.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"])), W1)
    rho = workloads.random_rho(128, seed=4)
    x = np.linspace(-5, 5, 256)
    assert np.array_equal(np.asarray(me.wigner(rho, x, shard="grid")), np.asarray(engine.wigner(rho, x)))


# ------------------------------------------------------------------------------------------------ snapshots
def test_snapshot_wigners_batched(engine):
    from bosonic_qiskit_b200 import wigner
    from fakes import FakeCVCircuit, FakeQuantumRegister, FakeQumodeRegister

    circuit = FakeCVCircuit(FakeQumodeRegister(1, 3), FakeQuantumRegister(1))
    circuit.cv_snapshot_id = 3
    psi = random_psi(4, 3, seed=5)

    class Result:
        def data(self):
            return {f"cv_snapshot_{k}": psi[k] for k in range(3)}

    out = wigner.snapshot_wigners(circuit, Result(), axes_steps=48)
    assert [label for label, _ in out] == ["cv_snapshot_0", "cv_snapshot_1", "cv_snapshot_2"]
    x = np.linspace(-6, 6, 48)
    for k, (_, W) in enumerate(out):
        assert rel_err(W, wigner_clenshaw(partial_trace_sv(psi[k], [3]), x)) < TOL


def test_local_prologue_equals_grid_wide_prologue(engine):
    """Tiny traces at a resident cutoff are done per CTA into shared memory (pro_mode 3); BQ_B200_LOCAL_PROLOGUE=0 keeps the
    grid-wide prologue: same summation order, bit-identical rho and W."""
    from bosonic_qiskit_b200 import Engine

    os.environ["BQ_B200_LOCAL_PROLOGUE"] = "0"
    try:
        ref_engine = Engine(0)
    finally:
        del os.environ["BQ_B200_LOCAL_PROLOGUE"]
    try:
        for n_qubits, traced, S, batch in ((5, [4], 200, 1), (5, [0], 33, 3), (6, [1, 4], 48, 2), (7, [6], 130, 8)):
            psi = random_psi(n_qubits, batch, seed=17 + n_qubits)
            x = np.linspace(-5, 5, S)
            W, rho = engine.state_to_wigner(psi, traced, x, return_rho=True)
            W0, rho0 = ref_engine.state_to_wigner(psi, traced, x, return_rho=True)
            assert np.array_equal(np.asarray(W), np.asarray(W0)) and np.array_equal(rho, rho0)
            assert rel_err(W[0], wigner_clenshaw(partial_trace_sv(psi[0], traced), x)) < TOL
    finally:
        ref_engine.close()
