"""Drop-in API cases shared by the CPU (oracle-backed engine) and GPU (CUDA engine) suites.

They replay the INPUTS of the reference's own hot-path tests (tests/test_util.py:11-87,216-261,
tests/test_circuit.py:80-108,326-347, tests/test_wigner.py) with states built directly in the
reference's index layout (no Aer in this image), and assert the same things the reference asserts --
plus numerical values, which the reference does not pin.
"""

from __future__ import annotations

import os

import numpy as np
import pytest

import workloads
from bosonic_qiskit_b200 import util, wigner
from bosonic_qiskit_b200._compat import DensityMatrix, Statevector
from fakes import FakeCVCircuit, FakeQuantumRegister, FakeQumodeRegister
from oracle import analytic as an
from oracle import partial_trace_sv, wigner_clenshaw


def basis_state(n_qubits: int, index: int) -> Statevector:
    v = np.zeros(1 << n_qubits, dtype=np.complex128)
    v[index] = 1.0
    return Statevector(v)


def make_1mode_1qubit(nq_per_mode=2):
    qmr = FakeQumodeRegister(1, nq_per_mode)
    qr = FakeQuantumRegister(1)
    return qmr, qr, FakeCVCircuit(qmr, qr)


def case_get_qubit_indices():
    # tests/test_circuit.py:80-108
    qmr = FakeQumodeRegister(2, 2)
    qbr = FakeQuantumRegister(2)
    circuit = FakeCVCircuit(qmr, qbr)
    assert circuit.get_qubit_indices([qmr[0]]) == [0, 1]
    assert circuit.get_qubit_indices([qmr[1]]) == [2, 3]
    assert circuit.get_qubit_indices([qbr[0]]) == [4]
    assert circuit.get_qubit_indices([qbr[1]]) == [5]
    assert util._find_qubit_indices(circuit) == [4, 5]
    assert circuit.qumode_qubit_indices == [0, 1, 2, 3]


def case_trace_out_zero():
    # tests/test_util.py:11-26: qubit |1>, qumode |0>  => index = 1*4 + 0
    qmr, qr, circuit = make_1mode_1qubit()
    state = basis_state(3, 4)
    trace = util.trace_out_qubits(circuit, state)
    assert state.dims() == (2, 2, 2)
    assert trace.dims() == (2, 2)
    np.testing.assert_almost_equal(trace.probabilities_dict()["00"], 1.0)


def case_trace_out_one():
    # tests/test_util.py:33-48: qubit |0>, qumode |1>
    qmr, qr, circuit = make_1mode_1qubit()
    state = basis_state(3, 1)
    trace = util.trace_out_qubits(circuit, state)
    assert trace.dims() == (2, 2)
    np.testing.assert_almost_equal(trace.probabilities_dict()["01"], 1.0)


def case_trace_out_qubit_and_qumode():
    # tests/test_util.py:55-87
    qmr, qbr, circuit = make_1mode_1qubit()
    state = basis_state(3, 1)
    trace = util.cv_partial_trace(circuit, state, qbr[0])
    assert trace.dims() == (2, 2)
    np.testing.assert_almost_equal(trace.probabilities_dict()["01"], 1.0)
    trace_q = util.cv_partial_trace(circuit, state, qmr[0])
    assert trace_q.dims() == (2,)
    np.testing.assert_almost_equal(trace_q.probabilities_dict()["0"], 1.0)
    trace_m = util.trace_out_qumodes(circuit, state)
    assert np.allclose(trace_m.data, trace_q.data)


def case_notebook_golden():
    # tutorials/wigner-function/wigner-function.ipynb:309-313,353-361
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_cutoff4.npz"))
    qmr, qr, circuit = make_1mode_1qubit()
    dm = util.trace_out_qubits(circuit, Statevector(g["psi"]))
    assert np.abs(dm.data - g["rho"]).max() < 2e-8


def case_avg_photon_num():
    # tests/test_util.py:216-232: |3>, |4>, |5> on three qumodes in two registers (3 qubits per mode)
    qmr1 = FakeQumodeRegister(2, 3)
    qmr2 = FakeQumodeRegister(1, 3)
    circ = FakeCVCircuit(qmr1, qmr2)
    state = basis_state(9, 3 + (4 << 3) + (5 << 6))
    assert np.allclose(util.avg_photon_num(circ, state), [3, 4, 5])
    # DensityMatrix input takes the one-trace-per-qumode route (util.py:693 DensityMatrix branch)
    small = FakeCVCircuit(FakeQumodeRegister(2, 2))
    sv = basis_state(4, 2 + (1 << 2))
    assert np.allclose(util.avg_photon_num(small, DensityMatrix(sv)), [2, 1])
    # a superposition, against the closed form
    psi = workloads.random_state(1 << 4, seed=11)
    got = util.avg_photon_num(small, Statevector(psi), decimals=6)
    p = np.abs(psi.reshape(4, 4)) ** 2  # [n1, n0]
    want = [np.round((p.sum(0) * np.arange(4)).sum(), 6), np.round((p.sum(1) * np.arange(4)).sum(), 6)]
    assert np.allclose(got, want)


def case_qumode_avg_photon_num():
    # tests/test_util.py:235-261
    rng = np.random.default_rng(5)
    for _ in range(5):
        decimals = int(rng.integers(1, 6))
        dim = 8
        vector = rng.uniform(-1, 1, dim) + 1.0j * rng.uniform(-1, 1, dim)
        element_norm = (vector * vector.conj()).real
        avg_num = float(np.dot(element_norm, np.arange(dim)) / element_norm.sum())
        a = util.qumode_avg_photon_num(Statevector(vector), decimals)
        b = util.qumode_avg_photon_num(DensityMatrix(vector), decimals)
        assert a == b == np.round(avg_num, decimals)
    with pytest.raises(TypeError):
        util.qumode_avg_photon_num(np.ones(4), 2)


def case_sqr_like_qubit_expectation():
    # tests/test_circuit.py:326-347: qubit rotated by theta_n conditioned on Fock |n>, <Z> = cos(theta_n)
    m = FakeQumodeRegister(1, 4)
    q = FakeQuantumRegister(1)
    qc = FakeCVCircuit(m, q)
    cutoff = 16
    theta = np.pi / cutoff * np.arange(cutoff)
    for n in range(cutoff):
        psi = np.zeros(2 * cutoff, dtype=np.complex128)
        psi[n] = np.cos(theta[n] / 2)  # qubit |0>
        psi[cutoff + n] = np.sin(theta[n] / 2)  # qubit |1>
        dm = util.trace_out_qumodes(qc, Statevector(psi))
        z = (dm.data[0, 0] - dm.data[1, 1]).real
        assert np.isclose(z, np.cos(theta[n]))


def case_wigner_entry_points():
    x = np.linspace(-6, 6, 200)
    # Statevector input (wigner.py:123-124)
    c = an.cat_amplitudes(2.0, 16, +1)
    W = wigner.wigner(Statevector(c))
    assert W.shape == (200, 200) and W.dtype == np.float64
    assert np.abs(W - wigner_clenshaw(np.outer(c, c.conj()), x)).max() < 1e-12
    # DensityMatrix input
    rho = workloads.random_rho(8, seed=1)
    W2 = wigner.wigner(DensityMatrix(rho), axes_min=-4, axes_max=4, axes_steps=50)
    assert np.abs(W2 - wigner_clenshaw(rho, np.linspace(-4, 4, 50))).max() < 1e-12
    # raw 1-D array: np.outer(state, state) WITHOUT conjugation (wigner.py:130-131)
    v = an.coherent_amplitudes(0.5 + 0.8j, 16)
    W3 = wigner.wigner(v, axes_steps=64)
    assert np.abs(W3 - wigner_clenshaw(np.outer(v, v), np.linspace(-6, 6, 64))).max() < 1e-12
    # raw 2-D array
    W4 = wigner.wigner(rho, axes_steps=32, g=2.0)
    assert np.abs(W4 - wigner_clenshaw(rho, np.linspace(-6, 6, 32), g=2.0)).max() < 1e-12
    # _wigner: yvec is effectively xvec (wigner.py:187-188)
    W5 = wigner._wigner(DensityMatrix(rho), np.linspace(-3, 3, 20))
    assert W5.shape == (20, 20)
    with pytest.raises(ValueError):
        wigner._wigner(DensityMatrix(rho), np.linspace(-3, 3, 20), yvec=np.linspace(-1, 1, 4))
    with pytest.raises(TypeError):
        wigner.wigner(rho, method="bogus")
    # method="fft": the tuple (W, yvec) qutip.wigner returns for it, on the p axis derived from the x grid
    from oracle import wigner_fft

    Wf, yf = wigner.wigner(rho, axes_min=-5, axes_max=5, axes_steps=40, method="fft")
    Wo, yo = wigner_fft(np.asarray(rho), np.linspace(-5, 5, 40))
    assert Wf.shape == (40, 40) and yf.shape == (40,)
    np.testing.assert_allclose(yf, yo, rtol=1e-14, atol=0)
    assert np.abs(Wf - Wo).max() <= 1e-10 * np.abs(Wo).max()


def case_wigner_projections():
    # wigner.py:290-375 -- qubit (|0>+|1>)/sqrt2 entangled with two coherent states of one 16-level qumode
    qmr, qr, circuit = make_1mode_1qubit(nq_per_mode=4)
    a0, a1 = an.coherent_amplitudes(1.0 + 0.5j, 16), an.coherent_amplitudes(-1.2j, 16)
    x = np.concatenate([a0, a1]) / np.sqrt(2)  # qubit is the most significant bit
    y_z = np.concatenate([a0, -a1]) / np.sqrt(2)  # Z on the qubit
    y_x = np.concatenate([a1, a0]) / np.sqrt(2)  # X on the qubit
    xvec = np.linspace(-5, 5, 96)
    got = wigner.wigner_projections(circuit, Statevector(x), Statevector(y_z), Statevector(y_x), xvec)
    want = []
    for y in (y_z, y_x):
        t = [partial_trace_sv(u * v.conj(), [4]) for u in (x, y) for v in (x, y)]  # xx, xy, yx, yy
        want.append((t[0] + t[1] + t[2] + t[3]) / 4)
        want.append((t[0] - t[1] - t[2] + t[3]) / 4)
    assert len(got) == 4
    for g, rho in zip(got, want):
        ref = wigner_clenshaw(rho, xvec)
        assert g.shape == (96, 96)
        assert np.abs(g - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300)
    # default grid is the reference's linspace(-6, 6, 200); unknown kwargs fail like _wigner(**kwargs) would
    assert wigner.wigner_projections(circuit, x, y_z, y_x)[0].shape == (200, 200)
    with pytest.raises(TypeError):
        wigner.wigner_projections(circuit, x, y_z, y_x, bogus=1)


def case_wigner_mle():
    rng = np.random.default_rng(3)
    base = an.coherent_amplitudes(1.0, 8)
    shots = [Statevector(base + 0.01 * (rng.standard_normal(8) + 1j * rng.standard_normal(8))) for _ in range(6)]
    W = wigner.wigner_mle(shots, axes_steps=40)
    mean = np.stack([s.data for s in shots]).mean(0)
    mean /= np.linalg.norm(mean)
    assert W is not None  # what the reference asserts (tests/test_wigner.py:149-150)
    assert np.abs(W - wigner_clenshaw(np.outer(mean, mean), np.linspace(-6, 6, 40))).max() < 1e-12


class _FakeResult:
    def __init__(self, frames, label):
        self.results = [object()]
        self._data = {f"{label}{k}": Statevector(f) for k, f in enumerate(frames)}

    def data(self):
        return self._data


def case_simulate_paths(monkeypatch):
    """simulate_wigner / simulate_wigner_multiple_statevectors with the Aer call replaced by canned states."""
    qmr, qr, circuit = make_1mode_1qubit(nq_per_mode=3)  # cutoff 8 + 1 qubit
    wl = workloads.config3(frames=5, steps=40)
    frames = [np.concatenate([f[:8], f[32:40]]) for f in wl["psi"]]  # cut to cutoff 8 (+ qubit block)
    frames = [f / np.linalg.norm(f) for f in frames]
    x = np.linspace(-5, 5, 40)

    def fake_simulate(circ, shots=1, noise_passes=None, conditional_state_vector=False, return_fockcounts=True, **kw):
        if conditional_state_vector:
            return {"0x0": Statevector(frames[0])}, _FakeResult(frames, "segment_"), None
        return Statevector(frames[0]), _FakeResult(frames, "segment_"), None

    monkeypatch.setattr(wigner, "_simulate_impl", fake_simulate)
    W, state = wigner.simulate_wigner(circuit, x, shots=1, trace=True)
    ref0 = wigner_clenshaw(partial_trace_sv(frames[0], [3]), x)
    assert np.abs(W - ref0).max() < 1e-12 and np.allclose(state.data, frames[0])
    Wc, _ = wigner.simulate_wigner(circuit, x, shots=1, conditional_state="0x0", trace=True)
    assert np.abs(Wc - ref0).max() < 1e-12
    # trace=False: the whole register is one fictitious mode of dimension 16 (quirk 3)
    Wn, _ = wigner.simulate_wigner(circuit, x, shots=1, trace=False)
    assert np.abs(Wn - wigner_clenshaw(np.outer(frames[0], frames[0].conj()), x)).max() < 1e-12
    Ws = wigner.simulate_wigner_multiple_statevectors(circuit, x, 1, "segment_", 5, trace=True)
    assert isinstance(Ws, list) and len(Ws) == 5
    for k in range(5):
        assert np.abs(Ws[k] - wigner_clenshaw(partial_trace_sv(frames[k], [3]), x)).max() < 1e-12

    def no_state(*a, **k):
        return None, _FakeResult([], "segment_"), None

    monkeypatch.setattr(wigner, "_simulate_impl", no_state)
    with pytest.raises(RuntimeError, match="No state vector returned"):
        wigner.simulate_wigner(circuit, x, shots=1)


def case_animate_frame_loops(monkeypatch):
    """The frame producers of animate_wigner, with discretisation + Aer replaced by canned circuits/states."""
    from bosonic_qiskit_b200 import animate

    qmr, qr, circuit = make_1mode_1qubit(nq_per_mode=3)
    rng = np.random.default_rng(0)
    frames = [workloads.random_state(16, seed=s) for s in range(4)]

    def fake_single(circuit, segments_per_gate, epsilon, sequential_subcircuit, statevector_per_segment,
                    statevector_label, noise_passes):
        return circuit, len(frames)

    def fake_simulate(circ, shots=1, noise_passes=None, **kw):
        return Statevector(frames[0]), _FakeResult(frames, "segment_"), None

    monkeypatch.setattr(animate, "_discretize_single_circuit_impl", fake_single)
    monkeypatch.setattr(wigner, "_simulate_impl", fake_simulate)
    w_fock, xvec = animate.wigner_frames_without_measure(circuit, animation_segments=4, axes_steps=30)
    assert len(w_fock) == 4 and w_fock[0].shape == (30, 30)
    for k in range(4):
        assert np.abs(w_fock[k] - wigner_clenshaw(partial_trace_sv(frames[k], [3]), xvec)).max() < 1e-12
