"""Drop-in for the partial-trace entry points of ``bosonic_qiskit.util``.

Same names, argument meaning and return types as the reference
(``src/bosonic_qiskit/util.py:548-619,672-736``); the two third-party calls
behind them -- ``qiskit.quantum_info.partial_trace`` at ``util.py:580,596,619,693``
-- are routed to the B200 library (K1 / K2 / K5 in ``csrc/ptrace_kernels.cu``).
``simulate()`` and everything else in the reference's ``util`` is unchanged and
is NOT re-implemented here.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from ._compat import DensityMatrix, Statevector, is_density_matrix, is_statevector
from .engine import get_engine


def partial_trace(state, qargs):
    """``qiskit.quantum_info.partial_trace(state, qargs)`` on the GPU.

    ``state``: Statevector / DensityMatrix (or a 1-D / 2-D array).  Returns a
    ``DensityMatrix`` with ``dims=(2,)*n_kept``; tracing every qubit returns the
    scalar trace, as the qiskit function does.
    """
    qargs = [int(q) for q in qargs]
    data = np.asarray(getattr(state, "data", state))
    eng = get_engine()
    if data.ndim == 1 and not is_density_matrix(state):
        n = int(data.shape[0]).bit_length() - 1
        _validate_qargs(qargs, n)
        rho = eng.partial_trace_sv(data, qargs)
    elif data.ndim == 2:
        n = int(data.shape[0]).bit_length() - 1
        _validate_qargs(qargs, n)
        rho = eng.partial_trace_dm(data, qargs)
    else:
        raise ValueError("state must be a statevector or a density matrix")
    if len(qargs) == n:
        return rho.reshape(()).item()
    return DensityMatrix(rho, dims=(2,) * (n - len(qargs)))


def _validate_qargs(qargs, n):
    if len(set(qargs)) != len(qargs):
        raise ValueError("Duplicate qargs")  # qiskit raises QiskitError for the same misuse
    for q in qargs:
        if q < 0 or q >= n:
            raise ValueError(f"qargs {q} out of range for {n} qubits")


def _find_qubit_indices(circuit) -> list[int]:
    """Indices of the circuit's qubits that are NOT in a QumodeRegister (``util.py:548-564``)."""
    to_exclude = set(circuit.qumode_qubits)
    indices = []
    for index, qubit in enumerate(circuit.qubits):
        if qubit not in to_exclude:
            indices.append(index)
    return indices


def trace_out_qumodes(circuit, state_vector):
    """Reduced density matrix of the qubits: trace out every qumode (``util.py:567-580``)."""
    indices = circuit.qumode_qubit_indices
    return partial_trace(state_vector, indices)


def trace_out_qubits(circuit, state_vector):
    """Reduced density matrix of the qumodes: trace out every plain qubit (``util.py:583-596``)."""
    indices = _find_qubit_indices(circuit)
    return partial_trace(state_vector, indices)


def cv_partial_trace(circuit, state_vector, qubits):
    """Reduced density matrix after tracing the given Qiskit qubits (``util.py:599-619``)."""
    if not isinstance(qubits, Sequence):
        qubits = [qubits]
    indices = circuit.get_qubit_indices(qubits)
    return partial_trace(state_vector, indices)


def mean_trace_out_qubits(circuit, state_vectors, weights=None):
    """Shot-averaged reduced density matrix of the qumodes: ``sum_k w_k trace_out_qubits(circuit, psi_k)`` over the
    per-shot statevectors ``simulate(..., per_shot_state_vector=True)`` returns (``util.py:523-532``), the quantity
    ``tutorials/bosonic-kitten-code`` cell 14 builds with a Python loop over shots.  One batched partial-trace launch
    with the accumulation fused into its reduction (``bq_partial_trace_sv_accumulate_host``); ``weights=None`` is the
    plain mean.  Returns a ``DensityMatrix`` -- the mixed-state input of ``wigner.wigner`` under noise."""
    indices = _find_qubit_indices(circuit)
    psi = np.stack([np.asarray(getattr(sv, "data", sv), dtype=np.complex128).reshape(-1) for sv in state_vectors])
    n = int(psi.shape[1]).bit_length() - 1
    _validate_qargs(indices, n)
    rho = get_engine().partial_trace_sv_mean(psi, indices, weights)
    return DensityMatrix(rho, dims=(2,) * (n - len(indices)))


def avg_photon_num(circuit, state, decimals: int = 2) -> list[float]:
    """Average photon number of each qumode (``util.py:672-696``).

    The reference runs one partial trace per qumode over the same state
    (``util.py:688-694``); for a Statevector these are issued as ONE batched
    launch (``bq_partial_trace_sv_multi``), then one <n> reduction.
    """
    qumode_qubits = circuit.qumode_qubits_indices_grouped
    traced_sets = []
    for qumode in range(len(qumode_qubits)):
        traced = []
        for traced_qumode in range(len(qumode_qubits)):
            if traced_qumode != qumode:
                traced.extend(qumode_qubits[traced_qumode])
        traced_sets.append(traced)
    if not traced_sets:
        return []

    eng = get_engine()
    data = np.asarray(getattr(state, "data", state))
    same_len = len({len(t) for t in traced_sets}) == 1
    if data.ndim == 1 and same_len:
        rhos = eng.partial_trace_sv_multi(data, traced_sets)
        values = eng.photon_number(rhos)
        return [_round_photon(v, decimals) for v in values]
    # qumodes of different cutoffs, or a DensityMatrix input: one trace per qumode
    averages = []
    for traced in traced_sets:
        traced_state = partial_trace(state, traced)
        averages.append(qumode_avg_photon_num(traced_state, decimals))
    return averages


def _round_photon(avg_photon: complex, decimals: int):
    if round(avg_photon.imag, 6) != 0:
        raise Exception(
            "Magnitude of average photon is complex, check inputs. Imaginary portion = {}".format(avg_photon.imag)
        )
    return np.round(avg_photon.real, decimals)


def qumode_avg_photon_num(state, decimals: int = 2) -> float:
    """<n> of one qumode's state via the number operator (``util.py:699-736``)."""
    eng = get_engine()
    if is_statevector(state):
        rho = eng.partial_trace_sv(np.asarray(state.data), [])  # |psi><psi|; <n> divides by the norm
    elif is_density_matrix(state):
        rho = np.asarray(state.data)
    else:
        raise TypeError("Only Statevector or DensityMatrix are accepted as valid types.")
    return _round_photon(complex(eng.photon_number(rho)), decimals)


__all__ = [
    "partial_trace",
    "trace_out_qumodes",
    "trace_out_qubits",
    "cv_partial_trace",
    "mean_trace_out_qubits",
    "avg_photon_num",
    "qumode_avg_photon_num",
    "DensityMatrix",
    "Statevector",
]
