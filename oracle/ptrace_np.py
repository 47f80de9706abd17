"""NumPy restatement of ``qiskit.quantum_info.partial_trace`` (qiskit 2.1.2).

ORACLE -- test infrastructure, not product code (see ``oracle/__init__.py``).

Reference call sites this stands in for:
``src/bosonic_qiskit/util.py:580`` (trace_out_qumodes), ``:596``
(trace_out_qubits), ``:619`` (cv_partial_trace), ``:693`` (avg_photon_num).

Semantics restated (third-party, source not in /root/reference):
* Statevector branch: ``arr = psi.reshape((2,)*n)`` (axis k <-> qubit n-1-k,
  Qiskit little-endian), ``rho = tensordot(arr, conj(arr), axes=(T, T))`` with
  ``T = n-1-qargs``; the kept qubits keep their significance order.
* DensityMatrix branch: ``rho_keep[i, j] = sum_t rho[(i,t), (j,t)]``.
"""

from __future__ import annotations

import numpy as np


def _n_qubits(dim: int) -> int:
    n = int(dim).bit_length() - 1
    if dim <= 0 or (1 << n) != dim:
        raise ValueError(f"dimension {dim} is not a power of two")
    return n


def _check_qargs(qargs, n: int) -> list[int]:
    q = [int(x) for x in qargs]
    if len(set(q)) != len(q):
        raise ValueError("duplicate qubit index in qargs")
    for x in q:
        if x < 0 or x >= n:
            raise ValueError(f"qubit index {x} out of range for {n} qubits")
    return q


def partial_trace_sv(psi, qargs) -> np.ndarray:
    """rho_keep = Psi Psi^dagger over the traced qubits ``qargs`` of a statevector.

    Follows qiskit ``partial_trace`` (Statevector branch); returns a
    ``(D, D)`` complex128 array, ``D = 2**(n-len(qargs))``.  When every qubit
    is traced the scalar ``<psi|psi>`` is returned as a 1x1 array.
    """
    psi = np.ascontiguousarray(np.asarray(psi, dtype=np.complex128).reshape(-1))
    n = _n_qubits(psi.shape[0])
    q = _check_qargs(qargs, n)
    arr = psi.reshape((2,) * n) if n else psi.reshape(())
    axes = [n - 1 - x for x in q]
    rho = np.tensordot(arr, arr.conj(), axes=(axes, axes))
    d = 1 << (n - len(q))
    return np.asarray(rho).reshape(d, d)


def partial_trace_dm(rho, qargs) -> np.ndarray:
    """Partial trace of a density matrix over qubits ``qargs`` (DensityMatrix branch)."""
    rho = np.asarray(rho, dtype=np.complex128)
    if rho.ndim != 2 or rho.shape[0] != rho.shape[1]:
        raise ValueError("density matrix must be square")
    n = _n_qubits(rho.shape[0])
    q = _check_qargs(qargs, n)
    arr = rho.reshape((2,) * (2 * n))
    # row axis of qubit x is n-1-x, column axis is 2n-1-x
    keep = [x for x in range(n - 1, -1, -1) if x not in q]  # descending significance
    letters = {}
    nxt = iter("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
    sub_in = [None] * (2 * n)
    for x in range(n):
        if x in q:
            l = next(nxt)
            sub_in[n - 1 - x] = l
            sub_in[2 * n - 1 - x] = l
        else:
            lr, lc = next(nxt), next(nxt)
            letters[x] = (lr, lc)
            sub_in[n - 1 - x] = lr
            sub_in[2 * n - 1 - x] = lc
    out = "".join(letters[x][0] for x in keep) + "".join(letters[x][1] for x in keep)
    res = np.einsum("".join(sub_in) + "->" + out, arr)
    d = 1 << len(keep)
    return np.asarray(res).reshape(d, d)


def partial_trace(state, qargs) -> np.ndarray:
    """Dispatch on rank like the qiskit function dispatches on type."""
    a = np.asarray(getattr(state, "data", state))
    if a.ndim == 1:
        return partial_trace_sv(a, qargs)
    return partial_trace_dm(a, qargs)
